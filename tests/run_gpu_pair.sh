#!/bin/bash
# CTA-pair GEMM bring-up, fail-fast: correctness + A/B timings first; the suite and the bench only when that passes
mkdir -p gpurun_out
O=gpurun_out
timeout 180 python tests/pair_bringup.py > $O/pair_bringup.log 2>&1; rc=$?
echo "pair bring-up exit=$rc"; tail -45 $O/pair_bringup.log
if [ $rc -ne 0 ]; then exit 1; fi
if [ -n "$QUICK" ]; then exit 0; fi
timeout 700 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; rc=$?; echo "pytest exit=$rc"; grep -v "^[0-9]* *$" $O/pytest_gpu.log | tail -8
if [ $rc -ne 0 ]; then exit 1; fi
timeout 400 python bench.py --steps ${STEPS:-8} --warmup 3 --cpu-baseline-chunks 0 > $O/bench.json 2> $O/bench.err; echo "bench exit=$?"; tail -3 $O/bench.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench.json"))
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "lat", d["latency"].get("p50_ms"), d["clocks"])
    for k, v in d["kernel_classes"].items():
        print(k, round(v["frac"], 3), v["launches_per_step"], round(v["ms_per_step"], 2))
except Exception as e:
    print("no bench", e)
PY

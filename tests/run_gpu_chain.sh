#!/bin/bash
# chain kernel bring-up, fail-fast: A/B bit-identity test first; the rest only runs when it passes
mkdir -p gpurun_out
O=gpurun_out
timeout 240 python -m pytest tests/test_gpu_parity.py -x -q -k "decode_chain" > $O/pytest_chain.log 2>&1; rc=$?
echo "chain A/B exit=$rc"; tail -15 $O/pytest_chain.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 600 python -m pytest tests -m gpu -x -q -k "not decode_chain" > $O/pytest_gpu.log 2>&1; rc=$?; echo "pytest exit=$rc"; grep -v "^[0-9]* *$" $O/pytest_gpu.log | tail -8
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

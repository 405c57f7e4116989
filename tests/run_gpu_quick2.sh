#!/bin/bash
# parity tests + GEMM micro-benchmark (stamps) + timeline + short bench
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest exit=$?"; tail -3 $O/pytest_gpu.log
timeout 300 python tests/gemm_bench.py > $O/gemm_bench.txt 2>&1; head -8 $O/gemm_bench.txt | cut -c1-330
timeout 600 python bench.py --timeline $O/timeline.txt --warmup 2 > $O/timeline.log 2>&1; echo "timeline exit=$?"
timeout 600 python bench.py --steps 4 --warmup 3 --latency-chunks 10 --cpu-baseline-chunks 0 > $O/bench.json 2> $O/bench.err; echo "bench exit=$?"; tail -2 $O/bench.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench.json"))
    print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "lat", d["latency"])
    for k, v in d["kernel_classes"].items():
        print(f"  {k:14s} {v['ms_per_step']:8.2f} ms  frac {v['frac']:.3f}  launches {v['launches_per_step']}")
except Exception as e:
    print("no bench json", e)
PY

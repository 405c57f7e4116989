#!/bin/bash
# parity tests, then bench with and without programmatic dependent launch
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest exit=$?"; tail -3 $O/pytest_gpu.log
for pdl in 1 0; do
  ISST_PDL=$pdl timeout 600 python bench.py --steps 4 --warmup 3 --latency-chunks 10 --cpu-baseline-chunks 0 > $O/bench_pdl$pdl.json 2> $O/bench_pdl$pdl.err; echo "bench pdl=$pdl exit=$?"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_pdl$pdl.json"))
    print("PDL=$pdl value", round(d["value"],1), "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), "lat p50", round(d["latency"]["p50_ms"],2))
except Exception as e:
    print("no bench json", e)
PY
done
cp $O/bench_pdl1.json $O/bench.json

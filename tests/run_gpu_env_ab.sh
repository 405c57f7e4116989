#!/bin/bash
# parity tests, then short benches under different values of one tuning env var: VAR=NAME VALS="a b c"
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest exit=$?"; tail -3 $O/pytest_gpu.log
for v in $VALS; do
  if [ "$v" = "default" ]; then unset $VAR; else export $VAR=$v; fi
  timeout 600 python bench.py --steps 4 --warmup 3 --latency-chunks ${LAT:-0} --cpu-baseline-chunks 0 > $O/bench_$v.json 2> $O/bench_$v.err; echo "bench $VAR=$v exit=$?"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_$v.json"))
    print("$VAR=$v value", round(d["value"],1), "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), "lat", d.get("latency"))
    print("   ", {k: round(x["ms_per_step"],2) for k, x in d["kernel_classes"].items()})
except Exception as e:
    print("no bench json", e)
PY
done

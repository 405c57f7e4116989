#!/bin/bash
# greedy bench at several batch sizes (streams per GPU): VALS="64 128 256"
mkdir -p gpurun_out
O=gpurun_out
for v in $VALS; do
  timeout 900 python bench.py --streams $v --steps 3 --warmup 3 --latency-chunks 0 --cpu-baseline-chunks 0 > $O/bench_s$v.json 2> $O/bench_s$v.err; echo "bench streams=$v exit=$?"; tail -2 $O/bench_s$v.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_s$v.json"))
    print("streams=$v value", round(d["value"],1), "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1))
    print("   ", {k: round(x["ms_per_step"],2) for k, x in d["kernel_classes"].items()})
except Exception as e:
    print("no bench json", e)
PY
done

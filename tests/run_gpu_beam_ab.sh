#!/bin/bash
# beam-4 bench under different values of one tuning env var: VAR=NAME VALS="a b c"
mkdir -p gpurun_out
O=gpurun_out
for v in $VALS; do
  if [ "$v" = "default" ]; then unset $VAR; else export $VAR=$v; fi
  timeout 600 python bench.py --beam 4 --streams ${BS:-64} --steps 3 --warmup 3 --latency-chunks 0 --cpu-baseline-chunks 0 > $O/bench_beam_$v.json 2> $O/bench_beam_$v.err; echo "bench $VAR=$v exit=$?"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_beam_$v.json"))
    print("$VAR=$v value", round(d["value"],1), "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1))
    print("   ", {k: round(x["ms_per_step"],2) for k, x in d["kernel_classes"].items()})
except Exception as e:
    print("no bench json", e)
PY
done

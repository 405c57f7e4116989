#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest exit=$?"; tail -3 $O/pytest_gpu.log
ISST_DEC_FUSE=0 timeout 600 python -m pytest tests -m gpu -x -q -k "tiny or sliding or golden or reference" > $O/pytest_gpu_nofuse.log 2>&1; echo "pytest(no fuse) exit=$?"; tail -2 $O/pytest_gpu_nofuse.log
for v in 1 0; do
  ISST_DEC_FUSE=$v timeout 600 python bench.py --steps 4 --warmup 3 --latency-chunks 10 --cpu-baseline-chunks 0 > $O/bench_fuse$v.json 2> $O/bench_fuse$v.err; echo "bench fuse=$v exit=$?"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_fuse$v.json"))
    print("FUSE=$v value", round(d["value"],1), "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), "lat p50", round(d["latency"]["p50_ms"],2))
    print("   ", {k: round(x["ms_per_step"],2) for k, x in d["kernel_classes"].items()})
except Exception as e:
    print("no bench json", e)
PY
done
cp $O/bench_fuse1.json $O/bench.json
timeout 600 python bench.py --timeline $O/timeline.txt --warmup 2 > $O/timeline.log 2>&1; echo "timeline exit=$?"

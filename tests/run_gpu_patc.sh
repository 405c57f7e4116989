#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
ISST_PREFILL_TC=1 timeout 400 python -m pytest tests -m gpu -x -q -k "tiny or sliding or golden or reference or batched or production or drift" > $O/pytest_patc.log 2>&1; echo "pytest(tc) exit=$?"; tail -12 $O/pytest_patc.log
for v in 1 0; do
  ISST_PREFILL_TC=$v timeout 150 python bench.py --steps 4 --warmup 3 --latency-chunks 0 --cpu-baseline-chunks 0 > $O/bench_patc$v.json 2> $O/bench_patc$v.err; echo "bench tc=$v exit=$?"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_patc$v.json"))
    print("TC=$v value", round(d["value"],1), "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1))
    print("   ", {k: round(x["ms_per_step"],2) for k, x in d["kernel_classes"].items()})
except Exception as e:
    print("no bench json", e)
PY
done

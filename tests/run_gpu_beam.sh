#!/bin/bash
# full GPU parity suite, then the bench with the reference's shipped decoding (--beam 4) and the greedy north-star line
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s > $O/pytest_gpu.log 2>&1; echo "pytest exit=$?"; tail -4 $O/pytest_gpu.log
timeout 900 python bench.py --beam 4 --streams ${BS:-64} --steps 4 --warmup 3 --latency-chunks 5 --cpu-baseline-chunks 0 > $O/bench_beam4.json 2> $O/bench_beam4.err; echo "bench beam exit=$?"; tail -3 $O/bench_beam4.err
python - <<'PY'
import json
for f in ("bench_beam4",):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1), "lat", d.get("latency"))
        print("   ", {k: round(x["ms_per_step"], 2) for k, x in d["kernel_classes"].items()})
    except Exception as e:
        print("no bench json", f, e)
PY

#!/bin/bash
# N-GPU bench exactly as the driver launches it (torchrun, one rank per GPU): N=${N:-8}
N=${N:-8}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/n${N}_gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 4 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "n$N exit=$?"
tail -3 gpurun_out/bench_n$N.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_n$N.json"))
    print("N=$N value", round(d["value"],1), "ms/step", round(d["ms_per_step"],2), "n_gpus", d["n_gpus"], "e2e", round(d["e2e"]["value"],1), "lat", d.get("latency"), "clocks", d.get("clocks"))
except Exception as e:
    print("no json", e)
PY

#!/bin/bash
# BASELINE.json configs[4]: 512 streams partitioned over N GPUs (strong scaling): N=${N:-2}
N=${N:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --total-streams 512 --steps 3 --warmup 3 --latency-chunks 0 --cpu-baseline-chunks 0 > gpurun_out/bench_cfg5_n$N.json 2> gpurun_out/bench_cfg5_n$N.err; echo "cfg5 n$N exit=$?"
tail -2 gpurun_out/bench_cfg5_n$N.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_cfg5_n$N.json"))
    print("cfg5 N=$N value", round(d["value"],1), "ms/step", round(d["ms_per_step"],2), "streams/gpu", d["config"]["streams_per_gpu"], "total", d["config"]["streams_total"], d["scaling"], "e2e", round(d["e2e"]["value"],1))
except Exception as e:
    print("no json", e)
PY

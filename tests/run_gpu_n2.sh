#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/n2_gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 4 --warmup 3 --latency-chunks 0 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "n2 exit=$?"
tail -3 gpurun_out/bench_n2.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_n2.json"))
    print("N=2 value", round(d["value"],1), "ms/step", round(d["ms_per_step"],2), "n_gpus", d["n_gpus"], "e2e", round(d["e2e"]["value"],1))
except Exception as e:
    print("no json", e)
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo "ref exit=$?"; cut -c1-300 gpurun_out/bench_ref_n2.json

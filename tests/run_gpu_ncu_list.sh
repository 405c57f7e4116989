#!/bin/bash
mkdir -p gpurun_out
ISST_PDL=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/launches.csv python bench.py --ncu-step --warmup 1 > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit=$?"

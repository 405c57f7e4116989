#!/bin/bash
# per-kernel times of one steady-state beam-4 step (ncu, serialised: compare shares)
mkdir -p gpurun_out
O=gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $O/launches_beam4.csv python bench.py --beam 4 --streams ${BS:-64} --ncu-step --warmup 1 > $O/ncu_launches_beam4.log 2>&1; echo "ncu exit=$?"
python tools/ncu_summarize.py $O/launches_beam4.csv 2>/dev/null | head -40

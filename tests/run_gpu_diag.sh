#!/bin/bash
# Runs every diagnostic stage under its own timeout so one hung kernel cannot eat the GPU lease.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/diag.log 2>&1
for spec in "simple:gemm_simple" "tc:gemm_tc" "simple:stream" "simple:stream_evict" "tc:stream" "tc:stream_evict"; do
  mode=${spec%%:*}; stage=${spec##*:}
  echo "=== ISST_GEMM=$mode stage=$stage ===" >> gpurun_out/diag.log
  ISST_GEMM=$mode timeout 300 python tests/gpu_diag.py $stage >> gpurun_out/diag.log 2>&1
  echo "exit=$?" >> gpurun_out/diag.log
done
tail -5 gpurun_out/diag.log

#!/bin/bash
# encoder attention bring-up, fail-fast
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -k "stream_vs_oracle_tiny or sliding" > gpurun_out/pytest_enc.log 2>&1; rc=$?
echo "enc tiny exit=$rc"; grep -v "^[0-9]* *$" gpurun_out/pytest_enc.log | tail -12
if [ $rc -ne 0 ]; then exit 1; fi
timeout 400 python -m pytest tests/test_enc_variants.py tests/test_gpu_parity.py -m gpu -x -q -k "variant or enc or latency_multipliers or update_multiplier or production or golden" > gpurun_out/pytest_enc2.log 2>&1; rc=$?
echo "enc more exit=$rc"; grep -v "^[0-9]* *$" gpurun_out/pytest_enc2.log | tail -8
if [ $rc -ne 0 ]; then exit 1; fi
bash tests/run_gpu_ab.sh "enc_attention_tc=1" "enc_attention_tc=0"

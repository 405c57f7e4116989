mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "not production and not gemm" > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
for s in 1 2 3; do ISST_DEC_SPLITS=$s python tests/decode_attn_bench.py; done 2>&1 | tee gpurun_out/dec_sweep.log
N_STREAMS=1 python tests/decode_attn_bench.py 2>&1 | tee -a gpurun_out/dec_sweep.log

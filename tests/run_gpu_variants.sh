#!/bin/bash
# encoder flag variants (--xpos 1 / --rope 0) and checkpoint ingestion on the GPU, then the whole GPU suite
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_enc_variants.py tests/test_checkpoint.py -m gpu -x -q -s > $O/pytest_variants.log 2>&1; echo "variants exit=$?"
tail -25 $O/pytest_variants.log
if [ "${FULL:-1}" = "1" ]; then
  timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest exit=$?"; tail -6 $O/pytest_gpu.log
fi

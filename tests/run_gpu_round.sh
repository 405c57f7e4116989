#!/bin/bash
# One gpurun call: smoke, GPU parity tests, bench line, ncu launch list + full captures of the top kernels.
# Every stage runs under its own timeout so one hung kernel cannot eat the lease.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $O/gpu.txt 2>&1
nproc >> $O/gpu.txt; free -g | head -2 >> $O/gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit=$?" | tee -a $O/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q -s > $O/pytest_gpu.log 2>&1; echo "pytest exit=$?" | tee -a $O/pytest_gpu.log
tail -5 $O/pytest_gpu.log
timeout 600 python bench.py --steps ${STEPS:-6} --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench exit=$?" | tee -a $O/bench.err
tail -3 $O/bench.err; cat $O/bench.json
if [ "${NCU:-1}" = "1" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
      --log-file $O/launches.csv python bench.py --ncu-step --warmup 1 > $O/ncu_launches.log 2>&1; echo "ncu launches exit=$?"
  # the fused decode-layer chain: one launch = o_proj -> RMSNorm -> gate/up -> down -> RMSNorm -> next QKV (33 per decode forward)
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
      -k regex:decode_chain -s 40 -c 3 -f -o $O/prof_gemm_decode python bench.py --ncu-step --warmup 1 --prime 3 > $O/ncu_gemm.log 2>&1; echo "ncu chain exit=$?"
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
      -k regex:decode_attention -s 32 -c 2 -f -o $O/prof_decode_attn python bench.py --ncu-step --warmup 1 > $O/ncu_attn.log 2>&1; echo "ncu attn exit=$?"
fi
ls -la $O
if [ "${NCU:-1}" = "1" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
      -k regex:prefill_attention_tc -s 2 -c 2 -f -o $O/prof_prefill_attn python bench.py --ncu-step --warmup 1 > $O/ncu_pattn.log 2>&1; echo "ncu prefill attn exit=$?"
  # tensor-bound GEMMs: the 4 GEMMs of one prefill layer (1408 tokens) on the CTA-pair kernel.  One step launches ~100
  # gemm_pair kernels before the LLM prefill (post_proj + 24 x 4 encoder + proj; the strided conv views stay on gemm_sk);
  # skip into prefill layer 1-2 (32 x 4 launches), any 4 consecutive launches there are one layer's o / gate-up / down / qkv
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
      -k regex:gemm_pair -s 104 -c 4 -f -o $O/prof_gemm_prefill python bench.py --ncu-step --warmup 1 --prime 3 > $O/ncu_gemm_prefill.log 2>&1; echo "ncu gemm prefill exit=$?"
  timeout 600 python bench.py --timeline $O/timeline.txt --warmup 2 > $O/timeline.log 2>&1; echo "timeline exit=$?"
  # beam search (the reference's shipped decoding): bench line + full capture of the shared-prefix group attention
  timeout 600 python bench.py --beam 4 --steps 4 --warmup 3 --latency-chunks 10 --cpu-baseline-chunks 0 > $O/bench_beam4.json 2> $O/bench_beam4.err; echo "bench beam4 exit=$?"
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
      -k regex:decode_attention_group -s 32 -c 1 -f -o $O/prof_group_attn python bench.py --beam 4 --ncu-step --warmup 1 > $O/ncu_gattn.log 2>&1; echo "ncu group attn exit=$?"
fi
# digest on the box (the raw reports of a full round exceed the 64 MiB gpurun copies back), keep only the summaries
if [ "${NCU:-1}" = "1" ]; then
  PROF_OUT=$O/profiles python tools/make_profiles.py ${TAG:-round} > $O/make_profiles.log 2>&1; echo "digest exit=$?"
  rm -f $O/*.ncu-rep
fi
du -sh $O

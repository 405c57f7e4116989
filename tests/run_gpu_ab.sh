#!/bin/bash
# A/B of per-context engine options on the bench step: bash tests/run_gpu_ab.sh "decode_splits=4" "decode_splits=0" ...
mkdir -p gpurun_out
for opt in "$@"; do
  tag=$(echo "$opt" | tr '= ' '__')
  args=""
  for kv in $opt; do args="$args --engine-opt $kv"; done
  timeout 300 python bench.py --steps 6 --warmup 3 --latency-chunks 6 --cpu-baseline-chunks 0 --eager-chunks 0 $args > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err
  python - "$opt" gpurun_out/ab_$tag.json <<'PY'
import json, sys
d = json.load(open(sys.argv[2]))
k = d["kernel_classes"]
print(f"{sys.argv[1]:32s} {d['value']:7.1f} speech-s/s {d['ms_per_step']:6.2f} ms  lat p50 {d['latency'].get('p50_ms', 0):5.1f} ms  attn_decode {k['attn_decode']['ms_per_step']:5.2f}  gemm_stream {k['gemm_stream']['ms_per_step']:5.2f}  attn_prefill {k['attn_prefill']['ms_per_step']:4.2f}  attn_enc {k['attn_encoder']['ms_per_step']:4.2f}  sm {d['clocks']['sm_mhz']}")
PY
done

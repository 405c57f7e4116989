// Memory-bound row kernels: conv0 + LayerNorm + GELU, LayerNorm / RMSNorm rows, embedding gather +
// speech splice, logits processors + arg-max.
#pragma once
#include "common.cuh"
#include "gemm_tcgen05.cuh"

namespace isst {

// ----------------------------------------------------------------------------------------------
// conv0: Conv1d(1 -> C, k, stride, bias) -> Fp32LayerNorm(C) -> GELU on the window [tail | new samples].
// fairseq ConvFeatureExtractionModel block 0, layer_norm mode (call site
// patch_speech_encoder.py:245-251).  Only the minimal window (RF-1 carried samples + the new chunk) is
// convolved instead of the reference's two-chunk ring: identical frames, half the work.
// One warp per output frame; weights transposed [k][C] in shared memory.
// ----------------------------------------------------------------------------------------------
constexpr int kConv0FramesPerCta = 64;
constexpr int kConv0MaxK = 16;

__global__ void __launch_bounds__(256)
conv0_ln_gelu_kernel(const float* __restrict__ pcm,      // [n][n_new] new samples (fp32, rounded to bf16 here
                                                         //  like `.to(dtype=bf16)`, agents/infinisst.py:222)
                     const float* __restrict__ tail,     // [max_streams][n_tail] carried samples (already rounded)
                     const int* __restrict__ slots, int n_new, int n_tail,
                     const float* __restrict__ w_t,      // [k][C]
                     const float* __restrict__ bias, const float* __restrict__ ln_w,
                     const float* __restrict__ ln_b, bf16* __restrict__ out,   // [n][T0][C]
                     int C, int k, int stride, int T0) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float c0_smem[];
  float* s_w = c0_smem;                 // k*C
  float* s_x = s_w + k * C;             // window span of this CTA
  const int b = blockIdx.y;
  const int slot = slots[b];
  const int f0 = blockIdx.x * kConv0FramesPerCta;
  const int nf = min(kConv0FramesPerCta, T0 - f0);
  const int span = (nf - 1) * stride + k;
  for (int i = threadIdx.x; i < k * C; i += blockDim.x) s_w[i] = w_t[i];
  for (int i = threadIdx.x; i < span; i += blockDim.x) {
    const int idx = f0 * stride + i;
    s_x[i] = idx < n_tail ? tail[static_cast<size_t>(slot) * n_tail + idx]
                          : bf16_round(pcm[static_cast<size_t>(b) * n_new + (idx - n_tail)]);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int per = C / 32;               // channels per lane (C % 32 == 0, <= 16)
  for (int f = warp; f < nf; f += nwarps) {
    float x[kConv0MaxK];
#pragma unroll
    for (int j = 0; j < kConv0MaxK; ++j) x[j] = j < k ? s_x[f * stride + j] : 0.f;
    float y[16];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (i < per) {
        const int ch = lane + 32 * i;
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < kConv0MaxK; ++j)
          if (j < k) acc += s_w[j * C + ch] * x[j];
        y[i] = bf16_round(acc + bias[ch]);        // conv output in the model dtype
        sum += y[i];
      }
    }
    const float mean = warp_sum(sum) / C;
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (i < per) { const float d = y[i] - mean; var += d * d; }
    const float rstd = rsqrtf(warp_sum(var) / C + 1e-5f);
    bf16* o = out + (static_cast<size_t>(b) * T0 + f0 + f) * C;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (i < per) {
        const int ch = lane + 32 * i;
        const float z = bf16_round((y[i] - mean) * rstd * ln_w[ch] + ln_b[ch]);   // Fp32LayerNorm(...).type_as(x)
        o[ch] = __float2bfloat16_rn(gelu_erf(z));
      }
    }
  }
}

// tail <- last n_tail samples of [tail | new]
__global__ void update_tail_kernel(const float* __restrict__ pcm, float* tail, const int* __restrict__ slots,
                                   int n_new, int n_tail) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x;
  const int slot = slots[b];
  float* t = tail + static_cast<size_t>(slot) * n_tail;
  // n_new >= n_tail in every supported configuration (chunk of 15360 samples vs 399 carried)
  for (int i = threadIdx.x; i < n_tail; i += blockDim.x)
    t[i] = bf16_round(pcm[static_cast<size_t>(b) * n_new + (n_new - n_tail + i)]);
}

// ----------------------------------------------------------------------------------------------
// Row normalisation: one CTA (128 threads) per row, up to 4096 columns held in registers.
//   kRms = false: LayerNorm(w, b);  kRms = true: LlamaRMSNorm (fp32 normalise -> round -> * w)
//   kGelu: GELU(erf) after the norm (conv feature extractor / length adapter blocks)
//   gather: optional row indices into `in` (last-token gather before the final norm + lm_head)
// ----------------------------------------------------------------------------------------------
// Deferred split-K reduction (weight-streaming GEMMs): when `part` is given, the row is first completed as
//   x = bf16( in + sum_s part[s][row] )        (fp32 partial sums of the preceding GEMM, in split order)
// and written back to `x_out` (the residual stream) before it is normalised - the GEMM then needs no
// cross-CTA reduction phase of its own.
struct DeferredSum {
  const float* part;        // [n_part][rows][C] fp32, or null
  int n_part;
  long long stride;         // rows * C
  bf16* x_out;              // completed rows (may alias `in`), or null
};

// kIts = 16-byte column groups per thread: 4 with 128 threads when there are many rows (bandwidth), 1 with up
// to 512 threads when there are few (the decode step's 64 rows are latency-bound: every load of a thread is
// issued up front, including all split partials).
template <bool kRms, bool kGelu, int kIts>
__global__ void __launch_bounds__(kIts == 1 ? 512 : 128)
norm_rows_kernel(const bf16* in, bf16* out, const float* __restrict__ w, const float* __restrict__ bvec,
                 const int* __restrict__ gather, int C, float eps, DeferredSum ds) {
  pdl_launch_dependents();
  const int tid = threadIdx.x;
  const int nthr = blockDim.x;
  // weights do not depend on the predecessor kernel: fetch them ahead of the dependency wait
  float wv[kIts][8], bv[kIts][8];
#pragma unroll
  for (int it = 0; it < kIts; ++it) {
    const int c0 = (it * nthr + tid) * 8;
    if (c0 < C) {
      const float4 w0 = *reinterpret_cast<const float4*>(w + c0), w1 = *reinterpret_cast<const float4*>(w + c0 + 4);
      wv[it][0] = w0.x; wv[it][1] = w0.y; wv[it][2] = w0.z; wv[it][3] = w0.w;
      wv[it][4] = w1.x; wv[it][5] = w1.y; wv[it][6] = w1.z; wv[it][7] = w1.w;
      if (!kRms) {
        const float4 b0 = *reinterpret_cast<const float4*>(bvec + c0), b1 = *reinterpret_cast<const float4*>(bvec + c0 + 4);
        bv[it][0] = b0.x; bv[it][1] = b0.y; bv[it][2] = b0.z; bv[it][3] = b0.w;
        bv[it][4] = b1.x; bv[it][5] = b1.y; bv[it][6] = b1.z; bv[it][7] = b1.w;
      }
    }
  }
  pdl_wait();
  const int row = blockIdx.x;
  const size_t src_row = gather ? gather[row] : row;
  const bf16* x = in + src_row * C;
  bf16* o = out + static_cast<size_t>(row) * C;
  float v[kIts][8];
  float sum = 0.f, sq = 0.f;
#pragma unroll
  for (int it = 0; it < kIts; ++it) {
    const int c0 = (it * nthr + tid) * 8;
    if (c0 < C) {
      uint4 raw = *reinterpret_cast<const uint4*>(x + c0);
      if (ds.part) {
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        const float* p0 = ds.part + src_row * C + c0;
        if (kIts == 1 && ds.n_part <= 8) {
          float4 pa[8], pb[8];
#pragma unroll
          for (int sp = 0; sp < 8; ++sp) {
            if (sp < ds.n_part) {
              const float4* pp = reinterpret_cast<const float4*>(p0 + sp * ds.stride);
              pa[sp] = __ldcg(pp); pb[sp] = __ldcg(pp + 1);
            }
          }
#pragma unroll
          for (int sp = 0; sp < 8; ++sp) {
            if (sp < ds.n_part) {
              acc[0] += pa[sp].x; acc[1] += pa[sp].y; acc[2] += pa[sp].z; acc[3] += pa[sp].w;
              acc[4] += pb[sp].x; acc[5] += pb[sp].y; acc[6] += pb[sp].z; acc[7] += pb[sp].w;
            }
          }
        } else {
          for (int sp = 0; sp < ds.n_part; ++sp) {
            const float4* pp = reinterpret_cast<const float4*>(p0 + sp * ds.stride);
            const float4 a = __ldcg(pp), b = __ldcg(pp + 1);
            acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w;
            acc[4] += b.x; acc[5] += b.y; acc[6] += b.z; acc[7] += b.w;
          }
        }
        const uint32_t ru[4] = {raw.x, raw.y, raw.z, raw.w};
        uint32_t nu[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = unpack_bf16(ru[j]);
          nu[j] = pack_bf16(acc[2 * j] + f.x, acc[2 * j + 1] + f.y);
        }
        raw = make_uint4(nu[0], nu[1], nu[2], nu[3]);
        if (ds.x_out) *reinterpret_cast<uint4*>(ds.x_out + src_row * C + c0) = raw;
      }
      const uint32_t u[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 f = unpack_bf16(u[j]);
        v[it][2 * j] = f.x; v[it][2 * j + 1] = f.y;
        sum += f.x + f.y;
        sq += f.x * f.x + f.y * f.y;
      }
    }
  }
  __shared__ float red[2][16];
  __shared__ float stat[2];
  const int nwarp = nthr >> 5;
  sum = warp_sum(sum); sq = warp_sum(sq);
  if ((tid & 31) == 0) { red[0][tid >> 5] = sum; red[1][tid >> 5] = sq; }
  __syncthreads();
  if (tid == 0) {
    float s = 0.f, q = 0.f;
    for (int i = 0; i < nwarp; ++i) { s += red[0][i]; q += red[1][i]; }
    if (kRms) { stat[0] = 0.f; stat[1] = rsqrtf(q / C + eps); }
    else { stat[0] = s / C; stat[1] = -1.f; }   // variance in a second pass below
  }
  __syncthreads();
  float mean = stat[0], rstd = stat[1];
  if (!kRms) {
    float var = 0.f;
#pragma unroll
    for (int it = 0; it < kIts; ++it) {
      const int c0 = (it * nthr + tid) * 8;
      if (c0 < C) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d = v[it][j] - mean; var += d * d; }
      }
    }
    var = warp_sum(var);
    __syncthreads();
    if ((tid & 31) == 0) red[0][tid >> 5] = var;
    __syncthreads();
    float tv = 0.f;
    for (int i = 0; i < nwarp; ++i) tv += red[0][i];
    rstd = rsqrtf(tv / C + eps);
  }
#pragma unroll
  for (int it = 0; it < kIts; ++it) {
    const int c0 = (it * nthr + tid) * 8;
    if (c0 < C) {
      uint32_t pk[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float y[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int k = 2 * j + e;
          float t;
          if (kRms) t = wv[it][k] * bf16_round(v[it][k] * rstd);
          else t = (v[it][k] - mean) * rstd * wv[it][k] + bv[it][k];
          if (kGelu) t = gelu_erf(bf16_round(t));
          y[e] = t;
        }
        pk[j] = pack_bf16(y[0], y[1]);
      }
      *reinterpret_cast<uint4*>(o + c0) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
  }
}

// Many rows, narrow rows (conv stack C = 512, encoder C = 1024): one WARP per row, kG 16-byte groups per lane,
// shuffles only - no shared memory, no block barrier, 8 rows per CTA and all of a row's loads in flight at once.
template <bool kRms, bool kGelu, int kG>
__global__ void __launch_bounds__(256)
norm_rows_warp_kernel(const bf16* in, bf16* out, const float* __restrict__ w,   // `in` may alias `out`
                      const float* __restrict__ bvec, int rows, float eps) {
  pdl_launch_dependents();
  constexpr int C = kG * 256;
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  float wv[kG][8], bv[kG][8];
#pragma unroll
  for (int g = 0; g < kG; ++g) {
    const int c0 = (g * 32 + lane) * 8;
    const float4 w0 = *reinterpret_cast<const float4*>(w + c0), w1 = *reinterpret_cast<const float4*>(w + c0 + 4);
    wv[g][0] = w0.x; wv[g][1] = w0.y; wv[g][2] = w0.z; wv[g][3] = w0.w; wv[g][4] = w1.x; wv[g][5] = w1.y; wv[g][6] = w1.z; wv[g][7] = w1.w;
    if (!kRms) {
      const float4 b0 = *reinterpret_cast<const float4*>(bvec + c0), b1 = *reinterpret_cast<const float4*>(bvec + c0 + 4);
      bv[g][0] = b0.x; bv[g][1] = b0.y; bv[g][2] = b0.z; bv[g][3] = b0.w; bv[g][4] = b1.x; bv[g][5] = b1.y; bv[g][6] = b1.z; bv[g][7] = b1.w;
    }
  }
  pdl_wait();
  if (row >= rows) return;
  const bf16* x = in + static_cast<size_t>(row) * C;
  uint4 raw[kG];
#pragma unroll
  for (int g = 0; g < kG; ++g) raw[g] = *reinterpret_cast<const uint4*>(x + (g * 32 + lane) * 8);
  float v[kG][8];
  float sum = 0.f, sq = 0.f;
#pragma unroll
  for (int g = 0; g < kG; ++g) {
    const uint32_t u[4] = {raw[g].x, raw[g].y, raw[g].z, raw[g].w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_bf16(u[j]);
      v[g][2 * j] = f.x; v[g][2 * j + 1] = f.y;
      sum += f.x + f.y;
      sq += f.x * f.x + f.y * f.y;
    }
  }
  float mean = 0.f, rstd;
  if (kRms) {
    rstd = rsqrtf(warp_sum(sq) / C + eps);
  } else {
    mean = warp_sum(sum) / C;
    float var = 0.f;
#pragma unroll
    for (int g = 0; g < kG; ++g)
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float d = v[g][j] - mean; var += d * d; }
    rstd = rsqrtf(warp_sum(var) / C + eps);
  }
  bf16* o = out + static_cast<size_t>(row) * C;
#pragma unroll
  for (int g = 0; g < kG; ++g) {
    uint32_t pk[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float y[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int k = 2 * j + e;
        float t;
        if (kRms) t = wv[g][k] * bf16_round(v[g][k] * rstd);
        else t = (v[g][k] - mean) * rstd * wv[g][k] + bv[g][k];
        if (kGelu) t = gelu_erf(bf16_round(t));
        y[e] = t;
      }
      pk[j] = pack_bf16(y[0], y[1]);
    }
    *reinterpret_cast<uint4*>(o + (g * 32 + lane) * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
}

// ----------------------------------------------------------------------------------------------
// Embedding gather + speech splice (SpeechLlamaModel.forward, llm.py:86-115): row r takes
// speech[speech_row[r]] when speech_row[r] >= 0 (the <sp_patch> slots), else embed[ids[r]].
// ----------------------------------------------------------------------------------------------
__global__ void embed_splice_kernel(const int* __restrict__ ids, const int* __restrict__ speech_row,
                                    const bf16* __restrict__ embed, const bf16* __restrict__ speech,
                                    bf16* __restrict__ out, int D) {
  pdl_launch_dependents();
  pdl_wait();
  const int r = blockIdx.x;
  const int sr = speech_row ? speech_row[r] : -1;
  const bf16* src = sr >= 0 ? speech + static_cast<size_t>(sr) * D : embed + static_cast<size_t>(ids[r]) * D;
  for (int c = threadIdx.x * 8; c < D; c += blockDim.x * 8)
    *reinterpret_cast<uint4*>(out + static_cast<size_t>(r) * D + c) = *reinterpret_cast<const uint4*>(src + c);
}

// ----------------------------------------------------------------------------------------------
// Device-side greedy step: HF processor order RepetitionPenalty ->
// NoRepeatNGram -> EncoderNoRepeatNGram -> SuppressTokens, then arg-max (lowest index wins ties),
// EOS / max-length bookkeeping.  kSelParts CTAs per stream; the host only ever sees token ids.
// ----------------------------------------------------------------------------------------------
struct GenState {
  int* ctx_ids;        // [n][ctx_cap]  prompt of this call + generated tokens
  int* ctx_len;        // [n]
  const int* enc_ids;  // [n][enc_cap]  last `lookback` emitted target ids
  const int* enc_len;  // [n]
  int* active;         // [n] 1 while the stream is still generating
  int* out_tokens;     // [n][max_new]
  int* out_count;      // [n]
  int* next_token;     // [n] token to forward at the next decode step
  const int* forced;   // [n][max_new] or null (teacher forcing for parity tests)
  int* picked;         // [n][max_new] or null: the arg-max of the processed scores at every step, whatever `forced` says
  const int* suppress; // [n_suppress]
  const int* eos;      // [n_eos]
  int ctx_cap, enc_cap, max_new, n_suppress, n_eos;
  int ngram;
  float penalty;
  int step;
};

constexpr int kSelParts = 16;       // CTAs per stream: each scans a contiguous 1/16 of the vocabulary
constexpr int kSelThreads = 256;
constexpr int kSelCtxSmem = 1024;   // prompt + generated ids staged in shared memory

struct SelectWs {
  float* best;   // [n][kSelParts]
  int* idx;      // [n][kSelParts]
  int* count;    // [n] arrival counters (left at 0 by the last CTA of every stream)
};

// grid (kSelParts, n).  A CTA applies the processors to the ids that fall into its slice of the vocabulary
// (in place: nobody else touches that slice), takes the slice arg-max, and the last CTA of a stream to arrive
// merges the kSelParts candidates and does the token bookkeeping.
__global__ void __launch_bounds__(kSelThreads)
greedy_select_kernel(float* logits, int V, GenState g, SelectWs ws) {
  pdl_launch_dependents();
  pdl_wait();
  const int part = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  if (!g.active[b]) return;
  float* lg = logits + static_cast<size_t>(b) * V;
  const int per = (V + kSelParts - 1) / kSelParts;
  const int lo = part * per, hi = min(V, lo + per);
  const int n_ctx = g.ctx_len[b];
  const int* enc = g.enc_ids + static_cast<size_t>(b) * g.enc_cap;
  const int n_enc = g.enc_len[b];
  __shared__ int ctx_s[kSelCtxSmem];
  const int* ctx_g = g.ctx_ids + static_cast<size_t>(b) * g.ctx_cap;
  const bool staged = n_ctx <= kSelCtxSmem;
  if (staged) for (int i = tid; i < n_ctx; i += kSelThreads) ctx_s[i] = ctx_g[i];
  __syncthreads();
  const int* ctx = staged ? ctx_s : ctx_g;
  // (1) repetition penalty, once per distinct id
  if (g.penalty != 1.0f) {
    for (int i = tid; i < n_ctx; i += kSelThreads) {
      const int id = ctx[i];
      if (id < lo || id >= hi) continue;
      bool first = true;
      for (int j = 0; j < i; ++j) if (ctx[j] == id) { first = false; break; }
      if (first) { const float s = lg[id]; lg[id] = s < 0.f ? s * g.penalty : s / g.penalty; }
    }
  }
  __syncthreads();
  // (2)+(3) n-gram bans: tail = last ngram-1 ids of ctx
  if (g.ngram > 0 && n_ctx + 1 >= g.ngram) {
    const int m = g.ngram - 1;
    const int* tail = ctx + n_ctx - m;
    for (int t = tid; t + g.ngram <= n_ctx; t += kSelThreads) {
      const int id = ctx[t + m];
      if (id < lo || id >= hi) continue;
      bool eq = true;
      for (int j = 0; j < m; ++j) if (ctx[t + j] != tail[j]) { eq = false; break; }
      if (eq) lg[id] = -INFINITY;
    }
    for (int t = tid; t + g.ngram <= n_enc; t += kSelThreads) {
      const int id = enc[t + m];
      if (id < lo || id >= hi) continue;
      bool eq = true;
      for (int j = 0; j < m; ++j) if (enc[t + j] != tail[j]) { eq = false; break; }
      if (eq) lg[id] = -INFINITY;
    }
  }
  // (4) suppress tokens
  for (int i = tid; i < g.n_suppress; i += kSelThreads) {
    const int id = g.suppress[i];
    if (id >= lo && id < hi) lg[id] = -INFINITY;
  }
  __syncthreads();
  // (5) slice arg-max (lowest index wins ties); four independent loads in flight per thread
  float best = -INFINITY;
  int bi = 0x7fffffff;
  int i = lo + tid;
  for (; i + 3 * kSelThreads < hi; i += 4 * kSelThreads) {
    const float s0 = lg[i], s1 = lg[i + kSelThreads], s2 = lg[i + 2 * kSelThreads], s3 = lg[i + 3 * kSelThreads];
    if (s0 > best) { best = s0; bi = i; }
    if (s1 > best) { best = s1; bi = i + kSelThreads; }
    if (s2 > best) { best = s2; bi = i + 2 * kSelThreads; }
    if (s3 > best) { best = s3; bi = i + 3 * kSelThreads; }
  }
  for (; i < hi; i += kSelThreads) {
    const float s = lg[i];
    if (s > best) { best = s; bi = i; }
  }
  __shared__ float sb[kSelThreads / 32];
  __shared__ int si[kSelThreads / 32];
  __shared__ int last_flag;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
  }
  if ((tid & 31) == 0) { sb[tid >> 5] = best; si[tid >> 5] = bi; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < kSelThreads / 32; ++w)
      if (sb[w] > best || (sb[w] == best && si[w] < bi)) { best = sb[w]; bi = si[w]; }
    ws.best[b * kSelParts + part] = best;
    ws.idx[b * kSelParts + part] = bi;
    __threadfence();
    last_flag = atomicAdd(&ws.count[b], 1) == kSelParts - 1;
  }
  __syncthreads();
  if (!last_flag || tid != 0) return;
  __threadfence();
  best = -INFINITY; bi = 0x7fffffff;
  for (int p = 0; p < kSelParts; ++p) {
    const float pb = __ldcg(ws.best + b * kSelParts + p);
    const int pi = __ldcg(ws.idx + b * kSelParts + p);
    if (pb > best || (pb == best && pi < bi)) { best = pb; bi = pi; }
  }
  ws.count[b] = 0;
  int tok = (bi == 0x7fffffff) ? 0 : bi;
  if (g.picked) g.picked[static_cast<size_t>(b) * g.max_new + g.step] = tok;
  if (g.forced) tok = g.forced[static_cast<size_t>(b) * g.max_new + g.step];
  g.out_tokens[static_cast<size_t>(b) * g.max_new + g.step] = tok;
  g.out_count[b] = g.step + 1;
  g.next_token[b] = tok;
  if (n_ctx < g.ctx_cap) g.ctx_ids[static_cast<size_t>(b) * g.ctx_cap + n_ctx] = tok;
  g.ctx_len[b] = n_ctx + 1;
  bool stop = false;
  for (int e = 0; e < g.n_eos; ++e) if (tok == g.eos[e]) stop = true;
  if (stop) g.active[b] = 0;
}

// kv_len[slot] += T[b] for active streams (after all layers appended their K/V)
__global__ void advance_kv_len_kernel(int* kv_len, const int* __restrict__ slots, const int* __restrict__ Tn,
                                      const int* __restrict__ active, int n) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < n && (!active || active[b])) kv_len[slots[b]] += Tn[b];
}

}  // namespace isst

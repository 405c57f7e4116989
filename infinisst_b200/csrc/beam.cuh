// Beam search device side: candidate selection over (beams x vocabulary) and KV page copies.
//
// Replaces, per decode step of `generation_mixin_beam_search` (model/patches/patch_hf.py:826-878):
//   log_softmax in fp32 (:830-832) -> HF logits processors on the LOG-PROBS (:834; repetition penalty,
//   no-repeat-ngram, encoder-no-repeat-ngram, suppress) -> + beam score (:835-837) -> top
//   max(2, 1 + n_eos) * num_beams over the num_beams * V candidates of every sentence (:861-878).
// The scorer itself (`beam_search_process`, :43-157) is a handful of integers per sentence and stays on the
// host; it sees 2k (score, beam, token) triples per sentence instead of the reference's [k, V] score matrix.
// `_temporary_reorder_cache` (:910-913, an index_select of the whole KV cache per step) and the per-hypothesis
// KV snapshots (:113-128) become copies of the beams' private tail pages (kv_page_copy_kernel).
#pragma once
#include "common.cuh"
#include "rowops.cuh"

namespace isst {

constexpr int kBeamMaxKeep = 32;     // candidates kept per sentence: max(2, 1 + n_eos) * num_beams
constexpr int kBeamCandCap = 512;    // slice elements that pass the threshold pre-filter (overflow: slow exact path)

struct BeamSel {
  const int* ctx_ids;        // [R][ctx_cap]  prompt of this call + tokens generated on this beam
  const int* ctx_len;        // [R]
  const int* enc_ids;        // [G][enc_cap]  last `lookback` emitted target ids of the sentence
  const int* enc_len;        // [G]
  const float* beam_score;   // [R]
  const int* suppress;       // [n_suppress]
  int n_suppress, ctx_cap, enc_cap, ngram;
  float penalty;
  int write_back;            // 1: leave the processed log-probs in `logits` (teacher-forced parity runs read them back)
  int rows_per_group;        // live beams per sentence (1 at the first step: patch_hf.py:772-774)
  int n_keep;                // candidates per sentence
  float* part_max;           // [R][kSelParts]
  float* part_sum;           // [R][kSelParts]
  float* cand_s;             // [R][kSelParts][n_keep]
  int* cand_i;               // [R][kSelParts][n_keep]  token id or -1
  int* count;                // [G] arrival counters (self re-arming)
  float* out_s;              // [G][n_keep] sorted, best first
  int* out_i;                // [G][n_keep] row_in_group * V + token
};

// (max, sum exp) of one vocabulary slice of one row.  The slice (V / 16 elements: 31.3 per thread at V = 128 256) is
// read ONCE into registers; a slice longer than kLseRegs x 256 takes the remainder in a second sweep.
constexpr int kLseRegs = 32;
__global__ void __launch_bounds__(kSelThreads)
beam_lse_kernel(const float* __restrict__ logits, int V, BeamSel s) {
  pdl_launch_dependents();
  pdl_wait();
  const int part = blockIdx.x, r = blockIdx.y, tid = threadIdx.x;
  const float* lg = logits + static_cast<size_t>(r) * V;
  const int per = (V + kSelParts - 1) / kSelParts;
  const int lo = part * per, hi = min(V, lo + per);
  float x[kLseRegs];
  float m = -INFINITY;
#pragma unroll
  for (int k = 0; k < kLseRegs; ++k) {
    const int i = lo + tid + k * kSelThreads;
    x[k] = i < hi ? __ldg(lg + i) : -INFINITY;
    m = fmaxf(m, x[k]);
  }
  for (int i = lo + tid + kLseRegs * kSelThreads; i < hi; i += kSelThreads) m = fmaxf(m, lg[i]);
  __shared__ float red[kSelThreads / 32];
  m = warp_max(m);
  if ((tid & 31) == 0) red[tid >> 5] = m;
  __syncthreads();
  m = red[0];
#pragma unroll
  for (int w = 1; w < kSelThreads / 32; ++w) m = fmaxf(m, red[w]);
  __syncthreads();
  float sum = 0.f;
  if (m > -INFINITY) {
#pragma unroll
    for (int k = 0; k < kLseRegs; ++k) sum += expf(x[k] - m);          // exp(-inf) = 0 for the padding
    for (int i = lo + tid + kLseRegs * kSelThreads; i < hi; i += kSelThreads) sum += expf(lg[i] - m);
  }
  sum = warp_sum(sum);
  if ((tid & 31) == 0) red[tid >> 5] = sum;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int w = 0; w < kSelThreads / 32; ++w) t += red[w];
    s.part_max[r * kSelParts + part] = m;
    s.part_sum[r * kSelParts + part] = t;
  }
}

// Block-wide extraction of the `n_keep` largest (value, key) pairs in (value desc, key asc) order from `count`
// items served by `get(i, &value, &key)`.  No item is marked: round t takes the best pair that comes strictly
// after the pair taken in round t-1, so duplicates of a value are all found.  Thread 0 gets each result.
template <typename Get, typename Put>
__device__ __forceinline__ void block_top_n(int count, int n_keep, Get get, Put put) {
  __shared__ float sv[kSelThreads / 32];
  __shared__ long long sk[kSelThreads / 32];
  __shared__ float last_v_s;
  __shared__ long long last_k_s;
  const int tid = threadIdx.x;
  float last_v = INFINITY;
  long long last_k = -1;
  for (int t = 0; t < n_keep; ++t) {
    float bv = -INFINITY;
    long long bk = 0x7fffffffffffffffLL;
    for (int i = tid; i < count; i += kSelThreads) {
      float v;
      long long k;
      get(i, &v, &k);
      if (k < 0 || v != v) continue;
      const bool after = v < last_v || (v == last_v && k > last_k);
      if (after && (v > bv || (v == bv && k < bk))) { bv = v; bk = k; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const long long ok = __shfl_xor_sync(0xffffffffu, bk, o);
      if (ov > bv || (ov == bv && ok < bk)) { bv = ov; bk = ok; }
    }
    if ((tid & 31) == 0) { sv[tid >> 5] = bv; sk[tid >> 5] = bk; }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < kSelThreads / 32; ++w)
        if (sv[w] > bv || (sv[w] == bv && sk[w] < bk)) { bv = sv[w]; bk = sk[w]; }
      const bool found = bk != 0x7fffffffffffffffLL;
      put(t, found ? bv : -INFINITY, found ? bk : -1LL);
      last_v_s = bv; last_k_s = found ? bk : 0x7fffffffffffffffLL;
    }
    __syncthreads();
    last_v = last_v_s; last_k = last_k_s;
    __syncthreads();
  }
}

// The same selection in ONE round for items that sit in shared memory: every item counts the items that come before it
// in (value desc, key asc) order; that count is its output position.  `put` must have been pre-filled with (-inf, -1)
// for the positions no item claims.  count^2 / 256 broadcast reads per thread instead of n_keep block-wide reductions.
template <typename Put>
__device__ __forceinline__ void block_rank_top_n(const float* v, const int* k, int count, int n_keep, Put put) {
  for (int i = threadIdx.x; i < count; i += kSelThreads) {
    const float vi = v[i];
    const int ki = k[i];
    if (ki < 0 || vi != vi) continue;
    int rank = 0;
    for (int j = 0; j < count; ++j) {
      const float vj = v[j];
      const int kj = k[j];
      rank += (kj >= 0 && vj == vj && (vj > vi || (vj == vi && kj < ki))) ? 1 : 0;
    }
    if (rank < n_keep) put(rank, vi, ki);
  }
}
constexpr int kBeamMergeCap = 1024;  // candidates of a sentence merged in shared memory (rows x 16 slices x n_keep)

// log-probs of one slice, processors, slice top-n; the last CTA of a sentence merges.  The slice is read from global
// memory once and lives in (dynamic) shared memory from then on: `lg` is that copy, indexed by token id.
__global__ void __launch_bounds__(kSelThreads)
beam_topk_kernel(float* logits, int V, BeamSel s) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float beam_slice[];
  const int part = blockIdx.x, r = blockIdx.y, tid = threadIdx.x;
  const int g = r / s.rows_per_group, rg = r - g * s.rows_per_group;
  float* lg_global = logits + static_cast<size_t>(r) * V;
  const int per = (V + kSelParts - 1) / kSelParts;
  const int lo = part * per, hi = min(V, lo + per);
  float* lg = beam_slice - lo;
  // log-sum-exp of the row from the slice partials
  float M = -INFINITY;
#pragma unroll
  for (int p = 0; p < kSelParts; ++p) M = fmaxf(M, s.part_max[r * kSelParts + p]);
  float S = 0.f;
#pragma unroll
  for (int p = 0; p < kSelParts; ++p) {
    const float pm = s.part_max[r * kSelParts + p];
    if (pm > -INFINITY) S += s.part_sum[r * kSelParts + p] * expf(pm - M);
  }
  const float lse = M + logf(S);
  for (int i = lo + tid; i < hi; i += kSelThreads) lg[i] = __ldg(lg_global + i) - lse;
  __syncthreads();
  const int n_ctx = s.ctx_len[r];
  const int* enc = s.enc_ids + static_cast<size_t>(g) * s.enc_cap;
  const int n_enc = s.enc_len[g];
  __shared__ int ctx_s[kSelCtxSmem];
  const int* ctx_g = s.ctx_ids + static_cast<size_t>(r) * s.ctx_cap;
  const bool staged = n_ctx <= kSelCtxSmem;
  if (staged) for (int i = tid; i < n_ctx; i += kSelThreads) ctx_s[i] = ctx_g[i];
  __syncthreads();
  const int* ctx = staged ? ctx_s : ctx_g;
  if (s.penalty != 1.0f) {                         // repetition penalty on the log-prob, once per distinct id
    for (int i = tid; i < n_ctx; i += kSelThreads) {
      const int id = ctx[i];
      if (id < lo || id >= hi) continue;
      bool first = true;
      for (int j = 0; j < i; ++j) if (ctx[j] == id) { first = false; break; }
      if (first) { const float x = lg[id]; lg[id] = x < 0.f ? x * s.penalty : x / s.penalty; }
    }
  }
  __syncthreads();
  if (s.ngram > 0 && n_ctx + 1 >= s.ngram) {       // n-gram bans against the beam's own ids and the emitted target ids
    const int m = s.ngram - 1;
    const int* tail = ctx + n_ctx - m;
    for (int t = tid; t + s.ngram <= n_ctx; t += kSelThreads) {
      const int id = ctx[t + m];
      if (id < lo || id >= hi) continue;
      bool eq = true;
      for (int j = 0; j < m; ++j) if (ctx[t + j] != tail[j]) { eq = false; break; }
      if (eq) lg[id] = -INFINITY;
    }
    for (int t = tid; t + s.ngram <= n_enc; t += kSelThreads) {
      const int id = enc[t + m];
      if (id < lo || id >= hi) continue;
      bool eq = true;
      for (int j = 0; j < m; ++j) if (enc[t + j] != tail[j]) { eq = false; break; }
      if (eq) lg[id] = -INFINITY;
    }
  }
  for (int i = tid; i < s.n_suppress; i += kSelThreads) {
    const int id = s.suppress[i];
    if (id >= lo && id < hi) lg[id] = -INFINITY;
  }
  __syncthreads();
  if (s.write_back)
    for (int i = lo + tid; i < hi; i += kSelThreads) lg_global[i] = lg[i];
  const float bs = s.beam_score[r];
  float* cs = s.cand_s + (static_cast<size_t>(r) * kSelParts + part) * s.n_keep;
  int* ci = s.cand_i + (static_cast<size_t>(r) * kSelParts + part) * s.n_keep;
  // Slice top-n in two levels: a cheap lower bound T of the slice's n-th largest score first, so that only elements
  // >= T (a few dozen) enter the exact selection; ties that overflow the candidate list fall back to selecting over the
  // whole slice.
  __shared__ float tmax_s[kSelThreads / 32];
  __shared__ float cand_v[kBeamMergeCap];          // slice candidates (kBeamCandCap), later the sentence's merge list
  __shared__ int cand_k[kBeamMergeCap];
  __shared__ float thr_s;
  __shared__ int cand_n;
  float tmax = -INFINITY;
  for (int i = lo + tid; i < hi; i += kSelThreads) tmax = fmaxf(tmax, lg[i] + bs);
  // Lower bound of the slice's n_keep-th largest score: every warp takes the ceil(n_keep / 8) largest of its 32 thread
  // maxima (shuffles only); these are 8 x ceil(n_keep / 8) >= n_keep distinct elements, so the smallest of them bounds
  // the n_keep-th largest from below.  A warp with fewer finite values than that makes the bound -inf (exact fallback).
  {
    const int lane = tid & 31;
    float mv = tmax;
    bool live = tmax > -INFINITY;
    float last = INFINITY;
    const int rounds = (s.n_keep + 7) >> 3;
    for (int t = 0; t < rounds; ++t) {
      float bv = live ? mv : -INFINITY;
      int bk = live ? lane : 64;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int ok = __shfl_xor_sync(0xffffffffu, bk, o);
        if (ov > bv || (ov == bv && ok < bk)) { bv = ov; bk = ok; }
      }
      last = bv;                                            // -inf when the warp has run out of finite values
      if (lane == bk) live = false;
    }
    if (lane == 0) tmax_s[tid >> 5] = last;
  }
  if (tid == 0) cand_n = 0;
  for (int t = tid; t < s.n_keep; t += kSelThreads) { cs[t] = -INFINITY; ci[t] = -1; }
  __syncthreads();
  if (tid == 0) {
    float t = tmax_s[0];
#pragma unroll
    for (int w = 1; w < kSelThreads / 32; ++w) t = fminf(t, tmax_s[w]);
    thr_s = t;
  }
  __syncthreads();
  const float thr = thr_s;
  for (int i = lo + tid; i < hi; i += kSelThreads) {
    const float v = lg[i] + bs;
    if (v >= thr && v > -INFINITY) {
      const int pos = atomicAdd(&cand_n, 1);
      if (pos < kBeamCandCap) { cand_v[pos] = v; cand_k[pos] = i; }
    }
  }
  __syncthreads();
  if (cand_n <= kBeamCandCap) {
    block_rank_top_n(cand_v, cand_k, cand_n, s.n_keep, [&](int t, float v, int k) { cs[t] = v; ci[t] = k; });
  } else {
    block_top_n(hi - lo, s.n_keep,
                [&](int i, float* v, long long* k) { *v = lg[lo + i] + bs; *k = lo + i; },
                [&](int t, float v, long long k) { cs[t] = v; ci[t] = static_cast<int>(k); });
  }
  __shared__ int last_flag;
  if (tid == 0) {
    __threadfence();
    last_flag = atomicAdd(&s.count[g], 1) == kSelParts * s.rows_per_group - 1;
  }
  __syncthreads();
  if (!last_flag) return;
  __threadfence();
  const int n_c = s.rows_per_group * kSelParts * s.n_keep;
  const float* gs = s.cand_s + static_cast<size_t>(g) * s.rows_per_group * kSelParts * s.n_keep;
  const int* gi = s.cand_i + static_cast<size_t>(g) * s.rows_per_group * kSelParts * s.n_keep;
  const int per_row = kSelParts * s.n_keep;
  if (n_c <= kBeamMergeCap) {
    for (int i = tid; i < n_c; i += kSelThreads) {
      const int tok = __ldcg(gi + i);
      cand_v[i] = __ldcg(gs + i);
      cand_k[i] = tok < 0 ? -1 : (i / per_row) * V + tok;
    }
    for (int t = tid; t < s.n_keep; t += kSelThreads) { s.out_s[g * s.n_keep + t] = -INFINITY; s.out_i[g * s.n_keep + t] = -1; }
    __syncthreads();
    block_rank_top_n(cand_v, cand_k, n_c, s.n_keep,
                     [&](int t, float v, int k) { s.out_s[g * s.n_keep + t] = v; s.out_i[g * s.n_keep + t] = k; });
    __syncthreads();
  } else {
    block_top_n(n_c, s.n_keep,
                [&](int i, float* v, long long* k) {
                  const int tok = __ldcg(gi + i);
                  *v = __ldcg(gs + i);
                  *k = tok < 0 ? -1LL : static_cast<long long>(i / per_row) * V + tok;
                },
                [&](int t, float v, long long k) {
                  s.out_s[g * s.n_keep + t] = v;
                  s.out_i[g * s.n_keep + t] = static_cast<int>(k);
                });
  }
  if (tid == 0) s.count[g] = 0;
  (void)rg;
}

// Copies whole KV pages (all layers): pairs[2 * i] -> pairs[2 * i + 1].  grid (n_pairs, layers).
__global__ void __launch_bounds__(256)
kv_page_copy_kernel(bf16* pool, size_t layer_elems, int page_elems, const int* __restrict__ pairs) {
  pdl_launch_dependents();
  pdl_wait();
  const int src = pairs[2 * blockIdx.x], dst = pairs[2 * blockIdx.x + 1];
  const uint4* s = reinterpret_cast<const uint4*>(pool + blockIdx.y * layer_elems + static_cast<size_t>(src) * page_elems);
  uint4* d = reinterpret_cast<uint4*>(pool + blockIdx.y * layer_elems + static_cast<size_t>(dst) * page_elems);
  const int n = page_elems / 8;
  for (int i = threadIdx.x; i < n; i += blockDim.x) d[i] = s[i];
}

}  // namespace isst

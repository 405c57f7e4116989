// Attention kernels over position-stable rotated KV caches.
//
// The reference keeps K UN-rotated and re-applies RoPE to every cached key at its index in the current
// cache on every call (patch_llm.py:280-299, patch_speech_encoder.py:797-824), so that sliding-window
// eviction shifts all positions consistently.  A score only depends on the DIFFERENCE of the two
// positions, so the same scores are obtained with keys rotated ONCE, when they are appended, at their
// ABSOLUTE index a (tokens / frames ever appended to the stream), and queries rotated at their absolute
// index: the difference a_q - a_k equals the reference's pos_q - pos_k for every key that slid with the
// window.  The pinned system prompt does not slide: its keys keep absolute index == reference position,
// and the reference distance to them is (a_q - evicted) - a_k, so a second copy of the query rotated at
// a_q - evicted is used for the system-prompt keys (two query variants per turn instead of re-rotating
// ~1000 keys per layer per step).  Angles are formed in fp64 and reduced mod 2 pi before sin/cos, so an
// unbounded stream (absolute indices ~1e5-1e6) keeps fp32-accurate rotations.  Eviction stays a
// page-table edit with no data movement, and the kept index sets are exactly the reference's.
//
//  * rope_table_kernel: (cos, sin) of the new tokens' / frames' absolute positions
//  * llm_rope_append_kernel / enc_rope_append_kernel: rotate q in place, append rotated K and V
//  * chunk_attention_kernel<HD, ENC>: tensor-core (mma.sync m16n8k16 bf16) flash-style kernel for
//      ENC = true : wav2vec2 block-causal sliding-window attention over the per-layer KV ring
//                   (uni_mha_forward, patch_speech_encoder.py:692-933; mask closed form SURVEY §4.4)
//      ENC = false: Llama chunk-prefill over the paged KV with GQA row packing
//                   (llama_sdpa_attention_new_forward, patch_llm.py:231-336)
//  * decode attention lives in decode_attention.cuh; decode_combine_kernel merges its split partials.
#pragma once
#include "common.cuh"

namespace isst {

constexpr int kPageTokens = 16;   // tokens per KV page

// ----------------------------------------------------------------------------------------------
// Paged KV addressing.  Pool layout per layer: [page][K|V][kv_head][kPageTokens][HD] bf16.
// Logical index t of a stream -> slot: t < sys_len ? t : t - sys_len + ring_start  (sys prompt pinned
// in the first pages; ring_start = 16 * sys_pages + head_off moves forward on eviction).
// ----------------------------------------------------------------------------------------------
struct PagedKV {
  bf16* pool;                  // this layer's pool base
  const int* page_table;       // [max_streams][pages_per_stream]
  const int* kv_len;           // [max_streams] logical length BEFORE the tokens being processed
  const int* sys_len;          // [max_streams]
  const int* ring_start;       // [max_streams]
  int pages_per_stream;
  int kv_heads;
  int head_dim;
};
__device__ __forceinline__ int kv_slot(int t, int sys_len, int ring_start) {
  return t < sys_len ? t : t - sys_len + ring_start;
}
__device__ __forceinline__ size_t kv_offset(const PagedKV& kv, const int* table, int slot, int is_v, int head) {
  const int page = table[slot / kPageTokens];
  const int in = slot % kPageTokens;
  return ((((static_cast<size_t>(page) * 2 + is_v) * kv.kv_heads + head) * kPageTokens) + in) * kv.head_dim;
}

// ----------------------------------------------------------------------------------------------
// mma.sync helpers
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_bf16_16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t& r0, uint32_t& r1, const void* smem_row) {
  uint32_t addr = static_cast<uint32_t>(__cvta_generic_to_shared(smem_row));
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}

struct EncAttnParams {
  const bf16* qkv;        // [tok, 3*H*HD]  q (pre-scaled by HD^-0.5 through the folded weights) rotated in place
  bf16* out;              // [tok, H*HD]
  bf16* k_ring;           // this layer: [stream_slot][H][cap][HD], keys rotated at their absolute frame index
  bf16* v_ring;
  const int* slots;       // [n] stream slot per batch entry
  const int* prefix;      // [n] frames encoded before this chunk (cache.n_steps), per batch entry
  int T;                  // new frames per stream
  int H;
  int cap;                // ring capacity (>= max_cache + T)
  int max_cache;          // cache.max_steps
  int blocksize;
};

struct LlmAttnParams {
  const bf16* qkv;        // [tok, (H + 2*Hkv) * HD]; q rotated in place at the absolute index (ring variant)
  const bf16* q_sys;      // [tok, H * HD]: q rotated at (absolute index - evicted), used against the pinned prefix
  bf16* out;              // [tok, H*HD]
  PagedKV kv;
  const int* slots;       // [n]
  const int* tok_base;    // [n] first packed row of this stream
  const int* T;           // [n] new tokens of this stream
  int H;                  // q heads
  float scale_log2;       // HD^-0.5 * log2(e)
  // key splits (prefill_attention_tc_kernel, few streams: one CTA per (stream, kv head) would leave the SMs idle):
  // split s of every row tile takes a contiguous range of key tiles and leaves an un-normalised fp32 partial
  // [(row * H + q head) * key_splits + s]; decode_combine_kernel merges.  1 (or 0): the kernel writes `out` itself.
  int key_splits;
  float* part_o;
  float* part_ml;
  int l2_ahead;           // K/V tiles asked into L2 ahead of the shared-memory ring (prefill_attention_tc_kernel)
  int kv_row0;            // first row of this layer in the pool-wide KV tensor map (rows of head_dim elements)
};

__device__ __forceinline__ void cpa16(void* smem_dst, const void* gsrc, int src_bytes) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cpa_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cpa_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* smem_row) {
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(smem_row));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], const void* smem_row) {
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(smem_row));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}

// ----------------------------------------------------------------------------------------------
// (cos, sin) of absolute positions, fp64 angle reduced mod 2 pi.
//   LLM (ENC = false): token i of batch entry b sits at logical index kv_len[b] + i; ring table at
//     kv_len + i + evicted (absolute), sys table at kv_len + i (reference position, see header).
//   ENC: frame i of batch entry b sits at absolute frame prefix[b] + i.
// ----------------------------------------------------------------------------------------------
template <bool ENC>
__global__ void rope_table_kernel(float2* __restrict__ tab_ring, float2* __restrict__ tab_sys,
                                  const int* __restrict__ tok_base, const int* __restrict__ Tn,
                                  const int* __restrict__ base_pos, const int* __restrict__ evicted,
                                  const int* __restrict__ active, const float* __restrict__ inv_freq, int n_freq,
                                  int T_fixed) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.y;
  if (active && !active[b]) return;
  const int T = ENC ? T_fixed : Tn[b];
  const int row0 = ENC ? b * T_fixed : tok_base[b];
  const double two_pi = 6.283185307179586476925286766559;
  for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < T * n_freq; u += gridDim.x * blockDim.x) {
    const int i = u / n_freq, d = u % n_freq;
    const double f = static_cast<double>(inv_freq[d]);
    const long long a_sys = static_cast<long long>(base_pos[b]) + i;
    const long long a_ring = a_sys + (ENC ? 0 : evicted[b]);
    double ang = static_cast<double>(a_ring) * f;
    ang -= two_pi * floor(ang / two_pi);
    float sn, cs;
    sincosf(static_cast<float>(ang), &sn, &cs);
    tab_ring[static_cast<size_t>(row0 + i) * n_freq + d] = make_float2(cs, sn);
    if (!ENC) {
      double as = static_cast<double>(a_sys) * f;
      as -= two_pi * floor(as / two_pi);
      sincosf(static_cast<float>(as), &sn, &cs);
      tab_sys[static_cast<size_t>(row0 + i) * n_freq + d] = make_float2(cs, sn);
    }
  }
}

// Shared-memory bytes of chunk_attention_kernel<HD, *, NW, NS>: NS (K, V) stages of 64 keys; the query staging area
// aliases the last stage (the query fragments live in registers while the tiles stream).
template <int HD, int NW, int NS>
constexpr int chunk_attn_smem_bytes() {
  static_assert(NW * 16 <= 2 * 64, "query staging must fit one stage");
  return NS * 2 * 64 * (HD + 8) * 2;
}

// One CTA = NW warps x 16 query rows, walking the keys in 64-key tiles.  K (already rotated) and V tiles
// stream global -> shared memory with cp.async through an NS-stage ring (NS - 1 tiles in flight while the current
// one is consumed by the tensor cores); 2+ CTAs per SM keep further pipelines running.
// LLM: the pinned system-prompt keys [0, sys_len) are visited first with the q_sys query variant, then the
// query fragments are reloaded from the ring variant for the sliding part (see the header).
template <int HD, bool ENC, int NW, int NS>
__global__ void __launch_bounds__(NW * 32, HD == 128 ? 2 : 4)
chunk_attention_kernel(const EncAttnParams ep, const LlmAttnParams lp) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int KT = 64;             // keys per tile
  constexpr int LDS = HD + 8;        // padded smem row (elements)
  constexpr int NTHREADS = NW * 32;
  extern __shared__ __align__(16) uint8_t attn_smem[];
  bf16* sRaw = reinterpret_cast<bf16*>(attn_smem);          // [NS stages][K | V][KT][LDS]
  bf16* sQ = sRaw + (NS - 1) * 2 * KT * LDS;                // [NW*16][LDS] query staging: aliases a stage that is free at that time

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int b = blockIdx.z;
  const int head = blockIdx.y;       // ENC: attention head; LLM: kv head
  const int row0 = blockIdx.x * (NW * 16);

  // ---- per-stream geometry ----
  int slot, T, L, n_rows, kept = 0, prefix = 0, ring0 = 0, tok0, group = 1, sys_len = 0, ring_start = 0;
  const int* table = nullptr;
  if (ENC) {
    slot = ep.slots[b];
    T = ep.T;
    prefix = ep.prefix[b];
    kept = min(prefix, ep.max_cache);
    L = kept + T;
    n_rows = T;
    tok0 = b * T;
    ring0 = (prefix - kept) % ep.cap;   // ring slot of window index 0: frame f lives in slot f % cap
  } else {
    slot = lp.slots[b];
    T = lp.T[b];
    tok0 = lp.tok_base[b];
    group = lp.H / lp.kv.kv_heads;
    n_rows = group * T;
    L = lp.kv.kv_len[slot] + T;
    sys_len = min(lp.kv.sys_len[slot], L);
    ring_start = lp.kv.ring_start[slot];
    table = lp.kv.page_table + static_cast<size_t>(slot) * lp.kv.pages_per_stream;
  }
  if (row0 >= n_rows) return;

  // keys needed by this CTA: up to the largest visible index over its rows
  int key_end;
  if (ENC) key_end = L;
  else {
    const int rmax = min(row0 + NW * 16, n_rows) - 1;
    // rows are hq * T + i: a tile may span several heads, so the largest i is T-1 unless the tile is inside one head
    const int i_hi = (row0 / T == rmax / T) ? (rmax % T) : (T - 1);
    key_end = L - T + i_hi + 1;
  }
  // tile list: [0, sys_end) in 64-key tiles (query variant q_sys), then [sys_end, key_end)
  const int sys_end = ENC ? 0 : min(sys_len, key_end);
  const int n_sys_tiles = (sys_end + KT - 1) / KT;
  const int n_tiles = n_sys_tiles + (key_end - sys_end + KT - 1) / KT;
  auto tile_k0 = [&](int t) { return t < n_sys_tiles ? t * KT : sys_end + (t - n_sys_tiles) * KT; };
  auto tile_k1 = [&](int t) { return t < n_sys_tiles ? min(sys_end, t * KT + KT) : min(key_end, sys_end + (t - n_sys_tiles + 1) * KT); };

  // ---- tile loader (cp.async, zero fill beyond the tile's key range) ----
  // LLM: threads 0-127 = 8 groups x 16 chunks; a group copies 8 consecutive keys, whose slots touch at most two
  // pages -> two page-table lookups per thread and tile, issued one tile ahead (as in decode_attention.cuh).
  const int grp = tid >> 4, chunk = tid & 15;
  int nx_s0 = 0, nx_pa = 0, nx_pb = 0;
  auto lookup = [&](int t) {
    if (ENC || tid >= 128) return;
    const int k1 = tile_k1(t);
    const int jg = min(tile_k0(t) + 8 * grp, L - 1);
    nx_s0 = kv_slot(jg, sys_len, ring_start);
    const int last = kv_slot(min(jg + 7, k1 - 1 > jg ? k1 - 1 : jg), sys_len, ring_start);
    nx_pa = table[nx_s0 >> 4];
    nx_pb = table[last >> 4];
  };
  auto load_tile = [&](int t, int stage) {
    constexpr int CH = HD / 8;
    bf16* dK = sRaw + stage * 2 * KT * LDS;
    bf16* dV = dK + KT * LDS;
    const int k0 = tile_k0(t), k1 = tile_k1(t);
    if (ENC) {
      for (int u = tid; u < KT * CH; u += NTHREADS) {
        const int kl = u / CH, c = u % CH;
        const int j = k0 + kl;
        const bool ok = j < k1;
        const int rs = ok ? (ring0 + j) % ep.cap : 0;
        const size_t off = ((static_cast<size_t>(slot) * ep.H + head) * ep.cap + rs) * HD + c * 8;
        cpa16(dK + kl * LDS + c * 8, ep.k_ring + off, ok ? 16 : 0);
        cpa16(dV + kl * LDS + c * 8, ep.v_ring + off, ok ? 16 : 0);
      }
    } else {
      if (tid < 128) {
        const size_t page_elems = static_cast<size_t>(2) * lp.kv.kv_heads * kPageTokens * HD;
        const bf16* head_base = lp.kv.pool + static_cast<size_t>(head) * kPageTokens * HD + chunk * 8;
        const size_t v_off = static_cast<size_t>(lp.kv.kv_heads) * kPageTokens * HD;
        const int n_ok = k1 - (k0 + 8 * grp);
        const int s0 = nx_s0;
        const bf16* base_a = head_base + static_cast<size_t>(nx_pa) * page_elems;
        const bf16* base_b = head_base + static_cast<size_t>(nx_pb) * page_elems;
        bf16* sk = dK + (8 * grp) * LDS + chunk * 8;
        bf16* sv = dV + (8 * grp) * LDS + chunk * 8;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int sl = s0 + it;
          const bf16* src = (((sl ^ s0) & ~15) == 0 ? base_a : base_b) + (sl & 15) * HD;
          const bool ok = it < n_ok;
          if (!ok) src = lp.kv.pool;
          cpa16(sk + it * LDS, src, ok ? 16 : 0);
          cpa16(sv + it * LDS, src + v_off, ok ? 16 : 0);
        }
      }
      if (t + 1 < n_tiles) lookup(t + 1);
    }
  };
  if (n_tiles > 0) lookup(0);
#pragma unroll
  for (int s0 = 0; s0 < NS - 1; ++s0) {
    if (s0 < n_tiles) load_tile(s0, s0);
    cpa_commit();
  }

  // ---- query staging: rows r = hq * T + i ----
  auto stage_q = [&](bool sys_variant, bf16* sQ) {
    constexpr int CH = HD / 8;
    for (int u = tid; u < NW * 16 * CH; u += NTHREADS) {
      const int rl = u / CH, c = u % CH;
      const int r = row0 + rl;
      uint4 val = make_uint4(0, 0, 0, 0);
      if (r < n_rows) {
        if (ENC) {
          val = *reinterpret_cast<const uint4*>(ep.qkv + static_cast<size_t>(tok0 + r) * (3 * ep.H * HD) + head * HD + c * 8);
        } else {
          const int hq = r / T, i = r % T;
          const int qh = head * group + hq;
          if (sys_variant) val = *reinterpret_cast<const uint4*>(lp.q_sys + static_cast<size_t>(tok0 + i) * (lp.H * HD) + qh * HD + c * 8);
          else val = *reinterpret_cast<const uint4*>(lp.qkv + static_cast<size_t>(tok0 + i) * ((lp.H + 2 * lp.kv.kv_heads) * HD) + qh * HD + c * 8);
        }
      }
      *reinterpret_cast<uint4*>(sQ + rl * LDS + c * 8) = val;
    }
  };
  uint32_t qf[HD / 16][4];
  auto load_qf = [&](const bf16* sQ) {
    const bf16* q0 = sQ + (warp * 16 + g) * LDS;
    const bf16* q1 = q0 + 8 * LDS;
#pragma unroll
    for (int kk = 0; kk < HD / 16; ++kk) {
      qf[kk][0] = *reinterpret_cast<const uint32_t*>(q0 + kk * 16 + 2 * t4);
      qf[kk][1] = *reinterpret_cast<const uint32_t*>(q1 + kk * 16 + 2 * t4);
      qf[kk][2] = *reinterpret_cast<const uint32_t*>(q0 + kk * 16 + 8 + 2 * t4);
      qf[kk][3] = *reinterpret_cast<const uint32_t*>(q1 + kk * 16 + 8 + 2 * t4);
    }
  };
  stage_q(n_sys_tiles > 0, sQ);
  __syncthreads();
  load_qf(sQ);

  float o_acc[HD / 8][4];
#pragma unroll
  for (int n = 0; n < HD / 8; ++n) { o_acc[n][0] = o_acc[n][1] = o_acc[n][2] = o_acc[n][3] = 0.f; }
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};

  // query geometry of this thread's two rows (g and g + 8 of the warp's 16)
  int rq[2], qlo[2], qhi[2];   // visible key window [qlo, qhi) in window/logical indices
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int r = row0 + warp * 16 + g + 8 * h;
    rq[h] = r;
    if (r >= n_rows) { qlo[h] = 0; qhi[h] = 0; continue; }
    if (ENC) {
      const int p = prefix + r;                       // absolute frame index
      const int lo_abs = max(0, p - ep.max_cache);
      const int hi_abs = min((p / ep.blocksize + 1) * ep.blocksize, prefix + T);
      qlo[h] = lo_abs - (prefix - kept);
      qhi[h] = hi_abs - (prefix - kept);
    } else {
      const int i = r % T;
      qlo[h] = 0;
      qhi[h] = L - T + i + 1;                         // causal, bottom-right aligned
    }
  }
  const float sl2 = ENC ? 1.4426950408889634f : lp.scale_log2;
  const bool warp_live = row0 + warp * 16 < n_rows;

  // rows of this warp see every key of a tile when the tile lies inside [wlo, whi): such tiles skip the mask
  int wlo = max(qlo[0], qlo[1]), whi = min(qhi[0], qhi[1]);
  // (a dead row has qhi == 0, which keeps its warp on the masked path)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    wlo = max(wlo, __shfl_xor_sync(0xffffffffu, wlo, o));
    whi = min(whi, __shfl_xor_sync(0xffffffffu, whi, o));
  }

  for (int t = 0; t < n_tiles; ++t) {
    const int k0 = tile_k0(t);
    cpa_wait<NS - 2>();
    __syncthreads();   // tile t landed; every warp is done with tile t-1, whose stage is free now
    if (!ENC && t == n_sys_tiles && n_sys_tiles > 0) {
      // switch from the system-prompt segment to the sliding segment: the ring query variant is staged through the
      // stage that was just released (the next prefetch goes there afterwards)
      bf16* sQ2 = sRaw + ((t + NS - 1) % NS) * 2 * KT * LDS;
      stage_q(false, sQ2);
      __syncthreads();
      load_qf(sQ2);
      __syncthreads();
    }
    if (t + NS - 1 < n_tiles) load_tile(t + NS - 1, (t + NS - 1) % NS);
    cpa_commit();
    if (!warp_live) continue;   // warp has no valid rows (still takes part in loading)
    const bf16* sK = sRaw + (t % NS) * 2 * KT * LDS;
    const bf16* sV = sK + KT * LDS;

    // ---- S = Q K^T ----
    float s[KT / 8][4];
#pragma unroll
    for (int n = 0; n < KT / 8; ++n) { s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f; }
    {
      // ldmatrix.x4 over a 16-key x 16-dim block: (keys 0-7, lo) (keys 0-7, hi) (keys 8-15, lo) (keys 8-15, hi)
      const bf16* krow = sK + (((lane >> 4) << 3) + (lane & 7)) * LDS + (((lane >> 3) & 1) << 3);
#pragma unroll
      for (int n2 = 0; n2 < KT / 16; ++n2) {
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
          uint32_t kf[4];
          ldsm_x4(kf, krow + n2 * 16 * LDS + kk * 16);
          mma_bf16_16816(s[2 * n2], qf[kk], kf[0], kf[1]);
          mma_bf16_16816(s[2 * n2 + 1], qf[kk], kf[2], kf[3]);
        }
      }
    }
    // ---- mask + online softmax (rows g, g+8; cols n*8 + 2*t4 + {0,1}) ----
    const int k1 = tile_k1(t);
    float mx[2] = {m_run[0], m_run[1]};
    if (k0 >= wlo && k0 + KT <= whi && k0 + KT <= k1) {      // interior tile: every key visible to every row of the warp
#pragma unroll
      for (int n = 0; n < KT / 8; ++n) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          s[n][e] *= sl2;
          mx[e >> 1] = fmaxf(mx[e >> 1], s[n][e]);
        }
      }
    } else {
#pragma unroll
      for (int n = 0; n < KT / 8; ++n) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int h = e >> 1;
          const int j = k0 + n * 8 + 2 * t4 + (e & 1);
          const bool ok = (j >= qlo[h]) && (j < qhi[h]) && (j < k1);
          s[n][e] = ok ? s[n][e] * sl2 : -INFINITY;
          mx[h] = fmaxf(mx[h], s[n][e]);
        }
      }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
      mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
    }
    float corr[2], msafe[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      msafe[h] = (mx[h] == -INFINITY) ? 0.f : mx[h];
      corr[h] = (m_run[h] == -INFINITY) ? 0.f : exp2_fast(m_run[h] - msafe[h]);
      m_run[h] = mx[h];
      l_run[h] *= corr[h];
    }
#pragma unroll
    for (int n = 0; n < HD / 8; ++n) {
      o_acc[n][0] *= corr[0]; o_acc[n][1] *= corr[0];
      o_acc[n][2] *= corr[1]; o_acc[n][3] *= corr[1];
    }
    uint32_t pf[KT / 16][4];
    float ls[2] = {0.f, 0.f};
#pragma unroll
    for (int n = 0; n < KT / 8; ++n) {
      const float p0 = exp2_fast(s[n][0] - msafe[0]), p1 = exp2_fast(s[n][1] - msafe[0]);
      const float p2 = exp2_fast(s[n][2] - msafe[1]), p3 = exp2_fast(s[n][3] - msafe[1]);
      ls[0] += p0 + p1;
      ls[1] += p2 + p3;
      // C fragments of two adjacent n8 tiles form the A fragment of one k16 step
      pf[n >> 1][(n & 1) * 2 + 0] = pack_bf16(p0, p1);
      pf[n >> 1][(n & 1) * 2 + 1] = pack_bf16(p2, p3);
    }
    l_run[0] += ls[0];
    l_run[1] += ls[1];
    // ---- O += P V : ldmatrix.x4.trans over a 16-key x 16-dim block of V ----
    {
      // matrices (keys 0-7, dims lo) (keys 8-15, dims lo) (keys 0-7, dims hi) (keys 8-15, dims hi)
      // = B fragments (b0, b1) of the dim-lo n-tile and (b0, b1) of the dim-hi n-tile
      const bf16* vrow = sV + ((((lane >> 3) & 1) << 3) + (lane & 7)) * LDS + ((lane >> 4) << 3);
#pragma unroll
      for (int kk = 0; kk < KT / 16; ++kk) {
#pragma unroll
        for (int n2 = 0; n2 < HD / 16; ++n2) {
          uint32_t vf[4];
          ldsm_x4_trans(vf, vrow + kk * 16 * LDS + n2 * 16);
          mma_bf16_16816(o_acc[2 * n2], pf[kk], vf[0], vf[1]);
          mma_bf16_16816(o_acc[2 * n2 + 1], pf[kk], vf[2], vf[3]);
        }
      }
    }
  }
  cpa_wait<0>();

  // ---- normalise + store ----
  if (!warp_live) return;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    float l = l_run[h];
    l += __shfl_xor_sync(0xffffffffu, l, 1);
    l += __shfl_xor_sync(0xffffffffu, l, 2);
    const int r = rq[h];
    if (r >= n_rows) continue;
    const float inv = l > 0.f ? 1.f / l : 0.f;
    bf16* dst;
    if (ENC) dst = ep.out + static_cast<size_t>(tok0 + r) * (ep.H * HD) + head * HD;
    else {
      const int hq = r / T, i = r % T;
      dst = lp.out + static_cast<size_t>(tok0 + i) * (lp.H * HD) + (head * group + hq) * HD;
    }
#pragma unroll
    for (int n = 0; n < HD / 8; ++n) {
      *reinterpret_cast<uint32_t*>(dst + n * 8 + 2 * t4) =
          pack_bf16(o_acc[n][2 * h] * inv, o_acc[n][2 * h + 1] * inv);
    }
  }
}

// ----------------------------------------------------------------------------------------------
// Encoder: rotate q in place and append rotated K / plain V of the T new frames to the ring slots
// (prefix + i) % cap.  Interleaved-pair RoPE (rotary_embedding_torch, patch_speech_encoder.py:823-824) at
// the absolute frame index; `tab` = rope_table_kernel<true> output [n*T][HD/2] (cos, sin).
// ----------------------------------------------------------------------------------------------
// xPos (`xpos_base` != nullptr, --xpos 1): rotary_embedding_torch scales the rotated query at window index
// kept + i by base_d^((kept + i - T/2) / 512) - `get_scale(seq[-q_len:])` is centred on q_len // 2 - with the
// scale cast to the model dtype first; keys are stored rotated but un-scaled, their window-relative scale is
// applied per call by enc_xpos_keys_kernel.
__global__ void enc_rope_append_kernel(bf16* __restrict__ qkv, bf16* k_ring, bf16* v_ring,
                                       const int* __restrict__ slots, const int* __restrict__ prefix,
                                       const float2* __restrict__ tab, int T, int H, int HD, int cap,
                                       const float* __restrict__ xpos_base, int max_cache, float xpos_inv_scale_base) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.y;
  const int slot = slots[b];
  const int pre = prefix[b];   // per batch entry
  const int kept = min(pre, max_cache);
  const int chunks = H * HD / 8;
  for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < T * chunks; u += gridDim.x * blockDim.x) {
    const int i = u / chunks, c = u % chunks;
    const int head = (c * 8) / HD, d = (c * 8) % HD;
    bf16* row = qkv + static_cast<size_t>(b * T + i) * (3 * H * HD);
    const float2* cs = tab + static_cast<size_t>(b * T + i) * (HD / 2) + d / 2;
    const float2 c0 = cs[0], c1 = cs[1], c2 = cs[2], c3 = cs[3];
    const float cc[4] = {c0.x, c1.x, c2.x, c3.x}, ss[4] = {c0.y, c1.y, c2.y, c3.y};
    float qs[4] = {1.f, 1.f, 1.f, 1.f};
    if (xpos_base) {
      const float power = static_cast<float>(kept + i - T / 2) * xpos_inv_scale_base;
#pragma unroll
      for (int j = 0; j < 4; ++j) qs[j] = bf16_round(powf(xpos_base[d / 2 + j], power));
    }
    uint4 q = *reinterpret_cast<const uint4*>(row + c * 8);
    uint4 k = *reinterpret_cast<const uint4*>(row + H * HD + c * 8);
    uint32_t qw[4] = {q.x, q.y, q.z, q.w}, kw[4] = {k.x, k.y, k.z, k.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 x = unpack_bf16(qw[j]);
      qw[j] = pack_bf16((x.x * cc[j] - x.y * ss[j]) * qs[j], (x.y * cc[j] + x.x * ss[j]) * qs[j]);
      x = unpack_bf16(kw[j]);
      kw[j] = pack_bf16(x.x * cc[j] - x.y * ss[j], x.y * cc[j] + x.x * ss[j]);
    }
    *reinterpret_cast<uint4*>(row + c * 8) = make_uint4(qw[0], qw[1], qw[2], qw[3]);
    const size_t dst = ((static_cast<size_t>(slot) * H + head) * cap + (pre + i) % cap) * HD + d;
    *reinterpret_cast<uint4*>(k_ring + dst) = make_uint4(kw[0], kw[1], kw[2], kw[3]);
    *reinterpret_cast<uint4*>(v_ring + dst) = *reinterpret_cast<const uint4*>(row + 2 * H * HD + c * 8);
  }
}

// xPos key scaling (--xpos 1; not on the production path).  The reference re-scales every cached key on every call
// by get_scale(seq) ** -1 with seq = [0, L) the indices in the CURRENT window, centred on L // 2
// (patch_speech_encoder.py:823-824 -> rotate_queries_with_cached_keys): the factor of a key changes as the window
// slides, so it cannot be baked into the ring.  This pass writes the window's keys, scaled, into a scratch ring of
// the same geometry (same slot index), which the attention kernel then reads instead of k_ring.
//   scale = bf16(base_d ^ ((j - L/2) / 512));  k' = bf16(k_rot * bf16(1 / scale))
__global__ void enc_xpos_keys_kernel(const bf16* __restrict__ k_ring, bf16* __restrict__ k_scaled,
                                     const int* __restrict__ slots, const int* __restrict__ prefix, int T, int H, int HD,
                                     int cap, int max_cache, const float* __restrict__ xpos_base, float inv_scale_base) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.y;
  const int slot = slots[b];
  const int pre = prefix[b];
  const int kept = min(pre, max_cache);
  const int L = kept + T;
  const int ring0 = (pre - kept) % cap;
  const int cpr = HD / 8;                       // 16-byte chunks per key row
  for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < H * L * cpr; u += gridDim.x * blockDim.x) {
    const int c = u % cpr, j = (u / cpr) % L, head = u / (cpr * L);
    const float power = static_cast<float>(j - L / 2) * inv_scale_base;
    const size_t off = ((static_cast<size_t>(slot) * H + head) * cap + (ring0 + j) % cap) * HD + c * 8;
    const uint4 k = *reinterpret_cast<const uint4*>(k_ring + off);
    uint32_t kw[4] = {k.x, k.y, k.z, k.w};
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const float sc = bf16_round(1.f / bf16_round(powf(xpos_base[c * 4 + p], power)));
      const float2 x = unpack_bf16(kw[p]);
      kw[p] = pack_bf16(x.x * sc, x.y * sc);
    }
    *reinterpret_cast<uint4*>(k_scaled + off) = make_uint4(kw[0], kw[1], kw[2], kw[3]);
  }
}

// --rope 0 (not on the production path): x[b, i, :] += sinusoidal_positional_embedding(n_steps, T, D)[i]
// (patch_speech_encoder.py:448-461, 488-493).  The reference forms the table entirely in bfloat16: frequencies
// exp(bf16(arange(D/2)) * -(ln 1e4 / (D/2 - 1))), positions bf16(arange(n_steps, n_steps + T)) (exact only up to
// 256), their product, sin | cos - each step rounded to bf16; the sum with the bf16 activations rounds once more.
__global__ void enc_sinusoid_add_kernel(bf16* __restrict__ x, const int* __restrict__ prefix, int T, int D) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.y;
  const int half = D / 2;
  const float step = -(logf(10000.f) / static_cast<float>(half - 1));
  for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < T * D; u += gridDim.x * blockDim.x) {
    const int i = u / D, d = u % D;
    const int f = d < half ? d : d - half;
    const float freq = bf16_round(expf(bf16_round(bf16_round(static_cast<float>(f)) * step)));
    const float pos = bf16_round(static_cast<float>(prefix[b] + i));
    const float ang = bf16_round(pos * freq);
    const float e = bf16_round(d < half ? sinf(ang) : cosf(ang));
    bf16* px = x + (static_cast<size_t>(b) * T + i) * D + d;
    *px = __float2bfloat16_rn(__bfloat162float(*px) + e);
  }
}

// ----------------------------------------------------------------------------------------------
// LLM: rotate the q heads in place (ring variant), write the q_sys variant, and append rotated K and
// plain V of the packed new tokens to the paged cache.  Half-split RoPE (HF apply_rotary_pos_emb,
// patch_llm.py:294-299).  active[b] == 0 -> stream finished (EOS): nothing is appended.
// One unit = 8 elements d..d+7 of the first half of a head together with d+64..d+71.
// ----------------------------------------------------------------------------------------------
// `part` (optional): fp32 split-K partial sums [n_part][rows][ldq] of the QKV projection; the row is then
// bf16(sum_s part[s]) instead of the contents of `qkv` (deferred split reduction, see rowops.cuh).
__global__ void llm_rope_append_kernel(bf16* __restrict__ qkv, bf16* __restrict__ q_sys, PagedKV kv,
                                       const int* __restrict__ slots, const int* __restrict__ tok_base,
                                       const int* __restrict__ Tn, const int* __restrict__ active,
                                       const float2* __restrict__ tab_ring, const float2* __restrict__ tab_sys, int H,
                                       const float* __restrict__ part, int n_part, long long part_stride) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.y;
  if (active && !active[b]) return;
  const int slot = slots[b];
  const int T = Tn[b];
  const int HD = kv.head_dim, HALF = HD / 2;
  const int upr = HALF / 8;                              // units per head
  const int heads_all = H + 2 * kv.kv_heads;
  const int base = kv.kv_len[slot];
  const int sys_len = kv.sys_len[slot], ring_start = kv.ring_start[slot];
  const int* table = kv.page_table + static_cast<size_t>(slot) * kv.pages_per_stream;
  const int ldq = heads_all * HD;
  for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < T * heads_all * upr; u += gridDim.x * blockDim.x) {
    const int i = u / (heads_all * upr);
    const int hh = (u / upr) % heads_all;                // 0..H-1 q, H..H+Hkv-1 k, then v
    const int d = (u % upr) * 8;
    const int rowi = tok_base[b] + i;
    bf16* src = qkv + static_cast<size_t>(rowi) * ldq + hh * HD;
    uint4 lo, hi;
    if (part) {
      float al[8], ah[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { al[j] = 0.f; ah[j] = 0.f; }
      const float* p0 = part + static_cast<size_t>(rowi) * ldq + hh * HD + d;
      for (int sp0 = 0; sp0 < n_part; sp0 += 4) {          // four partials' loads in flight at once
        float4 a0[4], a1[4], b0[4], b1[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (sp0 + q < n_part) {
            const float4* pl = reinterpret_cast<const float4*>(p0 + (sp0 + q) * part_stride);
            const float4* ph = reinterpret_cast<const float4*>(p0 + (sp0 + q) * part_stride + HALF);
            a0[q] = __ldcg(pl); a1[q] = __ldcg(pl + 1); b0[q] = __ldcg(ph); b1[q] = __ldcg(ph + 1);
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (sp0 + q < n_part) {
            al[0] += a0[q].x; al[1] += a0[q].y; al[2] += a0[q].z; al[3] += a0[q].w;
            al[4] += a1[q].x; al[5] += a1[q].y; al[6] += a1[q].z; al[7] += a1[q].w;
            ah[0] += b0[q].x; ah[1] += b0[q].y; ah[2] += b0[q].z; ah[3] += b0[q].w;
            ah[4] += b1[q].x; ah[5] += b1[q].y; ah[6] += b1[q].z; ah[7] += b1[q].w;
          }
        }
      }
      lo = make_uint4(pack_bf16(al[0], al[1]), pack_bf16(al[2], al[3]), pack_bf16(al[4], al[5]), pack_bf16(al[6], al[7]));
      hi = make_uint4(pack_bf16(ah[0], ah[1]), pack_bf16(ah[2], ah[3]), pack_bf16(ah[4], ah[5]), pack_bf16(ah[6], ah[7]));
    } else {
      lo = *reinterpret_cast<const uint4*>(src + d);
      hi = *reinterpret_cast<const uint4*>(src + d + HALF);
    }
    const int sl = kv_slot(base + i, sys_len, ring_start);
    if (hh >= H + kv.kv_heads) {                         // V: plain copy
      bf16* dst = kv.pool + kv_offset(kv, table, sl, 1, hh - H - kv.kv_heads);
      *reinterpret_cast<uint4*>(dst + d) = lo;
      *reinterpret_cast<uint4*>(dst + d + HALF) = hi;
      continue;
    }
    const uint32_t wl[4] = {lo.x, lo.y, lo.z, lo.w}, wh[4] = {hi.x, hi.y, hi.z, hi.w};
    // keys of the pinned prefix keep absolute index == reference position (sys table); everything else
    // (ring keys, ring-variant queries) sits at logical index + evicted (ring table)
    const bool sys_key = hh >= H && base + i < sys_len;
    const float2* cr = (sys_key ? tab_sys : tab_ring) + static_cast<size_t>(rowi) * HALF + d;
    uint32_t ol[4], oh[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 a = unpack_bf16(wl[j]), bb = unpack_bf16(wh[j]);
      const float2 c0 = cr[2 * j], c1 = cr[2 * j + 1];
      ol[j] = pack_bf16(a.x * c0.x - bb.x * c0.y, a.y * c1.x - bb.y * c1.y);
      oh[j] = pack_bf16(bb.x * c0.x + a.x * c0.y, bb.y * c1.x + a.y * c1.y);
    }
    if (hh < H) {                                        // q: in place (ring variant) + sys variant
      *reinterpret_cast<uint4*>(src + d) = make_uint4(ol[0], ol[1], ol[2], ol[3]);
      *reinterpret_cast<uint4*>(src + d + HALF) = make_uint4(oh[0], oh[1], oh[2], oh[3]);
      const float2* cs = tab_sys + static_cast<size_t>(rowi) * HALF + d;
      uint32_t sl_[4], sh_[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 a = unpack_bf16(wl[j]), bb = unpack_bf16(wh[j]);
        const float2 c0 = cs[2 * j], c1 = cs[2 * j + 1];
        sl_[j] = pack_bf16(a.x * c0.x - bb.x * c0.y, a.y * c1.x - bb.y * c1.y);
        sh_[j] = pack_bf16(bb.x * c0.x + a.x * c0.y, bb.y * c1.x + a.y * c1.y);
      }
      bf16* qs = q_sys + static_cast<size_t>(rowi) * (H * HD) + hh * HD;
      *reinterpret_cast<uint4*>(qs + d) = make_uint4(sl_[0], sl_[1], sl_[2], sl_[3]);
      *reinterpret_cast<uint4*>(qs + d + HALF) = make_uint4(sh_[0], sh_[1], sh_[2], sh_[3]);
    } else {                                             // k: rotated into the page
      bf16* dst = kv.pool + kv_offset(kv, table, sl, 0, hh - H);
      *reinterpret_cast<uint4*>(dst + d) = make_uint4(ol[0], ol[1], ol[2], ol[3]);
      *reinterpret_cast<uint4*>(dst + d + HALF) = make_uint4(oh[0], oh[1], oh[2], oh[3]);
    }
  }
}

// out[b, h, :] = sum_s w_s o_s / sum_s w_s l_s  over the key splits of decode attention
__global__ void decode_combine_kernel(const float* __restrict__ part_o, const float* __restrict__ part_ml,
                                      bf16* __restrict__ out, int H, int HD, int splits) {
  pdl_launch_dependents();
  pdl_wait();
  const int bh = blockIdx.x;   // b * H + h
  const int d = threadIdx.x;
  float mm = -INFINITY;
  for (int s = 0; s < splits; ++s) mm = fmaxf(mm, part_ml[(static_cast<size_t>(bh) * splits + s) * 2]);
  const float ms = (mm == -INFINITY) ? 0.f : mm;
  float ll = 0.f, oo = 0.f;
  for (int s = 0; s < splits; ++s) {
    const size_t pi = static_cast<size_t>(bh) * splits + s;
    const float pm = part_ml[pi * 2];
    const float w = (pm == -INFINITY) ? 0.f : exp2_fast(pm - ms);
    ll += w * part_ml[pi * 2 + 1];
    if (d < HD) oo += w * part_o[pi * HD + d];
  }
  if (d < HD) out[static_cast<size_t>(bh) * HD + d] = __float2bfloat16_rn(ll > 0.f ? oo / ll : 0.f);
}

}  // namespace isst

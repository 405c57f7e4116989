// Attention kernels with RoPE-on-read over un-rotated KV caches.
//
//  * chunk_attention_kernel<HD, ENC>: tensor-core (mma.sync m16n8k16 bf16) flash-style kernel for
//      ENC = true : wav2vec2 block-causal sliding-window attention over the per-layer KV ring
//                   (uni_mha_forward, patch_speech_encoder.py:692-933; mask closed form SURVEY §4.4;
//                   interleaved-pair RoPE, rotate_queries_with_cached_keys :823-824)
//      ENC = false: Llama chunk-prefill over the paged KV with GQA row packing
//                   (llama_sdpa_attention_new_forward, patch_llm.py:231-336; half-split RoPE at
//                   positions 0..L-1 of the *current* cache, :287-299)
//  * decode_attention_kernel / decode_combine_kernel: single-token split-K decode over the paged KV
//    (HBM-bound).
//
// Keys are stored UN-rotated (patch_llm.py:280-284, patch_speech_encoder.py:797-821) and rotated while
// they are staged into shared memory, with window-relative positions, exactly like the reference
// re-rotates the whole cache on every call -- so sliding-window eviction needs no data movement.
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
  const bf16* qkv;        // [tok, 3*H*HD]  (q pre-scaled by HD^-0.5 through the folded weights)
  bf16* out;              // [tok, H*HD]
  bf16* k_ring;           // this layer: [stream_slot][H][cap][HD]
  bf16* v_ring;
  const int* slots;       // [n] stream slot per batch entry
  const int* prefix;      // [n] frames encoded before this chunk (cache.n_steps), per batch entry
  const float* rope_cos;  // [n_pos][HD/2]
  const float* rope_sin;
  int T;                  // new frames per stream
  int H;
  int cap;                // ring capacity (>= max_cache + T)
  int max_cache;          // cache.max_steps
  int blocksize;
};

struct LlmAttnParams {
  const bf16* qkv;        // [tok, (H + 2*Hkv) * HD]
  bf16* out;              // [tok, H*HD]
  PagedKV kv;
  const int* slots;       // [n]
  const int* tok_base;    // [n] first packed row of this stream
  const int* T;           // [n] new tokens of this stream
  const bf162* rope;      // [max_pos][HD/2] (cos, sin) bf16-rounded like HF (cos/sin cast to the model dtype)
  int H;                  // q heads
  float scale_log2;       // HD^-0.5 * log2(e)
};

__device__ __forceinline__ void cpa16(void* smem_dst, const void* gsrc, int src_bytes) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cpa_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cpa_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Shared-memory bytes of chunk_attention_kernel<HD, *, NW>: [Q staging | rotated K] (aliased) + 2 raw (K, V) stages.
template <int HD, int NW>
constexpr int chunk_attn_smem_bytes() {
  return ((NW * 16 > 64 ? NW * 16 : 64) + 4 * 64) * (HD + 8) * 2;
}

// One CTA = NW warps x 16 query rows, walking the keys in 64-key tiles.  Raw (un-rotated) K and V tiles
// stream global -> shared memory with cp.async through two stages, so the next tile is in flight while the
// current one is rotated (cooperative pass: raw K -> rotated K, window-relative positions) and consumed by
// the tensor cores.  2 CTAs per SM (HD 128) keep a second pipeline running on the same SM.
template <int HD, bool ENC, int NW>
__global__ void __launch_bounds__(NW * 32)
chunk_attention_kernel(const EncAttnParams ep, const LlmAttnParams lp) {
  constexpr int KT = 64;             // keys per tile
  constexpr int LDS = HD + 8;        // padded smem row (elements)
  constexpr int NTHREADS = NW * 32;
  constexpr int QROWS = NW * 16 > KT ? NW * 16 : KT;
  extern __shared__ __align__(16) uint8_t attn_smem[];
  bf16* sQ = reinterpret_cast<bf16*>(attn_smem);            // [NW*16][LDS] rotated queries (start only)
  bf16* sK = sQ;                                            // [KT][LDS] rotated keys of the current tile (aliases sQ)
  bf16* sRaw = sQ + QROWS * LDS;                            // [2 stages][K | V][KT][LDS] raw tiles

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
    // ring slot of window index 0: frame f lives in slot f % cap
    ring0 = (prefix - kept) % ep.cap;
  } else {
    slot = lp.slots[b];
    T = lp.T[b];
    tok0 = lp.tok_base[b];
    group = lp.H / lp.kv.kv_heads;
    n_rows = group * T;
    L = lp.kv.kv_len[slot] + T;
    sys_len = lp.kv.sys_len[slot];
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
  const int n_tiles = (key_end + KT - 1) / KT;

  // ---- raw tile loader (cp.async, zero fill beyond L) ----
  auto load_tile = [&](int t, int stage) {
    constexpr int CH = HD / 8;
    bf16* dK = sRaw + stage * 2 * KT * LDS;
    bf16* dV = dK + KT * LDS;
    for (int u = tid; u < KT * CH; u += NTHREADS) {
      const int kl = u / CH, c = u % CH;
      const int j = t * KT + kl;
      const bool ok = j < L;
      const bf16* ks;
      const bf16* vs;
      if (ENC) {
        const int rs = ok ? (ring0 + j) % ep.cap : 0;
        const size_t off = ((static_cast<size_t>(slot) * ep.H + head) * ep.cap + rs) * HD + c * 8;
        ks = ep.k_ring + off;
        vs = ep.v_ring + off;
      } else {
        const int sl = ok ? kv_slot(j, sys_len, ring_start) : 0;
        ks = lp.kv.pool + (ok ? kv_offset(lp.kv, table, sl, 0, head) : 0) + c * 8;
        vs = ks + static_cast<size_t>(lp.kv.kv_heads) * kPageTokens * HD;
      }
      cpa16(dK + kl * LDS + c * 8, ks, ok ? 16 : 0);
      cpa16(dV + kl * LDS + c * 8, vs, ok ? 16 : 0);
    }
  };
  load_tile(0, 0);
  cpa_commit();

  // ---- stage rotated Q: rows r = hq * T + i ----
  {
    constexpr int CH = HD / 8;                 // 16-byte chunks per row
    constexpr int UNITS = ENC ? CH : CH / 2;   // LLM handles chunk c together with c + CH/2
    for (int u = tid; u < NW * 16 * UNITS; u += NTHREADS) {
      const int rl = u / UNITS, c = u % UNITS;
      const int r = row0 + rl;
      bf16* dst = sQ + rl * LDS;
      if (r >= n_rows) {
        *reinterpret_cast<uint4*>(dst + c * 8) = make_uint4(0, 0, 0, 0);
        if (!ENC) *reinterpret_cast<uint4*>(dst + (c + CH / 2) * 8) = make_uint4(0, 0, 0, 0);
        continue;
      }
      if (ENC) {
        const int i = r;
        const bf16* src = ep.qkv + static_cast<size_t>(tok0 + i) * (3 * ep.H * HD) + head * HD + c * 8;
        uint4 raw = *reinterpret_cast<const uint4*>(src);
        const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
        const int pos = kept + i;
        const float* cs = ep.rope_cos + static_cast<size_t>(pos) * (HD / 2) + c * 4;
        const float* sn = ep.rope_sin + static_cast<size_t>(pos) * (HD / 2) + c * 4;
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float2 x = unpack_bf16(w[j]);
          o[j] = pack_bf16(x.x * cs[j] - x.y * sn[j], x.y * cs[j] + x.x * sn[j]);
        }
        *reinterpret_cast<uint4*>(dst + c * 8) = make_uint4(o[0], o[1], o[2], o[3]);
      } else {
        const int hq = r / T, i = r % T;
        const int qh = head * group + hq;
        const int ldq = (lp.H + 2 * lp.kv.kv_heads) * HD;
        const bf16* src = lp.qkv + static_cast<size_t>(tok0 + i) * ldq + qh * HD;
        uint4 lo = *reinterpret_cast<const uint4*>(src + c * 8);
        uint4 hi = *reinterpret_cast<const uint4*>(src + (c + CH / 2) * 8);
        const uint32_t wl[4] = {lo.x, lo.y, lo.z, lo.w}, wh[4] = {hi.x, hi.y, hi.z, hi.w};
        const int pos = L - T + i;
        const bf162* rp = lp.rope + static_cast<size_t>(pos) * (HD / 2) + c * 8;
        uint32_t ol[4], oh[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float2 a = unpack_bf16(wl[j]), bb = unpack_bf16(wh[j]);
          float2 cs0 = __bfloat1622float2(rp[2 * j]), cs1 = __bfloat1622float2(rp[2 * j + 1]);
          ol[j] = pack_bf16(a.x * cs0.x - bb.x * cs0.y, a.y * cs1.x - bb.y * cs1.y);
          oh[j] = pack_bf16(bb.x * cs0.x + a.x * cs0.y, bb.y * cs1.x + a.y * cs1.y);
        }
        *reinterpret_cast<uint4*>(dst + c * 8) = make_uint4(ol[0], ol[1], ol[2], ol[3]);
        *reinterpret_cast<uint4*>(dst + (c + CH / 2) * 8) = make_uint4(oh[0], oh[1], oh[2], oh[3]);
      }
    }
  }
  __syncthreads();

  // ---- Q fragments (A operand) ----
  uint32_t qf[HD / 16][4];
  {
    const bf16* q0 = sQ + (warp * 16 + g) * LDS;
    const bf16* q1 = q0 + 8 * LDS;
#pragma unroll
    for (int kk = 0; kk < HD / 16; ++kk) {
      qf[kk][0] = *reinterpret_cast<const uint32_t*>(q0 + kk * 16 + 2 * t4);
      qf[kk][1] = *reinterpret_cast<const uint32_t*>(q1 + kk * 16 + 2 * t4);
      qf[kk][2] = *reinterpret_cast<const uint32_t*>(q0 + kk * 16 + 8 + 2 * t4);
      qf[kk][3] = *reinterpret_cast<const uint32_t*>(q1 + kk * 16 + 8 + 2 * t4);
    }
  }
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

  for (int t = 0; t < n_tiles; ++t) {
    const int k0 = t * KT;
    cpa_wait<0>();
    __syncthreads();   // raw tile t landed; every warp is done with rotated tile t-1 (and with the Q staging)
    if (t + 1 < n_tiles) load_tile(t + 1, (t + 1) & 1);
    cpa_commit();
    const bf16* rK = sRaw + (t & 1) * 2 * KT * LDS;
    const bf16* sV = rK + KT * LDS;
    // ---- rotate raw K -> sK (window-relative positions: key j of the window sits at position j) ----
    {
      constexpr int CH = HD / 8;
      constexpr int UNITS = ENC ? CH : CH / 2;
      for (int u = tid; u < KT * UNITS; u += NTHREADS) {
        const int kl = u / UNITS, c = u % UNITS;
        const int j = min(k0 + kl, L - 1);            // rows beyond L are zero: any valid table row will do
        const bf16* src = rK + kl * LDS;
        bf16* dst = sK + kl * LDS;
        if (ENC) {
          uint4 raw = *reinterpret_cast<const uint4*>(src + c * 8);
          const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
          const float4 cs = *reinterpret_cast<const float4*>(ep.rope_cos + static_cast<size_t>(j) * (HD / 2) + c * 4);
          const float4 sn = *reinterpret_cast<const float4*>(ep.rope_sin + static_cast<size_t>(j) * (HD / 2) + c * 4);
          const float cc[4] = {cs.x, cs.y, cs.z, cs.w}, ss[4] = {sn.x, sn.y, sn.z, sn.w};
          uint32_t o[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float2 x = unpack_bf16(w[q]);
            o[q] = pack_bf16(x.x * cc[q] - x.y * ss[q], x.y * cc[q] + x.x * ss[q]);
          }
          *reinterpret_cast<uint4*>(dst + c * 8) = make_uint4(o[0], o[1], o[2], o[3]);
        } else {
          uint4 lo = *reinterpret_cast<const uint4*>(src + c * 8);
          uint4 hi = *reinterpret_cast<const uint4*>(src + (c + CH / 2) * 8);
          const uint32_t wl[4] = {lo.x, lo.y, lo.z, lo.w}, wh[4] = {hi.x, hi.y, hi.z, hi.w};
          const uint4* rp4 = reinterpret_cast<const uint4*>(lp.rope + static_cast<size_t>(j) * (HD / 2) + c * 8);
          const uint4 r0 = __ldg(rp4), r1 = __ldg(rp4 + 1);
          const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
          uint32_t ol[4], oh[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float2 a = unpack_bf16(wl[q]), bb = unpack_bf16(wh[q]);
            float2 cs0 = unpack_bf16(rr[2 * q]), cs1 = unpack_bf16(rr[2 * q + 1]);
            ol[q] = pack_bf16(a.x * cs0.x - bb.x * cs0.y, a.y * cs1.x - bb.y * cs1.y);
            oh[q] = pack_bf16(bb.x * cs0.x + a.x * cs0.y, bb.y * cs1.x + a.y * cs1.y);
          }
          *reinterpret_cast<uint4*>(dst + c * 8) = make_uint4(ol[0], ol[1], ol[2], ol[3]);
          *reinterpret_cast<uint4*>(dst + (c + CH / 2) * 8) = make_uint4(oh[0], oh[1], oh[2], oh[3]);
        }
      }
    }
    __syncthreads();
    if (!warp_live) continue;   // warp has no valid rows (still took part in staging)

    // ---- S = Q K^T ----
    float s[KT / 8][4];
#pragma unroll
    for (int n = 0; n < KT / 8; ++n) {
      s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
      const bf16* kr = sK + (n * 8 + g) * LDS + 2 * t4;
#pragma unroll
      for (int kk = 0; kk < HD / 16; ++kk) {
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(kr + kk * 16);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(kr + kk * 16 + 8);
        mma_bf16_16816(s[n], qf[kk], b0, b1);
      }
    }
    // ---- mask + online softmax (rows g, g+8; cols n*8 + 2*t4 + {0,1}) ----
    float mx[2] = {m_run[0], m_run[1]};
#pragma unroll
    for (int n = 0; n < KT / 8; ++n) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int h = e >> 1;
        const int j = k0 + n * 8 + 2 * t4 + (e & 1);
        const bool ok = (j >= qlo[h]) && (j < qhi[h]);
        // the reference forms scores in the model dtype before the fp32 softmax (patch_speech_encoder.py:853-890)
        s[n][e] = ok ? s[n][e] * sl2 : -INFINITY;
        mx[h] = fmaxf(mx[h], s[n][e]);
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
      corr[h] = (m_run[h] == -INFINITY) ? 0.f : exp2f(m_run[h] - msafe[h]);
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
      const float p0 = exp2f(s[n][0] - msafe[0]), p1 = exp2f(s[n][1] - msafe[0]);
      const float p2 = exp2f(s[n][2] - msafe[1]), p3 = exp2f(s[n][3] - msafe[1]);
      ls[0] += p0 + p1;
      ls[1] += p2 + p3;
      // C fragments of two adjacent n8 tiles form the A fragment of one k16 step
      pf[n >> 1][(n & 1) * 2 + 0] = pack_bf16(p0, p1);
      pf[n >> 1][(n & 1) * 2 + 1] = pack_bf16(p2, p3);
    }
    l_run[0] += ls[0];
    l_run[1] += ls[1];
    // ---- O += P V ----
#pragma unroll
    for (int kk = 0; kk < KT / 16; ++kk) {
      const bf16* vrow = sV + (kk * 16 + (lane & 15)) * LDS;
#pragma unroll
      for (int n = 0; n < HD / 8; ++n) {
        uint32_t b0, b1;
        ldmatrix_x2_trans(b0, b1, vrow + n * 8);
        mma_bf16_16816(o_acc[n], pf[kk], b0, b1);
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
// Encoder KV ring append: K/V of the T new frames -> slots (prefix + i) % cap.  (patch_speech_encoder.py:797-821)
// ----------------------------------------------------------------------------------------------
__global__ void enc_kv_append_kernel(const bf16* __restrict__ qkv, bf16* k_ring, bf16* v_ring,
                                     const int* __restrict__ slots, const int* __restrict__ prefix, int T, int H,
                                     int HD, int cap) {
  const int b = blockIdx.y;
  const int slot = slots[b];
  const int pre = prefix[b];   // per batch entry
  const int chunks = H * HD / 8;
  for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < T * chunks; u += gridDim.x * blockDim.x) {
    const int i = u / chunks, c = u % chunks;
    const int head = (c * 8) / HD, d = (c * 8) % HD;
    const bf16* src = qkv + static_cast<size_t>(b * T + i) * (3 * H * HD);
    const size_t dst = ((static_cast<size_t>(slot) * H + head) * cap + (pre + i) % cap) * HD + d;
    *reinterpret_cast<uint4*>(k_ring + dst) = *reinterpret_cast<const uint4*>(src + H * HD + c * 8);
    *reinterpret_cast<uint4*>(v_ring + dst) = *reinterpret_cast<const uint4*>(src + 2 * H * HD + c * 8);
  }
}

// ----------------------------------------------------------------------------------------------
// LLM paged KV append (un-rotated K, V) for packed new tokens.  (patch_llm.py:280-284)
// active[b] == 0 -> stream finished (EOS): nothing is appended.
// ----------------------------------------------------------------------------------------------
__global__ void llm_kv_append_kernel(const bf16* __restrict__ qkv, PagedKV kv, const int* __restrict__ slots,
                                     const int* __restrict__ tok_base, const int* __restrict__ Tn,
                                     const int* __restrict__ active, int H) {
  const int b = blockIdx.y;
  if (active && !active[b]) return;
  const int slot = slots[b];
  const int T = Tn[b];
  const int HD = kv.head_dim;
  const int chunks = kv.kv_heads * HD / 8;
  const int base = kv.kv_len[slot];
  const int sys_len = kv.sys_len[slot], ring_start = kv.ring_start[slot];
  const int* table = kv.page_table + static_cast<size_t>(slot) * kv.pages_per_stream;
  const int ldq = (H + 2 * kv.kv_heads) * HD;
  for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < T * chunks; u += gridDim.x * blockDim.x) {
    const int i = u / chunks, c = u % chunks;
    const int head = (c * 8) / HD, d = (c * 8) % HD;
    const bf16* src = qkv + static_cast<size_t>(tok_base[b] + i) * ldq + H * HD;
    const int sl = kv_slot(base + i, sys_len, ring_start);
    *reinterpret_cast<uint4*>(kv.pool + kv_offset(kv, table, sl, 0, head) + d) =
        *reinterpret_cast<const uint4*>(src + c * 8);
    *reinterpret_cast<uint4*>(kv.pool + kv_offset(kv, table, sl, 1, head) + d) =
        *reinterpret_cast<const uint4*>(src + kv.kv_heads * HD + c * 8);
  }
}

// ----------------------------------------------------------------------------------------------
// Decode attention (T = 1), split along the key axis.  One CTA = (split, kv_head, stream), 4 warps;
// each warp owns whole keys: 16 lanes... layout: a key's 128 elements are covered by 16 lanes x
// (4 + 4) elements (the half-split RoPE pair d / d+64 lives in the same lane), two keys per warp pass.
// Partial (m, l, o[HD]) per (stream, q head, split) -> decode_combine_kernel.
// Algorithmic bytes: 2 * L * Hkv * HD * 2 B per layer per stream.
// ----------------------------------------------------------------------------------------------
struct DecodeParams {
  const bf16* qkv;        // [n, (H + 2 Hkv) * HD] : the single new token of each stream (already appended to KV)
  PagedKV kv;             // kv_len = length BEFORE this token
  const int* slots;
  const bf162* rope;
  float* part_o;          // [n][H][splits][HD]
  float* part_ml;         // [n][H][splits][2]
  int H;
  int splits;
  float scale_log2;
};

template <int HD, int GROUP>
__global__ void __launch_bounds__(128)
decode_attention_kernel(const DecodeParams p) {
  static_assert(HD == 128, "decode kernel is specialised for head_dim 128");
  const int split = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int sub = lane >> 4;        // which of the two keys of this warp pass
  const int l16 = lane & 15;        // owns elements [4*l16, 4*l16+4) and [64 + 4*l16, 64 + 4*l16 + 4)
  const int slot = p.slots[b];
  const int L = p.kv.kv_len[slot] + 1;
  const int sys_len = p.kv.sys_len[slot], ring_start = p.kv.ring_start[slot];
  const int* table = p.kv.page_table + static_cast<size_t>(slot) * p.kv.pages_per_stream;
  const int per = (L + p.splits - 1) / p.splits;
  const int j0 = split * per, j1 = min(L, j0 + per);

  // rotated queries of the GROUP q heads sharing this kv head (position L-1), pre-multiplied by scale*log2e
  float q[GROUP][8];
  {
    const int ldq = (p.H + 2 * p.kv.kv_heads) * HD;
    const bf162* rp = p.rope + static_cast<size_t>(L - 1) * (HD / 2) + 4 * l16;
#pragma unroll
    for (int hq = 0; hq < GROUP; ++hq) {
      const bf16* src = p.qkv + static_cast<size_t>(b) * ldq + (head * GROUP + hq) * HD;
      uint2 lo = *reinterpret_cast<const uint2*>(src + 4 * l16);
      uint2 hi = *reinterpret_cast<const uint2*>(src + 64 + 4 * l16);
      const uint32_t wl[2] = {lo.x, lo.y}, wh[2] = {hi.x, hi.y};
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        float2 a = unpack_bf16(wl[j]), bb = unpack_bf16(wh[j]);
        float2 c0 = __bfloat1622float2(rp[2 * j]), c1 = __bfloat1622float2(rp[2 * j + 1]);
        // HF rounds the rotated q to bf16 (apply_rotary_pos_emb in the model dtype)
        q[hq][2 * j] = bf16_round(a.x * c0.x - bb.x * c0.y) * p.scale_log2;
        q[hq][2 * j + 1] = bf16_round(a.y * c1.x - bb.y * c1.y) * p.scale_log2;
        q[hq][4 + 2 * j] = bf16_round(bb.x * c0.x + a.x * c0.y) * p.scale_log2;
        q[hq][4 + 2 * j + 1] = bf16_round(bb.y * c1.x + a.y * c1.y) * p.scale_log2;
      }
    }
  }
  float m[GROUP], l[GROUP], o[GROUP][8];
#pragma unroll
  for (int hq = 0; hq < GROUP; ++hq) {
    m[hq] = -INFINITY; l[hq] = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) o[hq][e] = 0.f;
  }
  // each warp pass handles 2 keys; 4 warps -> 8 keys per CTA iteration
  for (int jb = j0 + warp * 2; jb < j1; jb += 8) {   // warp-uniform trip count (full-mask shuffles below)
    const int j = jb + sub;
    const bool valid = j < j1;
    float kr[8], vv[8];
    if (valid) {
      const int sl = kv_slot(j, sys_len, ring_start);
      const bf16* kp = p.kv.pool + kv_offset(p.kv, table, sl, 0, head);
      const bf16* vp = p.kv.pool + kv_offset(p.kv, table, sl, 1, head);
      uint2 klo = ld_nc_u2(kp + 4 * l16), khi = ld_nc_u2(kp + 64 + 4 * l16);
      uint2 vlo = ld_nc_u2(vp + 4 * l16), vhi = ld_nc_u2(vp + 64 + 4 * l16);
      const bf162* rp = p.rope + static_cast<size_t>(j) * (HD / 2) + 4 * l16;
      const uint32_t wl[2] = {klo.x, klo.y}, wh[2] = {khi.x, khi.y};
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        float2 a = unpack_bf16(wl[e]), bb = unpack_bf16(wh[e]);
        float2 c0 = __bfloat1622float2(rp[2 * e]), c1 = __bfloat1622float2(rp[2 * e + 1]);
        kr[2 * e] = a.x * c0.x - bb.x * c0.y;
        kr[2 * e + 1] = a.y * c1.x - bb.y * c1.y;
        kr[4 + 2 * e] = bb.x * c0.x + a.x * c0.y;
        kr[4 + 2 * e + 1] = bb.y * c1.x + a.y * c1.y;
      }
      float2 f;
      f = unpack_bf16(vlo.x); vv[0] = f.x; vv[1] = f.y;
      f = unpack_bf16(vlo.y); vv[2] = f.x; vv[3] = f.y;
      f = unpack_bf16(vhi.x); vv[4] = f.x; vv[5] = f.y;
      f = unpack_bf16(vhi.y); vv[6] = f.x; vv[7] = f.y;
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) { kr[e] = 0.f; vv[e] = 0.f; }
    }
#pragma unroll
    for (int hq = 0; hq < GROUP; ++hq) {
      float s = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) s += q[hq][e] * kr[e];
      // reduce over the 16 lanes of this key
      s += __shfl_xor_sync(0xffffffffu, s, 8);
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      if (!valid) s = -INFINITY;
      const float mn = fmaxf(m[hq], s);
      const float ms = (mn == -INFINITY) ? 0.f : mn;
      const float corr = (m[hq] == -INFINITY) ? 0.f : exp2f(m[hq] - ms);
      const float pe = exp2f(s - ms);
      m[hq] = mn;
      l[hq] = l[hq] * corr + pe;
#pragma unroll
      for (int e = 0; e < 8; ++e) o[hq][e] = o[hq][e] * corr + pe * vv[e];
    }
  }
  // ---- merge the 8 (warp, sub) partial states of this CTA through shared memory ----
  __shared__ float sm_m[GROUP][8], sm_l[GROUP][8];
  __shared__ float sm_o[GROUP][8][HD];
  const int ps = warp * 2 + sub;
#pragma unroll
  for (int hq = 0; hq < GROUP; ++hq) {
    if (l16 == 0) { sm_m[hq][ps] = m[hq]; sm_l[hq][ps] = l[hq]; }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      sm_o[hq][ps][4 * l16 + e] = o[hq][e];
      sm_o[hq][ps][64 + 4 * l16 + e] = o[hq][4 + e];
    }
  }
  __syncthreads();
  for (int idx = tid; idx < GROUP * HD; idx += 128) {
    const int hq = idx / HD, d = idx % HD;
    float mm = -INFINITY;
#pragma unroll
    for (int s8 = 0; s8 < 8; ++s8) mm = fmaxf(mm, sm_m[hq][s8]);
    const float ms = (mm == -INFINITY) ? 0.f : mm;
    float ll = 0.f, oo = 0.f;
#pragma unroll
    for (int s8 = 0; s8 < 8; ++s8) {
      const float w = (sm_m[hq][s8] == -INFINITY) ? 0.f : exp2f(sm_m[hq][s8] - ms);
      ll += w * sm_l[hq][s8];
      oo += w * sm_o[hq][s8][d];
    }
    const size_t pidx = (static_cast<size_t>(b) * p.H + head * GROUP + hq) * p.splits + split;
    p.part_o[pidx * HD + d] = oo;
    if (d == 0) { p.part_ml[pidx * 2] = mm; p.part_ml[pidx * 2 + 1] = ll; }
  }
}

// out[b, h, :] = sum_s w_s o_s / sum_s w_s l_s
__global__ void decode_combine_kernel(const float* __restrict__ part_o, const float* __restrict__ part_ml,
                                      bf16* __restrict__ out, int H, int HD, int splits) {
  const int bh = blockIdx.x;   // b * H + h
  const int d = threadIdx.x;
  float mm = -INFINITY;
  for (int s = 0; s < splits; ++s) mm = fmaxf(mm, part_ml[(static_cast<size_t>(bh) * splits + s) * 2]);
  const float ms = (mm == -INFINITY) ? 0.f : mm;
  float ll = 0.f, oo = 0.f;
  for (int s = 0; s < splits; ++s) {
    const size_t pi = static_cast<size_t>(bh) * splits + s;
    const float pm = part_ml[pi * 2];
    const float w = (pm == -INFINITY) ? 0.f : exp2f(pm - ms);
    ll += w * part_ml[pi * 2 + 1];
    if (d < HD) oo += w * part_o[pi * HD + d];
  }
  if (d < HD) out[static_cast<size_t>(bh) * HD + d] = __float2bfloat16_rn(ll > 0.f ? oo / ll : 0.f);
}

}  // namespace isst

// Single-token decode attention over the paged, un-rotated KV cache (SURVEY §2.3 L3-L6b;
// llama_sdpa_attention_new_forward, patch_llm.py:231-336 with T == 1).
//
// HBM-bandwidth-bound: per layer and stream the kernel must read K and V of every cached token once
// (2 * L * Hkv * HD * 2 B).  Design:
//   * grid (splits, kv_head, stream); 4 warps per CTA; the CTA walks its key range in 64-key tiles,
//     each warp owning 16 keys of a tile;
//   * K/V rows go global -> shared memory with cp.async (16 B, L1-bypassing) through a 3-stage ring,
//     so ~2 tiles (64 KB) per CTA are always in flight; 2 CTAs per SM;
//   * the 4 query heads of a GQA group are the M rows of mma.sync m16n8k16 (S = Q K^T), the output is
//     accumulated transposed (O^T = V^T P^T, N = the 4(+4 padding) query heads) so the fp32 accumulator
//     is 32 registers instead of 64;
//   * RoPE-on-read without per-key table traffic: a key at logical position j0 + i of a tile is rotated
//     by the in-tile angle theta * i only (a 64-row table that lives in 32 registers per thread, applied
//     with packed bf16 math), and the query is rotated by theta * (L - 1 - j0) once per tile
//     (R(a) q . R(b) k depends on a - b only).  Keys stay un-rotated in HBM, so sliding-window eviction
//     remains a page-table edit.
// Partial (m, l, o) per split go to decode_combine_kernel (attention.cuh).
#pragma once
#include "attention.cuh"

namespace isst {

struct DecodeParams2 {
  const bf16* qkv;          // [n, (H + 2 Hkv) * HD]: the new token of each stream (its K/V are already appended)
  PagedKV kv;               // kv_len = length BEFORE this token
  const int* slots;         // [n]
  const float* rope_cos;    // [max_pos][HD/2] fp32 angles table (HF LlamaRotaryEmbedding, fp32)
  const float* rope_sin;
  const bf162* rope_tile;   // [64][HD/2] (cos, sin) of the in-tile offsets 0..63, bf16 like HF's cos/sin cast
  float* part_o;            // [n][H][splits][HD]
  float* part_ml;           // [n][H][splits][2]
  int H;
  int splits;
  float scale_log2;
};

constexpr int kDecTile = 64;              // keys per tile
constexpr int kDecStages = 3;
constexpr int kDecLds = 128 + 8;          // padded row (elements): conflict-free ldmatrix
constexpr int kDecStageElems = 2 * kDecTile * kDecLds;                  // K tile + V tile
constexpr int kDecSmemBytes = kDecStages * kDecStageElems * 2 + 2 * 4 * 128 * 2;   // + double-buffered rotated Q

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* smem_row) {
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(smem_row));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_row) {
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(smem_row));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
// (lo, hi) -> (lo*c - hi*s, hi*c + lo*s) on packed bf16 pairs
__device__ __forceinline__ void rope_pair_bf16(uint32_t& lo, uint32_t& hi, uint32_t c, uint32_t s) {
  const bf162 l = *reinterpret_cast<bf162*>(&lo), h = *reinterpret_cast<bf162*>(&hi);
  const bf162 cc = *reinterpret_cast<bf162*>(&c), ss = *reinterpret_cast<bf162*>(&s);
  const bf162 nl = __hfma2(l, cc, __hneg2(__hmul2(h, ss)));
  const bf162 nh = __hfma2(h, cc, __hmul2(l, ss));
  lo = *reinterpret_cast<const uint32_t*>(&nl);
  hi = *reinterpret_cast<const uint32_t*>(&nh);
}

template <int GROUP>
__global__ void __launch_bounds__(128, 2)
decode_attention_mma_kernel(const DecodeParams2 p) {
  static_assert(GROUP == 4, "4:1 GQA");
  constexpr int HD = 128, LDS = kDecLds, TILE = kDecTile;
  extern __shared__ __align__(16) uint8_t dec_smem[];
  bf16* stage_base = reinterpret_cast<bf16*>(dec_smem);
  bf16* qbuf = stage_base + kDecStages * kDecStageElems;       // [2][GROUP][HD] rotated queries

  const int split = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int slot = p.slots[b];
  const int L = p.kv.kv_len[slot] + 1;
  const int sys_len = p.kv.sys_len[slot], ring_start = p.kv.ring_start[slot];
  const int* table = p.kv.page_table + static_cast<size_t>(slot) * p.kv.pages_per_stream;
  // key range of this split: whole tiles
  const int tiles_total = (L + TILE - 1) / TILE;
  const int tiles_per = (tiles_total + p.splits - 1) / p.splits;
  const int j_lo = split * tiles_per * TILE;
  const int j_hi = min(L, j_lo + tiles_per * TILE);
  const size_t pbase = (static_cast<size_t>(b) * p.H + head * GROUP) * p.splits + split;   // + hq * splits
  if (j_lo >= j_hi) {   // empty split: neutral partial
    for (int idx = tid; idx < GROUP * HD; idx += 128) {
      const int hq = idx / HD, d = idx % HD;
      p.part_o[(pbase + static_cast<size_t>(hq) * p.splits) * HD + d] = 0.f;
      if (d == 0) { p.part_ml[(pbase + static_cast<size_t>(hq) * p.splits) * 2] = -INFINITY; p.part_ml[(pbase + static_cast<size_t>(hq) * p.splits) * 2 + 1] = 0.f; }
    }
    return;
  }
  const int n_tiles = (j_hi - j_lo + TILE - 1) / TILE;

  // ---- in-tile RoPE table in registers: rows i = 16*warp + 8*n + g, freqs kk*16 + 8*r + 2*t4 (+1) ----
  uint32_t tcos[2][4][2], tsin[2][4][2];
#pragma unroll
  for (int n = 0; n < 2; ++n) {
    const bf162* row = p.rope_tile + static_cast<size_t>(16 * warp + 8 * n + g) * (HD / 2);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const float2 a = __bfloat1622float2(row[kk * 16 + 8 * r + 2 * t4]);        // (cos, sin) of freq d
        const float2 c = __bfloat1622float2(row[kk * 16 + 8 * r + 2 * t4 + 1]);    // freq d + 1
        tcos[n][kk][r] = pack_bf16(a.x, c.x);
        tsin[n][kk][r] = pack_bf16(a.y, c.y);
      }
  }

  // ---- tile loader: 64 keys x (K 256 B + V 256 B) = 2048 16-byte chunks, 16 per thread ----
  auto load_tile = [&](int t, int stage) {
    bf16* sK = stage_base + stage * kDecStageElems;
    bf16* sV = sK + TILE * LDS;
    const int j0 = j_lo + t * TILE;
    const int chunk = tid & 15;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int kl = (tid >> 4) + 8 * it;
      const int j = j0 + kl;
      const bool ok = j < j_hi;
      const bf16* ksrc = p.kv.pool;
      const bf16* vsrc = p.kv.pool;
      if (ok) {
        const int sl = kv_slot(j, sys_len, ring_start);
        ksrc = p.kv.pool + kv_offset(p.kv, table, sl, 0, head) + chunk * 8;
        vsrc = ksrc + static_cast<size_t>(p.kv.kv_heads) * kPageTokens * HD;      // V block of the same page
      }
      cp_async16(sK + kl * LDS + chunk * 8, ksrc, ok ? 16 : 0);                   // invalid keys: zero fill
      cp_async16(sV + kl * LDS + chunk * 8, vsrc, ok ? 16 : 0);
    }
  };

  // ---- rotated queries of one tile -> qbuf[buf]: Q' = R(theta * (L - 1 - j0)) q, 256 (d, d+64) pairs ----
  const int ldq = (p.H + 2 * p.kv.kv_heads) * HD;
  auto rotate_q = [&](int t, int buf) {
    const int delta = (L - 1) - (j_lo + t * TILE);
    const float* cs = p.rope_cos + static_cast<size_t>(delta) * (HD / 2);
    const float* sn = p.rope_sin + static_cast<size_t>(delta) * (HD / 2);
    bf16* dst = qbuf + buf * GROUP * HD;
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int idx = tid + 128 * it;          // 0..255
      const int hq = idx >> 6, d = idx & 63;
      const bf16* src = p.qkv + static_cast<size_t>(b) * ldq + (head * GROUP + hq) * HD;
      const float lo = __bfloat162float(src[d]), hi = __bfloat162float(src[d + 64]);
      const float c = cs[d], s = sn[d];
      dst[hq * HD + d] = __float2bfloat16_rn(lo * c - hi * s);
      dst[hq * HD + d + 64] = __float2bfloat16_rn(hi * c + lo * s);
    }
  };

  // prologue
#pragma unroll
  for (int s = 0; s < kDecStages - 1; ++s) {
    if (s < n_tiles) load_tile(s, s);
    cp_async_commit();
  }
  rotate_q(0, 0);

  float o[8][4];
#pragma unroll
  for (int mt = 0; mt < 8; ++mt) { o[mt][0] = o[mt][1] = o[mt][2] = o[mt][3] = 0.f; }
  float m_run = -INFINITY, l_run = 0.f;      // row g of S (query head g; rows >= GROUP are padding)

  for (int t = 0; t < n_tiles; ++t) {
    cp_async_wait<kDecStages - 2>();
    __syncthreads();                         // tile t landed for everyone; everyone is done with tile t-1
    {
      const int nt = t + kDecStages - 1;
      if (nt < n_tiles) load_tile(nt, nt % kDecStages);
      cp_async_commit();
    }
    if (t + 1 < n_tiles) rotate_q(t + 1, (t + 1) & 1);
    const bf16* sK = stage_base + (t % kDecStages) * kDecStageElems;
    const bf16* sV = sK + TILE * LDS;
    const bf16* q = qbuf + (t & 1) * GROUP * HD;

    // ---- S = Q' K'^T for this warp's 16 keys ----
    float s[2][4];
#pragma unroll
    for (int n = 0; n < 2; ++n) { s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f; }
    {
      uint32_t kf[8][4];
      // ldmatrix.x4: (keys 0-7, dims lo) (keys 0-7, dims hi) (keys 8-15, dims lo) (keys 8-15, dims hi)
      const bf16* krow = sK + (16 * warp + ((lane >> 4) << 3) + (lane & 7)) * LDS + (((lane >> 3) & 1) << 3);
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) ldmatrix_x4(kf[kk], krow + kk * 16);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
#pragma unroll
        for (int n = 0; n < 2; ++n)
#pragma unroll
          for (int r = 0; r < 2; ++r) rope_pair_bf16(kf[kk][2 * n + r], kf[kk + 4][2 * n + r], tcos[n][kk][r], tsin[n][kk][r]);
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        uint32_t qa[4] = {0u, 0u, 0u, 0u};
        if (g < GROUP) {
          qa[0] = *reinterpret_cast<const uint32_t*>(q + g * HD + kk * 16 + 2 * t4);
          qa[2] = *reinterpret_cast<const uint32_t*>(q + g * HD + kk * 16 + 8 + 2 * t4);
        }
        mma_bf16_16816(s[0], qa, kf[kk][0], kf[kk][1]);
        mma_bf16_16816(s[1], qa, kf[kk][2], kf[kk][3]);
      }
    }
    // ---- online softmax on row g (rows >= GROUP are zero padding; harmless) ----
    const int jw = j_lo + t * TILE + 16 * warp;
    float mx = m_run;
#pragma unroll
    for (int n = 0; n < 2; ++n)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = jw + n * 8 + 2 * t4 + e;
        s[n][e] = (j < j_hi) ? s[n][e] * p.scale_log2 : -INFINITY;
        mx = fmaxf(mx, s[n][e]);
      }
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    const float msafe = (mx == -INFINITY) ? 0.f : mx;
    const float corr = (m_run == -INFINITY) ? 0.f : exp2f(m_run - msafe);
    m_run = mx;
    const float p00 = exp2f(s[0][0] - msafe), p01 = exp2f(s[0][1] - msafe);
    const float p10 = exp2f(s[1][0] - msafe), p11 = exp2f(s[1][1] - msafe);
    l_run = l_run * corr + (p00 + p01 + p10 + p11);
    const uint32_t pb0 = pack_bf16(p00, p01);     // P^T B fragment: keys 2*t4, 2*t4+1 of query g
    const uint32_t pb1 = pack_bf16(p10, p11);     //                 keys 8 + 2*t4, +1
    // O^T columns are queries 2*t4, 2*t4+1: fetch their correction factors from the lanes that own those rows
    const float c_lo = __shfl_sync(0xffffffffu, corr, (2 * t4) * 4);
    const float c_hi = __shfl_sync(0xffffffffu, corr, (2 * t4 + 1) * 4);
    // ---- O^T += V^T P^T ----
    const bf16* vrow = sV + (16 * warp + ((lane >> 4) << 3) + (lane & 7)) * LDS + (((lane >> 3) & 1) << 3);
#pragma unroll
    for (int mt = 0; mt < 8; ++mt) {
      uint32_t va[4];
      ldmatrix_x4_trans(va, vrow + mt * 16);
      // matrices arrive as (keys 0-7,dims lo) (keys 0-7,dims hi) (keys 8-15,dims lo) (keys 8-15,dims hi)
      // = A fragment registers a0 (row g, k lo), a1 (row g+8, k lo), a2 (row g, k hi), a3 (row g+8, k hi)
      o[mt][0] *= c_lo; o[mt][1] *= c_hi; o[mt][2] *= c_lo; o[mt][3] *= c_hi;
      mma_bf16_16816(o[mt], va, pb0, pb1);
    }
  }
  cp_async_wait<0>();
  __syncthreads();                           // all warps done with the stage buffers: reuse them for the merge

  // ---- merge the 4 warps (each saw a quarter of every tile) ----
  float* sm_o = reinterpret_cast<float*>(dec_smem);            // [4 warps][GROUP][HD]
  float* sm_m = sm_o + 4 * GROUP * HD;                         // [4][GROUP]
  float* sm_l = sm_m + 4 * GROUP;
  l_run += __shfl_xor_sync(0xffffffffu, l_run, 1);
  l_run += __shfl_xor_sync(0xffffffffu, l_run, 2);
  if (g < GROUP && t4 == 0) { sm_m[warp * GROUP + g] = m_run; sm_l[warp * GROUP + g] = l_run; }
  if (t4 < 2) {
#pragma unroll
    for (int mt = 0; mt < 8; ++mt) {
      sm_o[(warp * GROUP + 2 * t4) * HD + mt * 16 + g] = o[mt][0];
      sm_o[(warp * GROUP + 2 * t4 + 1) * HD + mt * 16 + g] = o[mt][1];
      sm_o[(warp * GROUP + 2 * t4) * HD + mt * 16 + 8 + g] = o[mt][2];
      sm_o[(warp * GROUP + 2 * t4 + 1) * HD + mt * 16 + 8 + g] = o[mt][3];
    }
  }
  __syncthreads();
  for (int idx = tid; idx < GROUP * HD; idx += 128) {
    const int hq = idx / HD, d = idx % HD;
    float mm = -INFINITY;
#pragma unroll
    for (int w = 0; w < 4; ++w) mm = fmaxf(mm, sm_m[w * GROUP + hq]);
    const float ms = (mm == -INFINITY) ? 0.f : mm;
    float ll = 0.f, oo = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const float wm = sm_m[w * GROUP + hq];
      const float wgt = (wm == -INFINITY) ? 0.f : exp2f(wm - ms);
      ll += wgt * sm_l[w * GROUP + hq];
      oo += wgt * sm_o[(w * GROUP + hq) * HD + d];
    }
    const size_t pidx = pbase + static_cast<size_t>(hq) * p.splits;
    p.part_o[pidx * HD + d] = oo;
    if (d == 0) { p.part_ml[pidx * 2] = mm; p.part_ml[pidx * 2 + 1] = ll; }
  }
}

}  // namespace isst

// Single-token decode attention over the paged KV cache (SURVEY §2.3 L3-L6b;
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
//   * keys are stored rotated at their absolute index (attention.cuh header), so the inner loop has no
//     RoPE work at all: tiles of the pinned system prompt use the q_sys query variant, tiles of the
//     sliding part the ring variant.
// Partial (m, l, o) per split go to decode_combine_kernel (attention.cuh).
#pragma once
#include "attention.cuh"

namespace isst {

struct DecodeParams2 {
  const bf16* qkv;          // [n, (H + 2 Hkv) * HD]: the new token of each stream, q rotated in place (ring variant)
  const bf16* q_sys;        // [n, H * HD]: q rotated at (absolute index - evicted), for the pinned prefix keys
  PagedKV kv;               // kv_len = length BEFORE this token (its K/V are already appended)
  const int* slots;         // [n]
  bf16* out;                // [n][H * HD]: written directly when splits == 1 (no combine pass)
  float* part_o;            // [n][H][splits][HD]
  float* part_ml;           // [n][H][splits][2]
  int H;
  int splits;
  float scale_log2;
  // ---- fused RoPE + KV append (decode step inside llm_forward): the kernel itself completes the QKV row of its
  // (stream, kv head) from the GEMM's fp32 split partials, rotates q (both variants) and k, writes K / V of the
  // new token to the page and attends to it from shared memory; kv.kv_len then EXCLUDES the new token and
  // llm_rope_append_kernel is not launched.  fuse == 0: q / q_sys / cache were prepared by llm_rope_append_kernel.
  int fuse;
  const float* part;        // [n_part][n][ldq] fp32 split-K partials of the QKV GEMM, or null (then `qkv` holds bf16 rows)
  int n_part;
  long long part_stride;    // n * ldq
  const float2* tab_ring;   // [n][HD/2] (cos, sin) at the absolute index of the new token
  const float2* tab_sys;    // [n][HD/2] (cos, sin) at (absolute index - evicted): query variant for the pinned prefix
  const int* active;        // [n] or null: finished streams append nothing
};

constexpr int kDecTile = 64;              // keys per tile
constexpr int kDecStages = 3;
constexpr int kDecLds = 128 + 8;          // padded row (elements): conflict-free ldmatrix
constexpr int kDecStageElems = 2 * kDecTile * kDecLds;                  // K tile + V tile
constexpr int kDecSmemBytes = kDecStages * kDecStageElems * 2 + 2 * 4 * 128 * 2 + 2 * 128 * 2 + 2 * kDecStages * 8;   // + the two query variants + K/V of the new token + full / empty barriers
constexpr int kDecLoaders = 4;            // loader warps
constexpr int kDecThreads = 128 + 32 * kDecLoaders;   // 4 compute warps + the loader warps

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
// the mbarrier receives one arrival from this thread once all of its earlier cp.async copies have landed
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(bar))) : "memory");
}
__device__ __forceinline__ void dec_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(bar))), "r"(count));
}
__device__ __forceinline__ void dec_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(bar))) : "memory");
}
__device__ __forceinline__ void dec_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nDEC_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DEC_DONE;\nbra DEC_WAIT;\nDEC_DONE:\n}\n" ::"r"(
          static_cast<uint32_t>(__cvta_generic_to_shared(bar))),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void dec_bar_compute() { asm volatile("bar.sync 1, 128;" ::: "memory"); }   // the 4 compute warps
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
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
template <int GROUP>
__global__ void __launch_bounds__(kDecThreads, 2)
decode_attention_mma_kernel(const DecodeParams2 p) {
  pdl_launch_dependents();
  pdl_wait();
  static_assert(GROUP == 4, "4:1 GQA");
  constexpr int HD = 128, LDS = kDecLds, TILE = kDecTile;
  extern __shared__ __align__(16) uint8_t dec_smem[];
  bf16* stage_base = reinterpret_cast<bf16*>(dec_smem);
  bf16* qbuf = stage_base + kDecStages * kDecStageElems;       // [2][GROUP][HD] rotated queries
  bf16* newkv = qbuf + 2 * GROUP * HD;                         // [2][HD] rotated K and plain V of the new token (fused mode)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(newkv + 2 * HD);   // [kDecStages] tile landed (32 loader lanes arrive)
  uint64_t* empty_bar = full_bar + kDecStages;                  // [kDecStages] tile consumed (4 compute warps arrive)

  const int split = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int slot = p.slots[b];
  const int L_old = p.kv.kv_len[slot];
  const int L = p.fuse ? L_old : L_old + 1;      // keys read from the cache (fused: the new token comes from shared memory)
  const int sys_len = min(p.kv.sys_len[slot], L), ring_start = p.kv.ring_start[slot];
  const int* table = p.kv.page_table + static_cast<size_t>(slot) * p.kv.pages_per_stream;
  // tile list: [0, sys_len) in 64-key tiles (q_sys variant), then [sys_len, L) (ring variant); a split is a
  // contiguous range of whole tiles
  const int n_sys_tiles = (sys_len + TILE - 1) / TILE;
  const int tiles_total = n_sys_tiles + (L - sys_len + TILE - 1) / TILE;
  const int tiles_per = (tiles_total + p.splits - 1) / p.splits;
  const int t_lo = split * tiles_per, t_hi = min(tiles_total, t_lo + tiles_per);
  const size_t pbase = (static_cast<size_t>(b) * p.H + head * GROUP) * p.splits + split;   // + hq * splits
  if (t_lo >= t_hi) {   // empty split: neutral partial
    for (int idx = tid; idx < GROUP * HD; idx += kDecThreads) {
      const int hq = idx / HD, d = idx % HD;
      p.part_o[(pbase + static_cast<size_t>(hq) * p.splits) * HD + d] = 0.f;
      if (d == 0) { p.part_ml[(pbase + static_cast<size_t>(hq) * p.splits) * 2] = -INFINITY; p.part_ml[(pbase + static_cast<size_t>(hq) * p.splits) * 2 + 1] = 0.f; }
    }
    return;
  }
  const int n_tiles = t_hi - t_lo;
  auto tile_j0 = [&](int t) { return t < n_sys_tiles ? t * TILE : sys_len + (t - n_sys_tiles) * TILE; };
  auto tile_j1 = [&](int t) { return t < n_sys_tiles ? min(sys_len, t * TILE + TILE) : min(L, sys_len + (t - n_sys_tiles + 1) * TILE); };

  // ---- warp specialisation: warps 4.. only stream K / V tiles into the ring, warps 0-3 only compute ----
  if (tid == 0) {
    for (int s0 = 0; s0 < kDecStages; ++s0) { dec_mbar_init(&full_bar[s0], 32 * kDecLoaders); dec_mbar_init(&empty_bar[s0], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp >= 4) {
    constexpr int kPass = 4 / kDecLoaders;           // 16-key passes per loader warp and tile
    const int pass0 = (warp - 4) * kPass;
    // Tile loader: 64 keys x (K 256 B + V 256 B) = 2048 16-byte chunks.  Lane (g2 = lane / 16, chunk = lane % 16)
    // copies chunk `chunk` of the keys 8*grp .. 8*grp+7 for grp = 2*pass + g2; the 4 passes are split over the loader warps.
    // The keys of a tile sit in one segment (pinned prefix or ring), so the 8 slots of a group are consecutive and
    // touch at most two pages: two page-table lookups per group, fetched one tile ahead of their use.
    const int g2 = lane >> 4, chunk = lane & 15;
    const size_t page_elems = static_cast<size_t>(2) * p.kv.kv_heads * kPageTokens * HD;
    const bf16* head_base = p.kv.pool + static_cast<size_t>(head) * kPageTokens * HD + chunk * 8;
    const size_t v_off = static_cast<size_t>(p.kv.kv_heads) * kPageTokens * HD;     // V block of the same page
    int nx_s0[kPass], nx_pa[kPass], nx_pb[kPass];
    auto lookup = [&](int t) {
#pragma unroll
      for (int ps = 0; ps < kPass; ++ps) {
        const int grp = 2 * (pass0 + ps) + g2;
        const int jg = min(tile_j0(t) + 8 * grp, L - 1);
        nx_s0[ps] = kv_slot(jg, sys_len, ring_start);
        const int last = kv_slot(min(jg + 7, tile_j1(t) - 1 > jg ? tile_j1(t) - 1 : jg), sys_len, ring_start);
        nx_pa[ps] = table[nx_s0[ps] >> 4];
        nx_pb[ps] = table[last >> 4];
      }
    };
    lookup(t_lo);
    for (int ti = 0; ti < n_tiles; ++ti) {
      const int t = t_lo + ti, stage = ti % kDecStages;
      if (ti >= kDecStages) dec_mbar_wait(&empty_bar[stage], ((ti / kDecStages) - 1) & 1);
#pragma unroll
      for (int ps = 0; ps < kPass; ++ps) {
        const int grp = 2 * (pass0 + ps) + g2;
        bf16* sK = stage_base + stage * kDecStageElems + (8 * grp) * LDS + chunk * 8;
        bf16* sV = sK + TILE * LDS;
        const int n_ok = tile_j1(t) - (tile_j0(t) + 8 * grp);          // keys of this group inside the tile
        const int s0 = nx_s0[ps];
        const bf16* base_a = head_base + static_cast<size_t>(nx_pa[ps]) * page_elems;
        const bf16* base_b = head_base + static_cast<size_t>(nx_pb[ps]) * page_elems;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int sl = s0 + it;
          const bf16* src = (((sl ^ s0) & ~15) == 0 ? base_a : base_b) + (sl & 15) * HD;
          const bool ok = it < n_ok;
          if (!ok) src = p.kv.pool;
          cp_async16(sK + it * LDS, src, ok ? 16 : 0);                 // invalid keys: zero fill
          cp_async16(sV + it * LDS, src + v_off, ok ? 16 : 0);
        }
      }
      cp_async_arrive_noinc(&full_bar[stage]);
      if (t + 1 < t_hi) lookup(t + 1);
    }
    cp_async_wait_all();
    return;
  }

  const bool owns_new = p.fuse && t_hi == tiles_total;          // the split that holds the last tile also takes the new token
  if (p.fuse) {
    // ---- fused llm_rope_append: complete this (stream, kv head)'s slice of the QKV row, rotate, append ----
    // items: 256 q pairs (4 heads x 64), 64 k pairs, 64 v pairs; a pair is (d, d + 64) of one head (half-split RoPE)
    const int ldq = (p.H + 2 * p.kv.kv_heads) * HD;
    const bool writer = owns_new && (!p.active || p.active[b]);
    const int sl_new = kv_slot(L_old, p.kv.sys_len[slot], ring_start);
    const size_t koff = kv_offset(p.kv, table, sl_new, 0, head), voff = kv_offset(p.kv, table, sl_new, 1, head);
    // every thread owns three items (tid, tid + 128, tid + 256): all of their loads are issued before any use
    int col[3], dd[3];
    float a[3], bb[3];
    float2 cr[3], cs[3];
    const bool sys_new = L_old < p.kv.sys_len[slot];   // new token still inside the pinned prefix: prefix convention for its key
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int item = tid + 128 * k;
      if (item < 256) { dd[k] = item & 63; col[k] = (head * GROUP + (item >> 6)) * HD; }
      else if (item < 320) { dd[k] = item - 256; col[k] = (p.H + head) * HD; }
      else { dd[k] = item - 320; col[k] = (p.H + p.kv.kv_heads + head) * HD; }
      cr[k] = ((item >= 256 && sys_new) ? p.tab_sys : p.tab_ring)[static_cast<size_t>(b) * 64 + dd[k]];
      cs[k] = p.tab_sys[static_cast<size_t>(b) * 64 + dd[k]];
    }
    if (p.part) {
      // all partial loads of the thread's three items are issued before the first use (<= 8 k-splits; the sum runs in
      // split order like the row kernels')
      float va[3][8], vb[3][8];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float* p0 = p.part + static_cast<size_t>(b) * ldq + col[k] + dd[k];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          va[k][q] = 0.f; vb[k][q] = 0.f;
          if (q < p.n_part) { va[k][q] = __ldcg(p0 + q * p.part_stride); vb[k][q] = __ldcg(p0 + q * p.part_stride + 64); }
        }
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        float sa = ((va[k][0] + va[k][1]) + va[k][2]) + va[k][3], sb = ((vb[k][0] + vb[k][1]) + vb[k][2]) + vb[k][3];
#pragma unroll
        for (int q = 4; q < 8; ++q)
          if (q < p.n_part) { sa += va[k][q]; sb += vb[k][q]; }
        const float* p0 = p.part + static_cast<size_t>(b) * ldq + col[k] + dd[k];
        for (int q = 8; q < p.n_part; ++q) { sa += __ldcg(p0 + q * p.part_stride); sb += __ldcg(p0 + q * p.part_stride + 64); }
        a[k] = bf16_round(sa); bb[k] = bf16_round(sb);             // the projection output is a bf16 tensor in the reference
      }
    } else {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const bf16* src = p.qkv + static_cast<size_t>(b) * ldq + col[k] + dd[k];
        a[k] = __bfloat162float(src[0]); bb[k] = __bfloat162float(src[64]);
      }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int item = tid + 128 * k, d = dd[k];
      if (item >= 320) {                                         // V: plain
        const bf16 v0 = __float2bfloat16_rn(a[k]), v1 = __float2bfloat16_rn(bb[k]);
        newkv[HD + d] = v0; newkv[HD + d + 64] = v1;
        if (writer) { p.kv.pool[voff + d] = v0; p.kv.pool[voff + d + 64] = v1; }
        continue;
      }
      const bf16 lo = __float2bfloat16_rn(a[k] * cr[k].x - bb[k] * cr[k].y), hi = __float2bfloat16_rn(bb[k] * cr[k].x + a[k] * cr[k].y);
      if (item < 256) {
        const int hq = item >> 6;
        qbuf[hq * HD + d] = lo; qbuf[hq * HD + d + 64] = hi;
        qbuf[(GROUP + hq) * HD + d] = __float2bfloat16_rn(a[k] * cs[k].x - bb[k] * cs[k].y);
        qbuf[(GROUP + hq) * HD + d + 64] = __float2bfloat16_rn(bb[k] * cs[k].x + a[k] * cs[k].y);
      } else {
        newkv[d] = lo; newkv[d + 64] = hi;
        if (writer) { p.kv.pool[koff + d] = lo; p.kv.pool[koff + d + 64] = hi; }
      }
    }
  } else {
    const int ldq = (p.H + 2 * p.kv.kv_heads) * HD;
    // 2 variants x 4 heads x 128 dims = 128 16-byte chunks, one per thread
    const int v = tid >> 6, hq = (tid >> 4) & 3, c = tid & 15;
    const bf16* src = v == 0 ? p.qkv + static_cast<size_t>(b) * ldq + (head * GROUP + hq) * HD
                             : p.q_sys + static_cast<size_t>(b) * (p.H * HD) + (head * GROUP + hq) * HD;
    *reinterpret_cast<uint4*>(qbuf + (v * GROUP + hq) * HD + c * 8) = *reinterpret_cast<const uint4*>(src + c * 8);
  }
  dec_bar_compute();
  // A fragments of the current query variant live in registers (rows >= GROUP are zero padding)
  uint32_t qa0[8], qa2[8];
  auto load_q = [&](bool sys_variant) {
    const bf16* q = qbuf + (sys_variant ? GROUP * HD : 0) + (g & (GROUP - 1)) * HD + 2 * t4;
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
      qa0[kk] = g < GROUP ? *reinterpret_cast<const uint32_t*>(q + kk * 16) : 0u;
      qa2[kk] = g < GROUP ? *reinterpret_cast<const uint32_t*>(q + kk * 16 + 8) : 0u;
    }
  };
  bool q_is_sys = t_lo < n_sys_tiles;
  load_q(q_is_sys);

  float o[8][4];
#pragma unroll
  for (int mt = 0; mt < 8; ++mt) { o[mt][0] = o[mt][1] = o[mt][2] = o[mt][3] = 0.f; }
  float m_run = -INFINITY, l_run = 0.f;      // row g of S (query head g; rows >= GROUP are padding)

  for (int ti = 0; ti < n_tiles + (owns_new ? 1 : 0); ++ti) {
    const int t = t_lo + ti;
    const bool fresh = ti == n_tiles;        // fused mode, last iteration: the new token itself, from shared memory
    const bf16* sK;
    if (!fresh) {
      dec_mbar_wait(&full_bar[ti % kDecStages], (ti / kDecStages) & 1);   // tile landed
      sK = stage_base + (ti % kDecStages) * kDecStageElems;
      if (t == n_sys_tiles && ti > 0 && q_is_sys) { load_q(false); q_is_sys = false; }   // leaving the pinned prefix: ring variant
    } else {
      // every tile has landed and been consumed.  Stage 0 held real tiles: 64 finite K / V rows.  Row 0 becomes the
      // new token; the other 63 rows are masked.
      dec_bar_compute();
      bf16* s0 = stage_base;
      if (tid < 32) *reinterpret_cast<uint2*>(s0 + tid * 4) = *reinterpret_cast<const uint2*>(newkv + tid * 4);
      else if (tid < 64) *reinterpret_cast<uint2*>(s0 + TILE * LDS + (tid - 32) * 4) = *reinterpret_cast<const uint2*>(newkv + HD + (tid - 32) * 4);
      dec_bar_compute();
      sK = s0;
      const bool new_is_sys = L_old < p.kv.sys_len[slot];
      if (q_is_sys != new_is_sys) { load_q(new_is_sys); q_is_sys = new_is_sys; }
    }
    const bf16* sV = sK + TILE * LDS;
    // ---- S = Q K^T for this warp's 16 keys ----
    float s[2][4];
#pragma unroll
    for (int n = 0; n < 2; ++n) { s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f; }
    {
      // ldmatrix.x4: (keys 0-7, dims lo) (keys 0-7, dims hi) (keys 8-15, dims lo) (keys 8-15, dims hi)
      const bf16* krow = sK + (16 * warp + ((lane >> 4) << 3) + (lane & 7)) * LDS + (((lane >> 3) & 1) << 3);
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        uint32_t kf[4];
        ldmatrix_x4(kf, krow + kk * 16);
        const uint32_t qa[4] = {qa0[kk], 0u, qa2[kk], 0u};
        mma_bf16_16816(s[0], qa, kf[0], kf[1]);
        mma_bf16_16816(s[1], qa, kf[2], kf[3]);
      }
    }
    // ---- online softmax on row g (rows >= GROUP are zero padding; harmless) ----
    const int jw = (fresh ? L_old : tile_j0(t)) + 16 * warp;
    const int j1 = fresh ? L_old + 1 : tile_j1(t);
    float mx = m_run;
#pragma unroll
    for (int n = 0; n < 2; ++n)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = jw + n * 8 + 2 * t4 + e;
        s[n][e] = (j < j1) ? s[n][e] * p.scale_log2 : -INFINITY;
        mx = fmaxf(mx, s[n][e]);
      }
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    const float msafe = (mx == -INFINITY) ? 0.f : mx;
    const float corr = (m_run == -INFINITY) ? 0.f : exp2_fast(m_run - msafe);
    m_run = mx;
    const float p00 = exp2_fast(s[0][0] - msafe), p01 = exp2_fast(s[0][1] - msafe);
    const float p10 = exp2_fast(s[1][0] - msafe), p11 = exp2_fast(s[1][1] - msafe);
    l_run = l_run * corr + (p00 + p01 + p10 + p11);
    const uint32_t pb0 = pack_bf16(p00, p01);     // P^T B fragment: keys 2*t4, 2*t4+1 of query g
    const uint32_t pb1 = pack_bf16(p10, p11);     //                 keys 8 + 2*t4, +1
    // O^T columns are queries 2*t4, 2*t4+1: fetch their correction factors from the lanes that own those rows
    const float c_lo = __shfl_sync(0xffffffffu, corr, (2 * t4) * 4);
    const float c_hi = __shfl_sync(0xffffffffu, corr, (2 * t4 + 1) * 4);
    const bool rescale = __any_sync(0xffffffffu, corr != 1.f);
    // ---- O^T += V^T P^T ----
    const bf16* vrow = sV + (16 * warp + ((lane >> 4) << 3) + (lane & 7)) * LDS + (((lane >> 3) & 1) << 3);
#pragma unroll
    for (int mt = 0; mt < 8; ++mt) {
      uint32_t va[4];
      ldmatrix_x4_trans(va, vrow + mt * 16);
      // matrices arrive as (keys 0-7,dims lo) (keys 0-7,dims hi) (keys 8-15,dims lo) (keys 8-15,dims hi)
      // = A fragment registers a0 (row g, k lo), a1 (row g+8, k lo), a2 (row g, k hi), a3 (row g+8, k hi)
      if (rescale) { o[mt][0] *= c_lo; o[mt][1] *= c_hi; o[mt][2] *= c_lo; o[mt][3] *= c_hi; }
      mma_bf16_16816(o[mt], va, pb0, pb1);
    }
    if (!fresh) {                            // this warp is done with the stage: hand it back to the loader
      __syncwarp();
      if (lane == 0) dec_mbar_arrive(&empty_bar[ti % kDecStages]);
    }
  }
  dec_bar_compute();                         // all compute warps done with the stage buffers (every tile has landed): reuse them for the merge

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
  dec_bar_compute();
  float* sm_w = sm_l + 4 * GROUP;                              // [4 warps][GROUP] merge weights, then [GROUP] m, [GROUP] l
  if (tid < GROUP) {
    const int hq = tid;
    float mm = -INFINITY;
#pragma unroll
    for (int w = 0; w < 4; ++w) mm = fmaxf(mm, sm_m[w * GROUP + hq]);
    const float ms = (mm == -INFINITY) ? 0.f : mm;
    float ll = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const float wm = sm_m[w * GROUP + hq];
      const float wgt = (wm == -INFINITY) ? 0.f : exp2_fast(wm - ms);
      sm_w[w * GROUP + hq] = wgt;
      ll += wgt * sm_l[w * GROUP + hq];
    }
    const size_t pidx = pbase + static_cast<size_t>(hq) * p.splits;
    p.part_ml[pidx * 2] = mm;
    p.part_ml[pidx * 2 + 1] = ll;
    sm_w[4 * GROUP + hq] = ll > 0.f ? 1.f / ll : 0.f;
  }
  dec_bar_compute();
#pragma unroll
  for (int hq = 0; hq < GROUP; ++hq) {
    const int d = tid;                                         // 128 threads == HD
    float oo = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) oo += sm_w[w * GROUP + hq] * sm_o[(w * GROUP + hq) * HD + d];
    if (p.splits == 1) p.out[(static_cast<size_t>(b) * p.H + head * GROUP + hq) * HD + d] = __float2bfloat16_rn(oo * sm_w[4 * GROUP + hq]);
    else p.part_o[(pbase + static_cast<size_t>(hq) * p.splits) * HD + d] = oo;
  }
}

}  // namespace isst

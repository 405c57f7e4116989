// LLM chunk-prefill attention on the 5th-generation tensor cores (llama_sdpa_attention_new_forward,
// patch_llm.py:231-336 with T > 1: causal, bottom-right aligned, 4:1 GQA, head_dim 128).
//
// One CTA = one (row tile of 128 query rows, kv head, stream).  Query rows are r = hq * T + i (the 4 query heads of
// the GQA group x the T new tokens), so a steady-state turn (T = 22 -> 88 rows) is one 128-row UMMA tile and K / V
// of the kv head are streamed exactly once.  Per 64-key tile:
//     S[128 x 64]  = Q[128 x 128] K^T      tcgen05.mma, both operands K-major in shared memory, S in TMEM (double-buffered)
//     P = exp2(S * scale - m)                4 softmax warps: thread = query row = TMEM lane; P (bf16) -> shared memory
//     O[128 x 128] += P[128 x 64] V          tcgen05.mma, A = P (K-major), B = V ([key][dim] = MN-major), O stays in TMEM
// The running maximum is only raised when it grows by more than 2^8 (then O in TMEM is rescaled in place), so P may
// reach 256 instead of 1 - exact in the same relative precision - and almost every tile skips the correction.
// Roles: warps 0-3 softmax / output, warp 4 MMA issuer (+ TMEM allocation), warp 5 TMA producer.  K / V tiles are
// TMA-staged straight from the paged pool: the pool of all layers is ONE 2-D tensor [rows of 128 dims] whose row
// coordinate encodes (layer, page, K|V, kv head, token in page); a 64-key tile is 4 pages x (K, V) x two 64-dim halves =
// 16 boxes of 16 rows x 128 B, written by the TMA unit in the SWIZZLE_128B layout the UMMA descriptors expect (no
// generic-proxy stores, no proxy fence) into a 2-stage ring with full / empty mbarriers.  Tiles are aligned to pages in
// SLOT space (the ring region starts at an arbitrary token offset after evictions): columns that fall before the ring
// start, behind the last key or in the unused tail of the pinned prefix's last page are masked in the softmax.
// 112 KB of shared memory, 256 TMEM columns and < 128 registers per thread: two CTAs per SM, so one CTA's softmax
// overlaps the other's MMAs.
// Keys are stored rotated at their absolute index (attention.cuh header): tiles of the pinned system prompt use the
// q_sys query variant, tiles of the sliding part the ring variant; the softmax warps swap the variant in shared memory
// at the boundary.
#pragma once
#include "attention.cuh"
#include "gemm_tcgen05.cuh"

namespace isst {

constexpr int kPaKT = 64;                 // keys per tile
constexpr int kPaStages = 2;
constexpr int kPaThreads = 224;           // 4 softmax warps + MMA warp + 2 loader warps
constexpr int kPaQBytes = 128 * 256;      // one query variant: 2 halves x [128 rows][128 B]
constexpr int kPaStageBytes = 4 * 64 * 128;   // K half 0, K half 1, V half 0, V half 1: each [64 keys][128 B]
constexpr int kPaPBytes = 128 * 128;      // P tile [128 rows][64 keys] bf16
constexpr int kPaSmemBytes = kPaQBytes + kPaStages * kPaStageBytes + kPaPBytes + 256 /*barriers*/;
constexpr int kPaTmemCols = 256;          // S0 [0,64) S1 [64,128) O [128,256)
constexpr float kPaRescaleThreshold = 8.0f;   // log2 units

// MN-major, SWIZZLE_128B shared-memory matrix descriptor for the V tile ([key][dim], dim contiguous): canonical layout
// ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units (cute::UMMA::make_umma_desc<Major::MN>): 64 dims = one 128-byte
// row per key, 8 keys = 1024 B (SBO), the next 64 dims live in the second half tile (LBO).
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// c_format F32, a/b BF16, A K-major, B MN-major (bit 16)
__host__ __device__ constexpr uint32_t make_idesc_bmn(int umma_m, int umma_n) {
  return tc::make_idesc(umma_m, umma_n) | (1u << 16);
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int GROUP, bool kSplit = false>
__global__ void __launch_bounds__(kPaThreads, 2)
prefill_attention_tc_kernel(const __grid_constant__ CUtensorMap tm_kv, const LlmAttnParams lp) {
  using namespace tc;
  pdl_launch_dependents();
  pdl_wait();
  constexpr int HD = 128, KT = kPaKT, NS = kPaStages;
  extern __shared__ __align__(1024) uint8_t pa_smem_raw[];
  uint8_t* smem = pa_smem_raw;                          // 1024-byte aligned (SWIZZLE_128B atoms)
  uint8_t* sQ = smem;                                   // [half][128 rows][128 B]: the query variant in use
  uint8_t* sStage = smem + kPaQBytes;                   // [NS][K h0 | K h1 | V h0 | V h1][64][128 B]
  uint8_t* sP = sStage + NS * kPaStageBytes;            // [128 rows][128 B]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sP + kPaPBytes);   // [NS] tile landed (TMA transaction bytes)
  uint64_t* empty_bar = full_bar + NS;                  // [NS] tile consumed (tcgen05.commit after P V)
  uint64_t* s_bar = empty_bar + NS;                     // [2]  S buffer ready (tcgen05.commit after Q K^T)
  uint64_t* p_bar = s_bar + 2;                          // P written (128 softmax threads arrive)
  uint64_t* pv_bar = p_bar + 1;                         // P V of a tile done (tcgen05.commit): P and O may be touched
  uint64_t* q_bar = pv_bar + 1;                         // ring query variant staged (128 softmax threads arrive)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(q_bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int KS = kSplit ? lp.key_splits : 1;                 // key splits (few streams), see LlmAttnParams
  const int split = kSplit ? blockIdx.x % KS : 0;
  const int b = blockIdx.z, head = blockIdx.y, row0 = (kSplit ? blockIdx.x / KS : blockIdx.x) * 128;
  const int slot = lp.slots[b];
  const int T = lp.T[b];
  const int tok0 = lp.tok_base[b];
  const int n_rows = GROUP * T;
  if (row0 >= n_rows) return;
  const int L = lp.kv.kv_len[slot] + T;
  const int sys_len = min(lp.kv.sys_len[slot], L), ring_start = lp.kv.ring_start[slot];
  const int* table = lp.kv.page_table + static_cast<size_t>(slot) * lp.kv.pages_per_stream;
  // keys needed by this CTA: up to the largest visible index over its rows
  const int rmax = min(row0 + 128, n_rows) - 1;
  const int i_hi = (row0 / T == rmax / T) ? (rmax % T) : (T - 1);
  const int key_end = L - T + i_hi + 1;
  const int sys_end = min(sys_len, key_end);
  const int n_sys_tiles = (sys_end + KT - 1) / KT;
  // ring part: logical keys [sys_len, key_end) live in slots [ring_start, ring_slot_end); tiles of 64 slots start at the
  // page that holds ring_start
  const int ring_slot_end = ring_start + (key_end - sys_end);
  const int sb0 = (ring_start >> 4) << 4;
  const int n_tiles_all = n_sys_tiles + (key_end > sys_end ? (ring_slot_end - sb0 + KT - 1) / KT : 0);
  // this CTA's key tiles: [t_lo, t_lo + n_tiles); u = t - t_lo indexes ring stages / barrier phases
  const int tiles_per = (n_tiles_all + KS - 1) / KS;
  const int t_lo = kSplit ? min(n_tiles_all, split * tiles_per) : 0;
  const int n_tiles = kSplit ? min(n_tiles_all, t_lo + tiles_per) - t_lo : n_tiles_all;
  if (kSplit && n_tiles <= 0) {                                     // empty split (KS > 1 only): neutral partials
    for (int idx = tid; idx < 128 * HD; idx += kPaThreads) {
      const int r = row0 + idx / HD, d = idx % HD;
      if (r >= n_rows) break;
      const size_t pi = (static_cast<size_t>(tok0 + r % T) * lp.H + head * GROUP + r / T) * KS + split;
      lp.part_o[pi * HD + d] = 0.f;
      if (d == 0) { lp.part_ml[pi * 2] = -INFINITY; lp.part_ml[pi * 2 + 1] = 0.f; }
    }
    return;
  }
  // tile t: slot of column 0, logical key index of column 0 (may lie before the first valid key), valid key range
  auto tile_sb = [&](int t) { return t < n_sys_tiles ? t * KT : sb0 + (t - n_sys_tiles) * KT; };
  auto tile_jbase = [&](int t) { return t < n_sys_tiles ? t * KT : sb0 + (t - n_sys_tiles) * KT - ring_start + sys_len; };
  auto tile_jlo = [&](int t) { return t < n_sys_tiles ? t * KT : max(sys_len, tile_jbase(t)); };
  auto tile_jhi = [&](int t) { return t < n_sys_tiles ? min(sys_end, t * KT + KT) : min(key_end, tile_jbase(t) + KT); };

  // ---- set-up: barriers, TMEM, both query variants (generic stores into the swizzled K-major layout) ----
  if (tid == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_kv)) : "memory");
    for (int s = 0; s < NS; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1);
    mbar_init(p_bar, 128);
    mbar_init(pv_bar, 1);
    mbar_init(q_bar, 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                 "r"(static_cast<uint32_t>(kPaTmemCols))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // one query variant: 128 rows x 16 chunks of 16 B into the swizzled K-major layout; rows beyond n_rows are zero
  auto stage_q = [&](bool sys_variant, int t0, int nthr) {
    const int ldq = (lp.H + 2 * lp.kv.kv_heads) * HD;
    for (int u = t0; u < 128 * 16; u += nthr) {
      const int rl = u >> 4, ch = u & 15;
      const int r = row0 + rl;
      uint4 val = make_uint4(0u, 0u, 0u, 0u);
      if (r < n_rows) {
        const int hq = r / T, i = r % T;
        const int qh = head * GROUP + hq;
        const bf16* src = sys_variant ? lp.q_sys + static_cast<size_t>(tok0 + i) * (lp.H * HD) + qh * HD
                                      : lp.qkv + static_cast<size_t>(tok0 + i) * ldq + qh * HD;
        val = *reinterpret_cast<const uint4*>(src + ch * 8);
      }
      const int h = ch >> 3, c = ch & 7;
      *reinterpret_cast<uint4*>(sQ + h * (128 * 128) + rl * 128 + ((c ^ (rl & 7)) << 4)) = val;
    }
    fence_proxy_async_smem();            // generic-proxy stores -> visible to the tensor core (async proxy)
  };
  stage_q(t_lo < n_sys_tiles, tid, kPaThreads);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp >= 5) {
    // ================= TMA producer: one lane; a tile = 4 pages x (K h0, K h1, V h0, V h1) boxes of 16 rows x 128 B =====
    if (warp == 5 && lane == 0) {
      const int Hkv = lp.kv.kv_heads;
      // L2 look-ahead: the ring holds one tile beyond the one in use (two CTAs of 112 KB per SM), i.e. 32 KB per CTA in
      // flight - at HBM latency that is ~20 KB/us per SM; tiles u+2 .. are therefore asked into L2 ahead of their turn
      auto prefetch_tile = [&](int u) {
        const int pg0 = tile_sb(t_lo + u) >> 4;
#pragma unroll
        for (int pp = 0; pp < KT / kPageTokens; ++pp) {
          const int page = table[min(pg0 + pp, lp.kv.pages_per_stream - 1)];
          const int rk = lp.kv_row0 + ((page * 2) * Hkv + head) * kPageTokens;
          const int rv = rk + Hkv * kPageTokens;
          tma_prefetch_l2_2d(&tm_kv, 0, rk); tma_prefetch_l2_2d(&tm_kv, 64, rk);
          tma_prefetch_l2_2d(&tm_kv, 0, rv); tma_prefetch_l2_2d(&tm_kv, 64, rv);
        }
      };
      for (int u = 2; u < min(n_tiles, 2 + lp.l2_ahead); ++u) prefetch_tile(u);
      for (int u = 0; u < n_tiles; ++u) {
        if (lp.l2_ahead > 0 && u + 2 + lp.l2_ahead - 1 < n_tiles) prefetch_tile(u + 2 + lp.l2_ahead - 1);
        const int stage = u % NS;
        if (u >= NS) mbar_wait(&empty_bar[stage], ((u / NS) - 1) & 1);      // P V of the tile that used this stage is done
        uint8_t* dK = sStage + stage * kPaStageBytes;
        mbar_expect_tx(&full_bar[stage], kPaStageBytes);
        const int pg0 = tile_sb(t_lo + u) >> 4;
#pragma unroll
        for (int pp = 0; pp < KT / kPageTokens; ++pp) {
          // pages past the stream's table hold masked columns only: any valid page will do
          const int page = table[min(pg0 + pp, lp.kv.pages_per_stream - 1)];
          const int rk = lp.kv_row0 + ((page * 2) * Hkv + head) * kPageTokens;
          const int rv = rk + Hkv * kPageTokens;
          uint8_t* d = dK + pp * (kPageTokens * 128);
          tma_load_2d(d, &tm_kv, &full_bar[stage], 0, rk);
          tma_load_2d(d + 64 * 128, &tm_kv, &full_bar[stage], 64, rk);
          tma_load_2d(d + 2 * (64 * 128), &tm_kv, &full_bar[stage], 0, rv);
          tma_load_2d(d + 3 * (64 * 128), &tm_kv, &full_bar[stage], 64, rv);
        }
      }
    }
  } else if (warp == 4) {
    // ================= MMA issuer =================
    constexpr uint32_t idesc_s = make_idesc(128, KT);          // S = Q K^T : N = 64 keys
    constexpr uint32_t idesc_o = make_idesc_bmn(128, HD);      // Ot = P V  : N = 128 dims, B MN-major
    auto mma_s = [&](int t) {                                  // t: local tile index
      const int stage = t % NS;
      mbar_wait(&full_bar[stage], (t / NS) & 1);
      if (t_lo + t == n_sys_tiles && t > 0) mbar_wait(q_bar, 0);   // the softmax warps swapped in the ring variant
      tcgen05_fence_after();
      if (lane == 0) {
        const uint32_t q = smem_u32(sQ);
        const uint32_t k = smem_u32(sStage + stage * kPaStageBytes);
        const uint32_t tS = tmem_base + (t & 1) * KT;
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {                 // 8 k-steps of 16 dims; dims >= 64 live in the second half tile
          const uint32_t hsel = kk >> 2, koff = (kk & 3) * 32;
          umma_bf16(tS, make_smem_desc(q + hsel * (128 * 128) + koff), make_smem_desc(k + hsel * (64 * 128) + koff), idesc_s,
                    kk > 0 ? 1u : 0u);
        }
        umma_commit(&s_bar[t & 1]);
      }
      __syncwarp();
    };
    // Software-pipelined by one tile: S(t+1) = Q K(t+1)^T is issued BEFORE the softmax of tile t is awaited (S is
    // double-buffered in TMEM), so the softmax warps find their next S tile ready and run back to back; without it every
    // tile serialises S MMA -> softmax -> P V MMA.
    // Buffer safety: S(t+1) overwrites the buffer of S(t-1), whose softmax finished before P(t-1) was announced (waited
    // for in the previous iteration); K(t+1) sits in the other ring stage than K(t) / V(t).
    if (n_tiles > 0) mma_s(0);
    for (int t = 0; t < n_tiles; ++t) {
      if (t + 1 < n_tiles) mma_s(t + 1);
      mbar_wait(p_bar, t & 1);
      tcgen05_fence_after();
      if (lane == 0) {
        const int stage = t % NS;
        const uint32_t pa = smem_u32(sP);
        const uint32_t v = smem_u32(sStage + stage * kPaStageBytes + 2 * (64 * 128));
        const uint32_t tO = tmem_base + 2 * KT;
#pragma unroll
        for (int kk = 0; kk < KT / 16; ++kk)                   // 4 k-steps of 16 keys
          umma_bf16(tO, make_smem_desc(pa + kk * 32), make_smem_desc_mn(v + kk * (16 * 128), 64 * 128, 1024), idesc_o,
                    (t > 0 || kk > 0) ? 1u : 0u);                // O accumulates in TMEM over all tiles
        umma_commit(pv_bar);
        umma_commit(&empty_bar[stage]);
      }
      __syncwarp();
    }
  } else {
    // ================= softmax + output: thread = query row = TMEM lane =================
    const int rl = tid;                                        // 0..127
    const int r = row0 + rl;
    const bool live = r < n_rows;
    const int hq = live ? r / T : 0, i = live ? r % T : 0;
    const int qhi = live ? L - T + i + 1 : 0;                  // causal, bottom-right aligned
    const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
    float m_run = -INFINITY, l_run = 0.f;
    for (int t = 0; t < n_tiles; ++t) {                        // t: local tile index
      const int jb = tile_jbase(t_lo + t), jlo = tile_jlo(t_lo + t), jhi = min(tile_jhi(t_lo + t), qhi);
      mbar_wait(&s_bar[t & 1], (t >> 1) & 1);
      tcgen05_fence_after();
      uint32_t sr[4][16];
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) tmem_ld16_issue(tmem_base + lane_addr + (t & 1) * KT + q4 * 16, sr[q4]);
      tmem_ld_wait();
      float mx = -INFINITY;
      if (jb >= jlo && jb + KT <= jhi) {                       // interior tile for this row: no mask
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4)
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const float sv = __uint_as_float(sr[q4][e]) * lp.scale_log2;
            sr[q4][e] = __float_as_uint(sv);
            mx = fmaxf(mx, sv);
          }
      } else {
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4)
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const int j = jb + q4 * 16 + e;
            const float sv = (j >= jlo && j < jhi) ? __uint_as_float(sr[q4][e]) * lp.scale_log2 : -INFINITY;
            sr[q4][e] = __float_as_uint(sv);
            mx = fmaxf(mx, sv);
          }
      }
      // lazy running maximum: raise it only when the tile exceeds it by more than the threshold
      float corr = 1.f;
      const bool raise = mx > m_run + kPaRescaleThreshold || (m_run == -INFINITY && mx != -INFINITY);
      if (raise) {
        corr = (m_run == -INFINITY) ? 0.f : exp2_fast(m_run - mx);
        m_run = mx;
        l_run *= corr;
      }
      const float msafe = (m_run == -INFINITY) ? 0.f : m_run;
      // P V of the previous tile must be complete before P is overwritten / O is rescaled
      if (t > 0) { mbar_wait(pv_bar, (t - 1) & 1); tcgen05_fence_after(); }
      if (t > 0 && __any_sync(0xffffffffu, raise)) {
        // ---- rare: rescale this warp's 32 rows of O in TMEM (rows that did not raise use corr = 1) ----
#pragma unroll 1
        for (int c8 = 0; c8 < HD / 32; ++c8) {
          uint32_t orr[2][16];
          tmem_ld16_issue(tmem_base + lane_addr + 2 * KT + c8 * 32, orr[0]);
          tmem_ld16_issue(tmem_base + lane_addr + 2 * KT + c8 * 32 + 16, orr[1]);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e) orr[e >> 4][e & 15] = __float_as_uint(__uint_as_float(orr[e >> 4][e & 15]) * corr);
          tmem_st16(tmem_base + lane_addr + 2 * KT + c8 * 32, orr[0]);
          tmem_st16(tmem_base + lane_addr + 2 * KT + c8 * 32 + 16, orr[1]);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      }
      float ls = 0.f;
      // P row (64 bf16 = 128 B = 8 chunks) into the swizzled K-major layout
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint32_t pk[4];
#pragma unroll
        for (int e2 = 0; e2 < 4; ++e2) {
          const int idx = c * 8 + e2 * 2;
          const float p0 = exp2_fast(__uint_as_float(sr[idx >> 4][idx & 15]) - msafe);
          const float p1 = exp2_fast(__uint_as_float(sr[(idx + 1) >> 4][(idx + 1) & 15]) - msafe);
          ls += p0 + p1;
          pk[e2] = pack_bf16(p0, p1);
        }
        *reinterpret_cast<uint4*>(sP + rl * 128 + ((c ^ (rl & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
      l_run += ls;
      fence_proxy_async_smem();                                // P stores -> visible to the tensor core
      tcgen05_fence_before();                                  // S loads / O stores above are complete
      mbar_arrive(p_bar);
      if (t_lo + t + 1 == n_sys_tiles && t + 1 < n_tiles) {
        // leaving the pinned prefix: Q K^T of every prefix tile is complete (s_bar), swap in the ring variant
        stage_q(false, tid, 128);
        mbar_arrive(q_bar);
      }
    }
    // ---- O = O / l ----
    mbar_wait(pv_bar, (n_tiles - 1) & 1);
    tcgen05_fence_after();
    const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
    bf16* dst = lp.out + static_cast<size_t>(tok0 + i) * (lp.H * HD) + (head * GROUP + hq) * HD;
    const size_t pi = kSplit ? (static_cast<size_t>(tok0 + i) * lp.H + head * GROUP + hq) * KS + split : 0;
    if (kSplit && live) { lp.part_ml[pi * 2] = m_run; lp.part_ml[pi * 2 + 1] = l_run; }
#pragma unroll 1
    for (int c8 = 0; c8 < HD / 32; ++c8) {
      uint32_t orr[2][16];
      tmem_ld16_issue(tmem_base + lane_addr + 2 * KT + c8 * 32, orr[0]);
      tmem_ld16_issue(tmem_base + lane_addr + 2 * KT + c8 * 32 + 16, orr[1]);
      tmem_ld_wait();
      if (kSplit && live) {                                    // un-normalised fp32 partial of this key split
        float* po = lp.part_o + pi * HD + c8 * 32;
#pragma unroll
        for (int c = 0; c < 8; ++c)
          *reinterpret_cast<float4*>(po + c * 4) = make_float4(__uint_as_float(orr[c >> 2][(c & 3) * 4]), __uint_as_float(orr[c >> 2][(c & 3) * 4 + 1]),
                                                               __uint_as_float(orr[c >> 2][(c & 3) * 4 + 2]), __uint_as_float(orr[c >> 2][(c & 3) * 4 + 3]));
      } else if (live) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 v;
          v.x = pack_bf16(__uint_as_float(orr[c >> 1][(c & 1) * 8 + 0]) * inv, __uint_as_float(orr[c >> 1][(c & 1) * 8 + 1]) * inv);
          v.y = pack_bf16(__uint_as_float(orr[c >> 1][(c & 1) * 8 + 2]) * inv, __uint_as_float(orr[c >> 1][(c & 1) * 8 + 3]) * inv);
          v.z = pack_bf16(__uint_as_float(orr[c >> 1][(c & 1) * 8 + 4]) * inv, __uint_as_float(orr[c >> 1][(c & 1) * 8 + 5]) * inv);
          v.w = pack_bf16(__uint_as_float(orr[c >> 1][(c & 1) * 8 + 6]) * inv, __uint_as_float(orr[c >> 1][(c & 1) * 8 + 7]) * inv);
          *reinterpret_cast<uint4*>(dst + c8 * 32 + c * 8) = v;
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 4) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(static_cast<uint32_t>(kPaTmemCols))
                 : "memory");
  }
}

}  // namespace isst

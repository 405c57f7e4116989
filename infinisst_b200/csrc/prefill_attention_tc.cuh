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
// Roles: warps 0-3 softmax / output, warp 4 MMA issuer (+ TMEM allocation), warps 5-6 K/V loaders (cp.async into the
// 128-byte-swizzled layout the UMMA descriptors expect, 2-stage ring with full / empty mbarriers).  112 KB of shared
// memory, 256 TMEM columns and < 128 registers per thread: two CTAs per SM, so one CTA's softmax overlaps the other's MMAs.
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
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int GROUP, bool kSplit = false>
__global__ void __launch_bounds__(kPaThreads, 2)
prefill_attention_tc_kernel(const LlmAttnParams lp) {
  using namespace tc;
  pdl_launch_dependents();
  pdl_wait();
  constexpr int HD = 128, KT = kPaKT, NS = kPaStages;
  extern __shared__ __align__(1024) uint8_t pa_smem_raw[];
  uint8_t* smem = pa_smem_raw;                          // 1024-byte aligned (SWIZZLE_128B atoms)
  uint8_t* sQ = smem;                                   // [half][128 rows][128 B]: the query variant in use
  uint8_t* sStage = smem + kPaQBytes;                   // [NS][K h0 | K h1 | V h0 | V h1][64][128 B]
  uint8_t* sP = sStage + NS * kPaStageBytes;            // [128 rows][128 B]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sP + kPaPBytes);   // [NS] tile landed (64 loader lanes arrive)
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
  const int n_tiles_all = n_sys_tiles + (key_end - sys_end + KT - 1) / KT;
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
  auto tile_k0 = [&](int t) { return t < n_sys_tiles ? t * KT : sys_end + (t - n_sys_tiles) * KT; };
  auto tile_k1 = [&](int t) { return t < n_sys_tiles ? min(sys_end, t * KT + KT) : min(key_end, sys_end + (t - n_sys_tiles + 1) * KT); };

  // ---- set-up: barriers, TMEM, both query variants (generic stores into the swizzled K-major layout) ----
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(&full_bar[s], 64); mbar_init(&empty_bar[s], 1); }
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
    // ================= K / V loaders: 64 lanes, each tile = 64 keys x (K 256 B + V 256 B) =================
    // lane (grp = l / 16, chunk = l % 16) copies chunk `chunk` of the keys 8*g .. 8*g+7 for g = 4*pass + grp.
    const int l64 = tid - 160;
    const int grp = l64 >> 4, chunk = l64 & 15;
    const int hh = chunk >> 3, cc = chunk & 7;
    const size_t page_elems = static_cast<size_t>(2) * lp.kv.kv_heads * kPageTokens * HD;
    const bf16* head_base = lp.kv.pool + static_cast<size_t>(head) * kPageTokens * HD + chunk * 8;
    const size_t v_off = static_cast<size_t>(lp.kv.kv_heads) * kPageTokens * HD;
    auto issue = [&](int u) {
      const int stage = u % NS;
      const int t = t_lo + u;
      uint8_t* dK = sStage + stage * kPaStageBytes + hh * (64 * 128);
      uint8_t* dV = dK + 2 * (64 * 128);
      const int k0 = tile_k0(t), k1 = tile_k1(t);
#pragma unroll
      for (int ps = 0; ps < 2; ++ps) {
        const int g = 4 * ps + grp;
        const int jg = min(k0 + 8 * g, L - 1);
        const int s0 = kv_slot(jg, sys_len, ring_start);
        const int last = kv_slot(min(jg + 7, k1 - 1 > jg ? k1 - 1 : jg), sys_len, ring_start);
        const bf16* base_a = head_base + static_cast<size_t>(table[s0 >> 4]) * page_elems;
        const bf16* base_b = head_base + static_cast<size_t>(table[last >> 4]) * page_elems;
        const int n_ok = k1 - (k0 + 8 * g);
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int kr = 8 * g + it;
          const int sl = s0 + it;
          const bf16* src = (((sl ^ s0) & ~15) == 0 ? base_a : base_b) + (sl & 15) * HD;
          const bool ok = it < n_ok;
          if (!ok) src = lp.kv.pool;
          const int off = kr * 128 + ((cc ^ (kr & 7)) << 4);
          cpa16(dK + off, src, ok ? 16 : 0);            // invalid keys: zero fill
          cpa16(dV + off, src + v_off, ok ? 16 : 0);
        }
      }
      cpa_commit();
    };
    // classic multi-stage cp.async pipeline inside the loader: tile t is published (proxy fence + arrive) once its
    // group has landed, while the next NS-1 tiles are already in flight
    // Tile t+NS-1 is requested as soon as its slot is free (P V of tile t-1 done), then tile t is published once
    // this thread's copies of it have landed: NS-1 tiles stay in flight.  (The MMA warp finishes tile t-1 without
    // needing tile t, so waiting for the slot before publishing cannot deadlock.)
    for (int t = 0; t < NS - 1; ++t) {
      if (t < n_tiles) issue(t);
      else cpa_commit();                                 // keep the group count uniform
    }
    for (int t = 0; t < n_tiles; ++t) {
      const int nt = t + NS - 1;
      if (nt < n_tiles) {
        if (nt >= NS) mbar_wait(&empty_bar[nt % NS], ((nt / NS) - 1) & 1);
        issue(nt);
      } else {
        cpa_commit();
      }
      cpa_wait<NS - 1>();                                // tile t landed (this thread's copies)
      fence_proxy_async_smem();
      mbar_arrive(&full_bar[t % NS]);
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
    for (int t = 0; t < n_tiles; ++t) {
      mma_s(t);                                                // (the other CTA on this SM fills the tensor pipe meanwhile)
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
      const int k0 = tile_k0(t_lo + t), k1 = min(tile_k1(t_lo + t), qhi);
      mbar_wait(&s_bar[t & 1], (t >> 1) & 1);
      tcgen05_fence_after();
      uint32_t sr[4][16];
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) tmem_ld16_issue(tmem_base + lane_addr + (t & 1) * KT + q4 * 16, sr[q4]);
      tmem_ld_wait();
      float mx = -INFINITY;
      if (k0 + KT <= k1) {                                     // interior tile for this row: no mask
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
            const int j = k0 + q4 * 16 + e;
            const float sv = (j < k1) ? __uint_as_float(sr[q4][e]) * lp.scale_log2 : -INFINITY;
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

// LLM chunk-prefill attention on the 5th-generation tensor cores (llama_sdpa_attention_new_forward,
// patch_llm.py:231-336 with T > 1: causal, bottom-right aligned, 4:1 GQA, head_dim 128).
//
// One CTA = one (row tile of 128 query rows, kv head, stream).  Query rows are r = hq * T + i (the 4 query heads of
// the GQA group x the T new tokens), so a steady-state turn (T = 22 -> 88 rows) is one 128-row UMMA tile and K / V
// of the kv head are streamed exactly once.  Per 64-key tile:
//     S[128 x 64]  = Q[128 x 128] K^T      tcgen05.mma, both operands K-major in shared memory, S in TMEM (double-buffered)
//     P = exp2(S * scale - m)                4 softmax warps: thread = query row = TMEM lane; P (bf16) -> shared memory
//     Ot[128 x 128] = P[128 x 64] V          tcgen05.mma, A = P (K-major), B = V ([key][dim] = MN-major), Ot in TMEM
//     O = O * corr + Ot                      the running output lives in the softmax threads' registers
// Roles: warps 0-3 softmax / output, warp 4 MMA issuer (+ TMEM allocation), warps 5-6 K/V loaders (cp.async into the
// 128-byte-swizzled layout the UMMA descriptors expect, 3-stage ring with full / empty mbarriers).
// Keys are stored rotated at their absolute index (attention.cuh header): tiles of the pinned system prompt use the
// q_sys query variant, tiles of the sliding part the ring variant; both variants sit in shared memory.
#pragma once
#include "attention.cuh"
#include "gemm_tcgen05.cuh"

namespace isst {

constexpr int kPaKT = 64;                 // keys per tile
constexpr int kPaStages = 3;
constexpr int kPaThreads = 224;           // 4 softmax warps + MMA warp + 2 loader warps
constexpr int kPaQBytes = 128 * 256;      // one query variant: 2 halves x [128 rows][128 B]
constexpr int kPaStageBytes = 4 * 64 * 128;   // K half 0, K half 1, V half 0, V half 1: each [64 keys][128 B]
constexpr int kPaPBytes = 128 * 128;      // P tile [128 rows][64 keys] bf16
constexpr int kPaSmemBytes = 2 * kPaQBytes + kPaStages * kPaStageBytes + kPaPBytes + 1024 /*align*/ + 256 /*barriers*/;
constexpr int kPaTmemCols = 256;          // S0 [0,64) S1 [64,128) Ot [128,256)

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
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int GROUP>
__global__ void __launch_bounds__(kPaThreads, 1)
prefill_attention_tc_kernel(const LlmAttnParams lp) {
  using namespace tc;
  pdl_launch_dependents();
  pdl_wait();
  constexpr int HD = 128, KT = kPaKT, NS = kPaStages;
  extern __shared__ uint8_t pa_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(pa_smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                   // [variant 0 ring | 1 sys][half][128 rows][128 B]
  uint8_t* sStage = smem + 2 * kPaQBytes;               // [NS][K h0 | K h1 | V h0 | V h1][64][128 B]
  uint8_t* sP = sStage + NS * kPaStageBytes;            // [128 rows][128 B]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sP + kPaPBytes);   // [NS] tile landed (64 loader lanes arrive)
  uint64_t* empty_bar = full_bar + NS;                  // [NS] tile consumed (tcgen05.commit after P V)
  uint64_t* s_bar = empty_bar + NS;                     // [2]  S buffer ready (tcgen05.commit after Q K^T)
  uint64_t* p_bar = s_bar + 2;                          // P written (128 softmax threads arrive)
  uint64_t* pv_bar = p_bar + 1;                         // Ot ready (tcgen05.commit after P V)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(pv_bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.z, head = blockIdx.y, row0 = blockIdx.x * 128;
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
  const int n_tiles = n_sys_tiles + (key_end - sys_end + KT - 1) / KT;
  auto tile_k0 = [&](int t) { return t < n_sys_tiles ? t * KT : sys_end + (t - n_sys_tiles) * KT; };
  auto tile_k1 = [&](int t) { return t < n_sys_tiles ? min(sys_end, t * KT + KT) : min(key_end, sys_end + (t - n_sys_tiles + 1) * KT); };

  // ---- set-up: barriers, TMEM, both query variants (generic stores into the swizzled K-major layout) ----
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(&full_bar[s], 64); mbar_init(&empty_bar[s], 1); }
    mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1);
    mbar_init(p_bar, 128);
    mbar_init(pv_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                 "r"(static_cast<uint32_t>(kPaTmemCols))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  {
    // 2 variants x 128 rows x 16 chunks of 16 B; rows beyond n_rows are zero
    const int ldq = (lp.H + 2 * lp.kv.kv_heads) * HD;
    for (int u = tid; u < 2 * 128 * 16; u += kPaThreads) {
      const int v = u >> 11, rl = (u >> 4) & 127, ch = u & 15;
      const int r = row0 + rl;
      uint4 val = make_uint4(0u, 0u, 0u, 0u);
      if (r < n_rows) {
        const int hq = r / T, i = r % T;
        const int qh = head * GROUP + hq;
        const bf16* src = v == 0 ? lp.qkv + static_cast<size_t>(tok0 + i) * ldq + qh * HD
                                 : lp.q_sys + static_cast<size_t>(tok0 + i) * (lp.H * HD) + qh * HD;
        val = *reinterpret_cast<const uint4*>(src + ch * 8);
      }
      const int h = ch >> 3, c = ch & 7;
      *reinterpret_cast<uint4*>(sQ + v * kPaQBytes + h * (128 * 128) + rl * 128 + ((c ^ (rl & 7)) << 4)) = val;
    }
    fence_proxy_async_smem();            // generic-proxy stores -> visible to the tensor core (async proxy)
  }
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
    auto issue = [&](int t) {
      const int stage = t % NS;
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
    for (int t = 0; t < NS - 1; ++t) {
      if (t < n_tiles) issue(t);
      else cpa_commit();                                 // keep the group count uniform
    }
    for (int t = 0; t < n_tiles; ++t) {
      cpa_wait<NS - 2>();                                // tile t landed (this thread's copies); NS-2 younger groups may fly
      fence_proxy_async_smem();
      mbar_arrive(&full_bar[t % NS]);                    // publish BEFORE blocking on a free slot (the MMA warp needs tile
      const int nt = t + NS - 1;                         // t+1 to get past tile t, whose completion frees the slot)
      if (nt < n_tiles) {
        if (nt >= NS) mbar_wait(&empty_bar[nt % NS], ((nt / NS) - 1) & 1);
        issue(nt);
      } else {
        cpa_commit();                                    // keep the group count uniform
      }
    }
  } else if (warp == 4) {
    // ================= MMA issuer =================
    constexpr uint32_t idesc_s = make_idesc(128, KT);          // S = Q K^T : N = 64 keys
    constexpr uint32_t idesc_o = make_idesc_bmn(128, HD);      // Ot = P V  : N = 128 dims, B MN-major
    auto mma_s = [&](int t) {
      const int stage = t % NS;
      mbar_wait(&full_bar[stage], (t / NS) & 1);
      tcgen05_fence_after();
      if (lane == 0) {
        const uint32_t q = smem_u32(sQ + (t < n_sys_tiles ? kPaQBytes : 0));
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
    if (n_tiles > 0) mma_s(0);
    for (int t = 0; t < n_tiles; ++t) {
      if (t + 1 < n_tiles) mma_s(t + 1);                       // S of the next tile while the softmax warps work on this one
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
                    kk > 0 ? 1u : 0u);
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
    float o[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) o[d] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;
    for (int t = 0; t < n_tiles; ++t) {
      const int k0 = tile_k0(t), k1 = min(tile_k1(t), qhi);
      mbar_wait(&s_bar[t & 1], (t >> 1) & 1);
      tcgen05_fence_after();
      uint32_t sr[4][16];
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) tmem_ld16_issue(tmem_base + lane_addr + (t & 1) * KT + q4 * 16, sr[q4]);
      tmem_ld_wait();
      float mx = m_run;
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4)
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int j = k0 + q4 * 16 + e;
          const float sv = (j < k1) ? __uint_as_float(sr[q4][e]) * lp.scale_log2 : -INFINITY;
          sr[q4][e] = __float_as_uint(sv);
          mx = fmaxf(mx, sv);
        }
      const float msafe = (mx == -INFINITY) ? 0.f : mx;
      const float corr = (m_run == -INFINITY) ? 0.f : exp2f(m_run - msafe);
      m_run = mx;
      float ls = 0.f;
      // P row (64 bf16 = 128 B = 8 chunks) into the swizzled K-major layout
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint32_t pk[4];
#pragma unroll
        for (int e2 = 0; e2 < 4; ++e2) {
          const int idx = c * 8 + e2 * 2;
          const float p0 = exp2f(__uint_as_float(sr[idx >> 4][idx & 15]) - msafe);
          const float p1 = exp2f(__uint_as_float(sr[(idx + 1) >> 4][(idx + 1) & 15]) - msafe);
          ls += p0 + p1;
          pk[e2] = pack_bf16(p0, p1);
        }
        *reinterpret_cast<uint4*>(sP + rl * 128 + ((c ^ (rl & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
      l_run = l_run * corr + ls;
      fence_proxy_async_smem();                                // P stores -> visible to the tensor core
      tcgen05_fence_before();                                  // the S loads above are complete (tmem_ld_wait)
      mbar_arrive(p_bar);
      // ---- O = O * corr + Ot ----
      mbar_wait(pv_bar, t & 1);
      tcgen05_fence_after();
#pragma unroll
      for (int c8 = 0; c8 < HD / 32; ++c8) {
        uint32_t orr[2][16];
        tmem_ld16_issue(tmem_base + lane_addr + 2 * KT + c8 * 32, orr[0]);
        tmem_ld16_issue(tmem_base + lane_addr + 2 * KT + c8 * 32 + 16, orr[1]);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) o[c8 * 32 + e] = o[c8 * 32 + e] * corr + __uint_as_float(orr[e >> 4][e & 15]);
      }
      tcgen05_fence_before();
    }
    if (live) {
      const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
      bf16* dst = lp.out + static_cast<size_t>(tok0 + i) * (lp.H * HD) + (head * GROUP + hq) * HD;
#pragma unroll
      for (int c = 0; c < HD / 8; ++c) {
        uint4 v;
        v.x = pack_bf16(o[c * 8 + 0] * inv, o[c * 8 + 1] * inv);
        v.y = pack_bf16(o[c * 8 + 2] * inv, o[c * 8 + 3] * inv);
        v.z = pack_bf16(o[c * 8 + 4] * inv, o[c * 8 + 5] * inv);
        v.w = pack_bf16(o[c * 8 + 6] * inv, o[c * 8 + 7] * inv);
        *reinterpret_cast<uint4*>(dst + c * 8) = v;
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

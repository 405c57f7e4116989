// wav2vec2 block-causal sliding-window self-attention of one chunk on the 5th-generation tensor cores, K / V TMA-staged
// from the per-layer KV rings (uni_mha_forward, patch_speech_encoder.py:692-933; mask closed form SURVEY §4.4: frame
// p sees keys j in [max(0, p - max_cache), min((p / blocksize + 1) * blocksize, prefix + T)); head_dim 64).
//
// One CTA = one (128-row tile of the T new frames, head, stream); T = 48 * m, so m <= 2 is one tile.  Per 64-key tile:
//     S[128 x 64]  = Q[128 x 64] K^T        tcgen05.mma (4 k-steps), both operands K-major in shared memory, S in TMEM x2
//     P = exp2(S * log2 e - m)               4 softmax warps: thread = query row = TMEM lane (q is pre-scaled by hd^-0.5)
//     O[128 x 64] += P[128 x 64] V           tcgen05.mma, A = P (K-major, shared memory), B = V ([key][dim] = MN-major)
// K / V: the rings of all layers are ONE 2-D tensor [rows of 64 dims = 128 B]; row = ((layer * streams + slot) * H + head)
// * cap + ring slot.  A tile is 4 boxes of 16 ring slots (cap is a multiple of 16, so a box never straddles the wrap);
// tiles are aligned to 16 slots in RING space - the window starts at an arbitrary frame - and the columns before the
// window start / behind its end are masked together with the band.  Same pipeline as prefill_attention_tc_kernel: TMA
// producer warp, MMA warp (S of tile t+1 issued before the softmax of tile t is awaited), lazy running maximum.
// 80 KB of shared memory and 256 TMEM columns: two CTAs per SM.
#pragma once
#include "prefill_attention_tc.cuh"

namespace isst {

constexpr int kEaKT = 64;                  // keys per tile
constexpr int kEaStages = 3;
constexpr int kEaThreads = 192;            // 4 softmax warps + MMA warp + TMA warp
constexpr int kEaQBytes = 128 * 128;       // [128 rows][64 dims]
constexpr int kEaStageBytes = 2 * 64 * 128;    // K [64 keys][128 B] | V [64 keys][128 B]
constexpr int kEaPBytes = 128 * 128;
constexpr int kEaSmemBytes = kEaQBytes + kEaStages * kEaStageBytes + kEaPBytes + 256;
constexpr int kEaTmemCols = 256;           // S0 [0,64) S1 [64,128) O [128,192)

struct EncAttnTcParams {
  const bf16* qkv;        // [tok, 3*H*64]; q rotated in place (enc_rope_append_kernel)
  bf16* out;              // [tok, H*64]
  const int* slots;       // [n] stream slot per batch entry
  const int* prefix;      // [n] frames encoded before this chunk
  int T, H, cap, max_cache, blocksize;
  int k_row0, v_row0;     // first row of this layer in the K / V ring tensor maps (rows of 64 dims)
  int rows_per_slot;      // H * cap
};

__global__ void __launch_bounds__(kEaThreads, 2)
encoder_attention_tc_kernel(const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_v,
                            const EncAttnTcParams ep) {
  using namespace tc;
  pdl_launch_dependents();
  pdl_wait();
  constexpr int HD = 64, KT = kEaKT, NS = kEaStages;
  extern __shared__ __align__(1024) uint8_t ea_smem_raw[];
  uint8_t* smem = ea_smem_raw;
  uint8_t* sQ = smem;                                   // [128 rows][128 B]
  uint8_t* sStage = smem + kEaQBytes;                   // [NS][K | V][64][128 B]
  uint8_t* sP = sStage + NS * kEaStageBytes;            // [128 rows][128 B]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sP + kEaPBytes);   // [NS] tile landed (TMA transaction bytes)
  uint64_t* empty_bar = full_bar + NS;                  // [NS] tile consumed (tcgen05.commit after P V)
  uint64_t* s_bar = empty_bar + NS;                     // [2]  S buffer ready
  uint64_t* p_bar = s_bar + 2;                          // P written (128 softmax threads arrive)
  uint64_t* pv_bar = p_bar + 1;                         // P V of a tile done
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(pv_bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.z, head = blockIdx.y, row0 = blockIdx.x * 128;
  const int T = ep.T;
  if (row0 >= T) return;
  const int slot = ep.slots[b];
  const int prefix = ep.prefix[b];
  const int kept = min(prefix, ep.max_cache);
  const int L = kept + T;                               // window length: kept frames + the new ones
  const int ring0 = (prefix - kept) % ep.cap;           // ring slot of window index 0 (frame f lives in slot f % cap)
  // keys any row of this tile can see: window indices [w_lo, w_hi)
  const int p_first = prefix + row0, p_last = prefix + min(row0 + 128, T) - 1;
  const int w_lo = max(0, p_first - ep.max_cache) - (prefix - kept);
  const int w_hi = min((p_last / ep.blocksize + 1) * ep.blocksize, prefix + T) - (prefix - kept);
  // tiles of 64 ring slots, starting at the 16-slot group that holds window index w_lo
  const int s_first = ring0 + w_lo;                     // un-wrapped ring position of the first needed key
  const int sb0 = (s_first >> 4) << 4;
  const int n_tiles = (ring0 + w_hi - sb0 + KT - 1) / KT;
  auto tile_wbase = [&](int t) { return sb0 + t * KT - ring0; };     // window index of column 0 of tile t (may be < w_lo)

  if (tid == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_k)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_v)) : "memory");
    for (int s = 0; s < NS; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1);
    mbar_init(p_bar, 128);
    mbar_init(pv_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                 "r"(static_cast<uint32_t>(kEaTmemCols))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // Q tile: 128 rows x 8 chunks of 16 B into the swizzled K-major layout; rows beyond T are zero
  {
    const int ldq = 3 * ep.H * HD;
    for (int u = tid; u < 128 * 8; u += kEaThreads) {
      const int rl = u >> 3, c = u & 7;
      const int r = row0 + rl;
      uint4 val = make_uint4(0u, 0u, 0u, 0u);
      if (r < T) val = *reinterpret_cast<const uint4*>(ep.qkv + static_cast<size_t>(b * T + r) * ldq + head * HD + c * 8);
      *reinterpret_cast<uint4*>(sQ + rl * 128 + ((c ^ (rl & 7)) << 4)) = val;
    }
    fence_proxy_async_smem();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 5) {
    // ================= TMA producer: a tile = 4 boxes of 16 ring slots for K and for V =================
    if (lane == 0) {
      const int row_base = (slot * ep.H + head) * ep.cap;
      const int groups = ep.cap >> 4;                   // 16-slot groups of the ring
      for (int u = 0; u < n_tiles; ++u) {
        const int stage = u % NS;
        if (u >= NS) mbar_wait(&empty_bar[stage], ((u / NS) - 1) & 1);
        uint8_t* dK = sStage + stage * kEaStageBytes;
        mbar_expect_tx(&full_bar[stage], kEaStageBytes);
        const int g0 = (sb0 + u * KT) >> 4;
#pragma unroll
        for (int pp = 0; pp < KT / 16; ++pp) {
          const int rr = row_base + ((g0 + pp) % groups) * 16;
          tma_load_2d(dK + pp * (16 * 128), &tm_k, &full_bar[stage], 0, ep.k_row0 + rr);
          tma_load_2d(dK + 64 * 128 + pp * (16 * 128), &tm_v, &full_bar[stage], 0, ep.v_row0 + rr);
        }
      }
    }
  } else if (warp == 4) {
    // ================= MMA issuer =================
    constexpr uint32_t idesc_s = make_idesc(128, KT);          // S = Q K^T : N = 64 keys
    constexpr uint32_t idesc_o = make_idesc_bmn(128, HD);      // O = P V   : N = 64 dims, B MN-major
    auto mma_s = [&](int t) {
      const int stage = t % NS;
      mbar_wait(&full_bar[stage], (t / NS) & 1);
      tcgen05_fence_after();
      if (lane == 0) {
        const uint32_t q = smem_u32(sQ);
        const uint32_t k = smem_u32(sStage + stage * kEaStageBytes);
        const uint32_t tS = tmem_base + (t & 1) * KT;
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk)                   // 4 k-steps of 16 dims
          umma_bf16(tS, make_smem_desc(q + kk * 32), make_smem_desc(k + kk * 32), idesc_s, kk > 0 ? 1u : 0u);
        umma_commit(&s_bar[t & 1]);
      }
      __syncwarp();
    };
    if (n_tiles > 0) mma_s(0);
    for (int t = 0; t < n_tiles; ++t) {
      if (t + 1 < n_tiles) mma_s(t + 1);                       // one tile ahead of the softmax (S is double-buffered)
      mbar_wait(p_bar, t & 1);
      tcgen05_fence_after();
      if (lane == 0) {
        const int stage = t % NS;
        const uint32_t pa = smem_u32(sP);
        const uint32_t v = smem_u32(sStage + stage * kEaStageBytes + 64 * 128);
        const uint32_t tO = tmem_base + 2 * KT;
#pragma unroll
        for (int kk = 0; kk < KT / 16; ++kk)                   // 4 k-steps of 16 keys
          umma_bf16(tO, make_smem_desc(pa + kk * 32), make_smem_desc_mn(v + kk * (16 * 128), 64 * 128, 1024), idesc_o,
                    (t > 0 || kk > 0) ? 1u : 0u);
        umma_commit(pv_bar);
        umma_commit(&empty_bar[stage]);
      }
      __syncwarp();
    }
  } else {
    // ================= softmax + output: thread = query row = TMEM lane =================
    const int rl = tid;                                        // 0..127
    const int r = row0 + rl;
    const bool live = r < T;
    const int p = prefix + r;                                  // absolute frame index of this row
    const int qlo = live ? max(0, p - ep.max_cache) - (prefix - kept) : 0;
    const int qhi = live ? min((p / ep.blocksize + 1) * ep.blocksize, prefix + T) - (prefix - kept) : 0;
    const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
    constexpr float sl2 = 1.4426950408889634f;
    float m_run = -INFINITY, l_run = 0.f;
    for (int t = 0; t < n_tiles; ++t) {
      const int jb = tile_wbase(t);
      mbar_wait(&s_bar[t & 1], (t >> 1) & 1);
      tcgen05_fence_after();
      uint32_t sr[4][16];
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) tmem_ld16_issue(tmem_base + lane_addr + (t & 1) * KT + q4 * 16, sr[q4]);
      tmem_ld_wait();
      float mx = -INFINITY;
      if (jb >= qlo && jb + KT <= qhi) {                       // interior tile for this row: no mask
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4)
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const float sv = __uint_as_float(sr[q4][e]) * sl2;
            sr[q4][e] = __float_as_uint(sv);
            mx = fmaxf(mx, sv);
          }
      } else {
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4)
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const int j = jb + q4 * 16 + e;
            const float sv = (j >= qlo && j < qhi) ? __uint_as_float(sr[q4][e]) * sl2 : -INFINITY;
            sr[q4][e] = __float_as_uint(sv);
            mx = fmaxf(mx, sv);
          }
      }
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
      fence_proxy_async_smem();
      tcgen05_fence_before();
      mbar_arrive(p_bar);
    }
    // ---- O = O / l ----
    mbar_wait(pv_bar, (n_tiles - 1) & 1);
    tcgen05_fence_after();
    const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
    bf16* dst = ep.out + static_cast<size_t>(b * T + r) * (ep.H * HD) + head * HD;
#pragma unroll 1
    for (int c8 = 0; c8 < HD / 32; ++c8) {
      uint32_t orr[2][16];
      tmem_ld16_issue(tmem_base + lane_addr + 2 * KT + c8 * 32, orr[0]);
      tmem_ld16_issue(tmem_base + lane_addr + 2 * KT + c8 * 32 + 16, orr[1]);
      tmem_ld_wait();
      if (live) {
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
                 "r"(static_cast<uint32_t>(kEaTmemCols))
                 : "memory");
  }
}

}  // namespace isst

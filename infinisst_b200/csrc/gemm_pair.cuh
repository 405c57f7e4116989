// CTA-pair (cta_group::2) tcgen05 GEMM for the tensor-bound regime:  out[tok, feat] = act[tok, K] . W[feat, K]^T
//
// Same call sites as gemm_tcgen05.cuh (SURVEY §2.3 L2/L7/L8 at prefill, E10-E14), taken when there are more token
// rows than one 128-row tile.  What bounds those GEMMs with one CTA per tile is the operand bytes an SM has to
// ingest per flop (48 KB per 128 x 256 x 64 k-block); here the two SMs of a TPC work on ONE tile:
//
//   UMMA M = 256 output features (this CTA stages its own 128 weight rows)
//   UMMA N = tile_tok tokens, any multiple of 16 up to 256 (this CTA stages its own HALF of the token rows;
//            the MMA reads the B operand from both CTAs' shared memory)
//
// so a k-block costs each SM 16 KB of weights + tile_tok / 2 x 128 B of activations for 128 x tile_tok outputs
// (31 KB instead of 48 KB at 240 tokens), and the accumulator (tile_tok columns) stays double-buffered in TMEM.
// Tokens are the N dimension because N is free in steps of 16: 1408 prefill rows are 6 tiles of 240 (2 % padding)
// where 256-row M tiles would pad 8 %.
//
// kDual (Llama gate/up): the 128 weight rows a CTA stages are 64 gate rows (f ..) followed by the 64 up rows of the same
// features (f + dual_off ..), so ONE accumulator holds gate in TMEM lanes 0-63 and up in lanes 64-127; the epilogue moves
// the up values through shared memory to the gate threads, which write silu(gate) * up.  A pair tile is then 128 output
// features x tile_tok tokens, with the same operand bytes per flop as the plain case.
//
// Warp roles per CTA: warp0 = TMA producer (both CTAs; every load signals the LEADER's full barrier), warp1 = MMA
// issuer (leader CTA only; commits are multicast to both CTAs' barriers), warp2 = TMEM allocator, warps 4-11 =
// epilogue (each CTA drains its own 128 TMEM lanes = its 128 features).  Work distribution (data-parallel rounds +
// stream-K remainder, fp32 partials reduced in contributor order) is gemm_sk_kernel's, with the PAIR as the unit.
#pragma once
#include "gemm_tcgen05.cuh"

namespace isst {
namespace tc {

template <bool kDual>
struct PairCfg {
  static constexpr int kCapN = 256;                        // token columns per tile (both CTAs together)
  static constexpr int kFeatTile = kDual ? 128 : 256;      // output features per pair tile
  static constexpr int kWBytes = kBM * kBK * 2;            // this CTA's 128 weight rows of one k-block (dual: 64 gate rows, then 64 up rows)
  static constexpr int kActBytes = (kCapN / 2) * kBK * 2;  // this CTA's half of the token rows
  static constexpr int kStageBytes = kWBytes + kActBytes;  // 32 KB
  static constexpr int kStages = 6;
  static constexpr int kXBytes = kDual ? 2 * 2 * 16 * 64 * 4 : 0;   // up-projection exchange [column half][buffer][16 columns][64 features] fp32
  static constexpr int kAccAll = kCapN;                    // TMEM columns of one accumulator
  static constexpr int kTmemCols = 512;                    // two accumulators
  static constexpr int kSmemBytes = kStages * kStageBytes + kXBytes + 1024 + 512;
  static constexpr int kEpiWarps = 8;
  static constexpr int kEpiThreads = kEpiWarps * 32;
  static constexpr int kThreadsTotal = 128 + kEpiThreads;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA loads of a CTA pair: data lands in THIS CTA's shared memory, the bytes are counted on `bar_cluster`
// (a shared::cluster address - the leader's full barrier)
__device__ __forceinline__ void tma2_load_2d(void* dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma2_load_4d(void* dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1,
                                             int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives (once the MMAs issued so far have completed) on the barrier at this shared-memory offset in BOTH CTAs
__device__ __forceinline__ void umma2_commit_both(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask)
               : "memory");
}

// Epilogue of 16 token columns of one output feature f: bias, SiLU gate, GELU, residual, store.
// kFast: all 16 columns and the feature are in range (warp-uniform), no predicates.
template <bool kDual, bool kFast>
__device__ __forceinline__ void pair_chunk16(const GemmParams& p, int f, int tok0, int n, bool f_ok, float bias,
                                             const float* v0, const float* v1, const float* rv) {
  float y[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    float x = v0[i] + bias;
    if (kDual) x = __fdividef(x, 1.0f + __expf(-x)) * v1[i];      // SiLU(gate) * up
    y[i] = x;
  }
  if (p.act == 1) {
#pragma unroll
    for (int i = 0; i < 16; ++i) y[i] = gelu_erf(y[i]);
  }
  const long long o0 = static_cast<long long>(tok0) * p.ldo + f;
  if (p.out_f32) {
    float* of = reinterpret_cast<float*>(p.out) + o0;
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (kFast || (f_ok && i < n)) of[static_cast<long long>(i) * p.ldo] = y[i] + rv[i];
  } else {
    bf16* ob = reinterpret_cast<bf16*>(p.out) + o0;
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (kFast || (f_ok && i < n)) ob[static_cast<long long>(i) * p.ldo] = __float2bfloat16_rn(y[i] + rv[i]);
  }
}
// residual values of the chunk (issued before the TMEM wait; `resid` may alias `out`)
template <bool kFast>
__device__ __forceinline__ void pair_resid16(const GemmParams& p, int f, int tok0, int n, bool f_ok, float* rv) {
  if (p.resid) {
    const bf16* r = p.resid + static_cast<long long>(tok0) * p.ldr + f;
#pragma unroll
    for (int i = 0; i < 16; ++i)
      rv[i] = (kFast || (f_ok && i < n)) ? __bfloat162float(r[static_cast<long long>(i) * p.ldr]) : 0.f;
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) rv[i] = 0.f;
  }
}

// sk: tiles_tok = ceil(M_tok / tile_tok), tiles_feat = ceil(N_out / kFeatTile), G = pairs = gridDim.x / 2.
// kDual: tm_w is the weight map with 64-row boxes.
template <bool kDual>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tm_act, const __grid_constant__ CUtensorMap tm_w,
                 const GemmParams p, const SkParams sk, const int tile_tok) {
  using C = PairCfg<kDual>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* xbuf = reinterpret_cast<float*>(smem + C::kStages * C::kStageBytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes + C::kXBytes);
  uint64_t* empty_bar = full_bar + C::kStages;
  uint64_t* tfull_bar = empty_bar + C::kStages;     // [2] accumulator ready (both CTAs get the arrive)
  uint64_t* tempty_bar = tfull_bar + 2;             // [2] accumulator drained by BOTH CTAs (the leader's copy is used)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int G = gridDim.x >> 1, pair = blockIdx.x >> 1;
  const int nkb = sk.num_kb;
  const int half_tok = tile_tok >> 1;               // token rows staged by each CTA
  if (sk.dbg && threadIdx.x == 0) sk.dbg[blockIdx.x * 8 + 0] = gtimer();
  pdl_launch_dependents();

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_w)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_act)) : "memory");
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 2 * C::kEpiWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                 "r"(static_cast<uint32_t>(C::kTmemCols))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  cluster_sync_all();           // both CTAs' barriers are initialised and their TMEM allocated before any remote signal
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ================= TMA producer (both CTAs) =================
    if (lane == 0) {
      const uint32_t stage_tx = static_cast<uint32_t>(C::kWBytes + half_tok * kBK * 2);
      int stage = 0;
      uint32_t phase = 0;
      SkWalker w(sk, G, pair);
      long long tile;
      int kb_begin, kb_end;
      int pre = 0;                                   // stages whose weights were requested ahead of pdl_wait()
      int pre_row[C::kStages], pre_k0[C::kStages];
      bool waited = false;
      if (sk.dbg) sk.dbg[blockIdx.x * 8 + 1] = gtimer();
      while (w.next(tile, kb_begin, kb_end)) {
        const SkTile t = sk_tile(tile, sk);
        const int act_row0 = t.tt * tile_tok + static_cast<int>(rank) * half_tok;
        // plain: this CTA's 128 rows of the 256-feature tile; dual: 64 gate rows + the 64 up rows of the same features
        const int w_row0 = kDual ? t.tf * 128 + static_cast<int>(rank) * 64 : t.tf * 256 + static_cast<int>(rank) * kBM;
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          if (!waited && pre == C::kStages) {
            pdl_wait();
            waited = true;
#pragma unroll 1
            for (int s2 = 0; s2 < pre; ++s2)
              tma2_load_4d(smem + s2 * C::kStageBytes + C::kWBytes, &tm_act, mapa_u32(smem_u32(&full_bar[s2]), 0),
                           pre_k0[s2], 0, pre_row[s2], 0);
          }
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sw = smem + stage * C::kStageBytes;
          uint8_t* sa = sw + C::kWBytes;
          const uint32_t bar = mapa_u32(smem_u32(&full_bar[stage]), 0);
          if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * stage_tx);
          const int k0 = kb * kBK;
          tma2_load_2d(sw, &tm_w, bar, k0, w_row0);
          if (kDual) tma2_load_2d(sw + C::kWBytes / 2, &tm_w, bar, k0, w_row0 + p.dual_off);
          if (waited) {
            tma2_load_4d(sa, &tm_act, bar, k0, 0, act_row0, 0);
          } else {
            pre_row[pre] = act_row0; pre_k0[pre] = k0;
            ++pre;
          }
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
      }
      if (!waited) {
        pdl_wait();
#pragma unroll 1
        for (int s2 = 0; s2 < pre; ++s2)
          tma2_load_4d(smem + s2 * C::kStageBytes + C::kWBytes, &tm_act, mapa_u32(smem_u32(&full_bar[s2]), 0),
                       pre_k0[s2], 0, pre_row[s2], 0);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (leader CTA only) =================
    if (rank == 0) {
      const uint32_t idesc = make_idesc(256, tile_tok);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      SkWalker w(sk, G, pair);
      long long tile;
      int kb_begin, kb_end;
      while (w.next(tile, kb_begin, kb_end)) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t tacc = tmem_base + acc * C::kAccAll;
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          if (lane == 0) {
            const uint32_t sw = smem_u32(smem + stage * C::kStageBytes);
            const uint64_t d_w = make_smem_desc(sw);
            const uint64_t d_act = make_smem_desc(sw + C::kWBytes);
#pragma unroll
            for (int k = 0; k < kBK / kUmmaK; ++k) {
              const uint64_t koff = static_cast<uint64_t>((k * kUmmaK * 2) >> 4);
              umma2_bf16(tacc, d_w + koff, d_act + koff, idesc, (kb > kb_begin || k > 0) ? 1u : 0u);
            }
            umma2_commit_both(&empty_bar[stage]);
            if (kb == kb_end - 1) umma2_commit_both(&tfull_bar[acc]);
          }
          __syncwarp();
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ================= epilogue: 4 TMEM lane quarters x 2 column halves (both CTAs, own features) =================
    pdl_wait();
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    const int r = q * 32 + lane;                                  // TMEM lane
    const int et = threadIdx.x - 128;
    const int slot_id = pair * 2 + static_cast<int>(rank);        // workspace slot / counter lane of this CTA
    const int n_slots = 2 * G;
    const int* counters = p.counters + (p.counter_parity ? p.counter_half : 0);
    for (int i = blockIdx.x * C::kEpiThreads + et; i < p.counter_half; i += gridDim.x * C::kEpiThreads)
      p.counters[(p.counter_parity ? 0 : p.counter_half) + i] = 0;
    // column range of this warp's half inside a tile: 16-column chunks [c_lo, c_hi)
    const int chunks = tile_tok >> 4;
    const int c_lo = (half ? (chunks + 1) / 2 : 0) * 16;
    const int c_hi = (half ? chunks : (chunks + 1) / 2) * 16;
    // dual: lanes 0-63 hold the gate rows, lanes 64-127 the up rows of the same 64 features; the up values travel
    // through shared memory to the warp that holds the gate ([16 columns][64 features], double-buffered per half)
    const bool holds_out = !kDual || q < 2;
    const int fl = kDual ? (r & 63) : r;                          // feature index inside this CTA's slice
    float* xb = xbuf + half * (2 * 16 * 64);
    int xsel = 0;
    constexpr size_t kSlot = static_cast<size_t>(C::kAccAll) * kBM;
    const uint32_t tempty_leader[2] = {mapa_u32(smem_u32(&tempty_bar[0]), 0), mapa_u32(smem_u32(&tempty_bar[1]), 0)};
    int acc = 0;
    uint32_t acc_phase = 0;
    SkWalker w(sk, G, pair);
    long long tile;
    int kb_begin, kb_end;
    long long part_tile[2];
    int n_part = 0;
    bool first_seg = true;
    while (w.next(tile, kb_begin, kb_end)) {
      const SkTile t = sk_tile(tile, sk);
      const int tok_base = t.tt * tile_tok;
      const int f = t.tf * C::kFeatTile + static_cast<int>(rank) * (C::kFeatTile / 2) + fl;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tcgen05_fence_after();
      if (first_seg && sk.dbg && et == 0) sk.dbg[blockIdx.x * 8 + 2] = gtimer();
      first_seg = false;
      const uint32_t taddr = tmem_base + acc * C::kAccAll + (static_cast<uint32_t>(q * 32) << 16);
      if (kb_begin == 0 && kb_end == nkb) {
        // ---- whole k-range accumulated here: final epilogue straight from TMEM ----
        const bool f_ok = f < p.N_out;
        const float bias = (p.bias && f_ok) ? p.bias[f] : 0.f;
        const bool feats_full = t.tf * C::kFeatTile + C::kFeatTile <= p.N_out;      // warp-uniform
#pragma unroll 1
        for (int c = c_lo; c < c_hi; c += 16) {
          uint32_t r0[16];
          float v0[16], v1[16], rv[16];
          tmem_ld16_issue(taddr + c, r0);
          const int n = min(16, p.M_tok - tok_base - c);
          const bool fast = feats_full && n == 16;
          if (holds_out) {
            if (fast) pair_resid16<true>(p, f, tok_base + c, n, f_ok, rv);
            else pair_resid16<false>(p, f, tok_base + c, n, f_ok, rv);
          }
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) v0[i] = __uint_as_float(r0[i]);
          if (kDual) {
            float* xw = xb + xsel * (16 * 64);
            if (q >= 2) {
#pragma unroll
              for (int i = 0; i < 16; ++i) xw[i * 64 + fl] = v0[i];
            }
            asm volatile("bar.sync %0, 128;" ::"r"(2 + half) : "memory");
            if (q < 2) {
#pragma unroll
              for (int i = 0; i < 16; ++i) v1[i] = xw[i * 64 + fl];
            }
            xsel ^= 1;
          }
          if (holds_out) {
            if (fast) pair_chunk16<kDual, true>(p, f, tok_base + c, n, f_ok, bias, v0, v1, rv);
            else if (n > 0) pair_chunk16<kDual, false>(p, f, tok_base + c, n, f_ok, bias, v0, v1, rv);
          }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(tempty_leader[acc]);
      } else {
        // ---- partial tile (stream-K): park the fp32 partial, announce it; reduced after the walk ----
        const long long ut = (tile - sk.tiles_dp) * nkb;
        const int c_first = sk_cta_of(ut, sk.units_sk, sk.g_sk);
        float* mine = p.ws + static_cast<size_t>(pair == c_first ? n_slots + slot_id : slot_id) * kSlot;
#pragma unroll 1
        for (int c = c_lo; c < c_hi; c += 16) {
          float v[16];
          tmem_ld16(taddr + c, v);
#pragma unroll
          for (int i = 0; i < 16; ++i) mine[(c + i) * kBM + r] = v[i];
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(tempty_leader[acc]);
        __threadfence();
        asm volatile("bar.sync 1, %0;" ::"n"(C::kEpiThreads) : "memory");
        if (et == 0) atomicAdd(const_cast<int*>(&counters[c_first * 2 + rank]), 1);
        part_tile[n_part++] = tile;
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (sk.dbg && et == 0) sk.dbg[blockIdx.x * 8 + 3] = gtimer();
    // ---- reduce the shared tiles: every contributor (same rank of each contributing pair) takes its share of the
    //      chunks; the parked partials are plain memory, so a gate thread reads the up lane (r + 64) itself ----
    for (int pi = 0; pi < n_part; ++pi) {
      const long long tl = part_tile[pi];
      const long long ut = (tl - sk.tiles_dp) * nkb;
      const int c_first = sk_cta_of(ut, sk.units_sk, sk.g_sk);
      const int c_last = sk_cta_of(ut + nkb - 1, sk.units_sk, sk.g_sk);
      const int nc = c_last - c_first + 1;
      if (et == 0) {
        while (ld_acquire(&counters[c_first * 2 + rank]) < nc) __nanosleep(32);
      }
      asm volatile("bar.sync 1, %0;" ::"n"(C::kEpiThreads) : "memory");
      const SkTile t = sk_tile(tl, sk);
      const int tok_base = t.tt * tile_tok;
      const int f = t.tf * C::kFeatTile + static_cast<int>(rank) * (C::kFeatTile / 2) + fl;
      const bool f_ok = f < p.N_out;
      const float bias = (p.bias && f_ok) ? p.bias[f] : 0.f;
      const int workers = nc * 2;
      const int me = (pair - c_first) * 2 + half;
      const int ch0 = chunks * me / workers, ch1 = chunks * (me + 1) / workers;
      if (holds_out) {
#pragma unroll 1
        for (int ch = ch0; ch < ch1; ++ch) {
          const int c = ch * 16;
          const int n = min(16, p.M_tok - tok_base - c);
          float v0[16], v1[16], rv[16];
          pair_resid16<false>(p, f, tok_base + c, n, f_ok, rv);
#pragma unroll
          for (int i = 0; i < 16; ++i) { v0[i] = 0.f; v1[i] = 0.f; }
#pragma unroll 2
          for (int cc = c_first; cc <= c_last; ++cc) {
            const int sid = cc * 2 + static_cast<int>(rank);
            const float* src = p.ws + static_cast<size_t>(cc == c_first ? n_slots + sid : sid) * kSlot;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              v0[i] += __ldcg(&src[(c + i) * kBM + (kDual ? fl : r)]);
              if (kDual) v1[i] += __ldcg(&src[(c + i) * kBM + 64 + fl]);
            }
          }
          if (n > 0) pair_chunk16<kDual, false>(p, f, tok_base + c, n, f_ok, bias, v0, v1, rv);
        }
      }
    }
    if (sk.dbg && et == 0) sk.dbg[blockIdx.x * 8 + 5] = gtimer();
    tcgen05_fence_before();
  }
  // neither CTA may leave (or free its TMEM) while the other can still read its shared memory / write its TMEM
  __syncwarp();
  cluster_sync_all();
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(static_cast<uint32_t>(C::kTmemCols))
                 : "memory");
  }
  if (sk.dbg && threadIdx.x == 0) sk.dbg[blockIdx.x * 8 + 4] = gtimer();
}

}  // namespace tc
}  // namespace isst

// Fused decode-layer chain: ONE persistent kernel runs a sequence of weight-streaming GEMMs and RMSNorm row phases
// of the Llama decode step (<= 64 token rows), separated by in-kernel grid barriers:
//
//     o_proj -> [+residual, RMSNorm] -> gate/up (SiLU * up) -> down_proj -> [+residual, RMSNorm] -> QKV of the next layer
//     (last layer: ... -> down_proj -> [+residual, final RMSNorm on the last rows] -> lm_head)
//
// Replaces, per decode layer, the o_proj / gate / up / down / q / k / v nn.Linear calls and the two LlamaRMSNorm
// modules of the reference (HF LlamaDecoderLayer as patched by model/patches/patch_llm.py:231-336; SURVEY §2.3 L1, L2,
// L7, L8) - seven launches of the operator-per-kernel path - with one launch; decode attention stays its own kernel
// between two chains (it consumes the QKV split partials the chain leaves and produces the o_proj input).
//
// Why: at 64 rows every GEMM is pure weight streaming (HBM-bound).  With one kernel per operator each launch pays
// start-up (CTA launch, barrier init, TMEM allocation), accumulator drain and grid turnover of a 224 KB one-CTA-per-SM
// kernel: ~7 us of the ~25-50 us a GEMM takes.  Here the CTAs stay resident, the TMA producer walks straight from the
// last weight tile of one phase into the first weight tiles of the next (weights never depend on a previous phase), and
// only the ACTIVATION loads of a phase wait for the grid barrier that closes the phase before it - the shared-memory
// ring (5 x 40 KB per SM) keeps requests in flight across the barrier.
//
// Structure (384 threads, one CTA per SM, G = #SMs CTAs):
//   warp 0      TMA producer: per unit (TWO 128-feature tiles x k-range) and k-block one stage = [act tile BN x 64]
//               [weight tile A 128 x 64][weight tile B 128 x 64], all K-major SWIZZLE_128B.  The two tiles are gate / up
//               rows of the same features (SwiGLU) or two adjacent feature tiles: one activation tile serves 32 KB of
//               weights - an SM ingests ~70 KB/us from L2/HBM whatever the mix (measured: 47 KB/us of weights with one
//               8 KB activation tile per 16 KB weight tile), so the activation share decides how fast weights stream
//   warp 1      tcgen05.mma issuer: two accumulators D[128 features x BN tokens], fp32 in TMEM, double-buffered so
//               the epilogue of a unit overlaps the MMAs of the next
//   warp 2      TMEM allocator
//   warps 4-11  epilogue: fp32 split-K partials / SiLU(gate) * up in bf16 / fp32 logits; the RMSNorm row phases
//               (sum of the split partials + residual -> residual stream, normalise, scale) - one row per CTA
// Grid barrier: one monotonically increasing 64-bit counter PER PHASE INDEX in global memory (a CTA without work in a
// phase runs ahead and arrives early for later phases - with a single counter those early arrivals would be mistaken for
// the missing arrivals of the phase being waited for); every CTA arrives once per phase, waiters poll with
// ld.acquire.gpu; readers of another CTA's generic-proxy stores through TMA add fence.proxy.async.
// All CTAs are co-resident (grid <= #SMs, 1 CTA / SM), so spinning cannot deadlock; under programmatic dependent launch
// the kernel may start while its predecessor drains (its weight prefetch then overlaps the predecessor's tail).
// The k-ranges, accumulation order and the order in which the row phase sums the partials are those of the
// operator-per-kernel path (gemm_sk_kernel deferred mode + norm_rows_kernel), so both paths give identical bits.
#pragma once
#include "gemm_tcgen05.cuh"

namespace isst {
namespace chain {

using namespace tc;

constexpr int kMaxPhases = 8;
constexpr int kMaxMaps = 4;
constexpr int kMaxTiles = 64;          // tile pairs of a fold_reduce GEMM (n_out <= 16384)
enum { PH_GEMM = 0, PH_ROWS = 1 };
enum { EPI_PART = 0, EPI_SILU = 1, EPI_F32 = 2 };

struct Phase {
  int kind;
  // ---- PH_GEMM: out[tok, f] = act[tok, :] . W[f, :] over feature tiles of 128 rows ----
  int wmap, amap;            // tensor map indices
  int tiles;                 // units of TWO weight tiles: rows [tile * tile_rows, +128) and the same + sub_off
  int tile_rows;             // 128 (gate/up: second tile = `up` rows of the same features) or 256 (adjacent feature tiles)
  int sub_off;               // row offset of the second tile: n_out (gate/up) or 128
  int num_kb;                // K / 64
  int splits;                // k-ranges per unit tile; units = tiles * splits
  int dual;                  // 1: W holds [gate; up] -> silu(gate) * up
  int epi;                   // EPI_*
  int n_out;
  void* out;                 // EPI_PART: float [splits][n_tok][n_out]; EPI_SILU: bf16 [n_tok][n_out]; EPI_F32: float [n_tok][n_out]
  // ---- PH_ROWS: x = bf16(x_in[src] + sum_s part[s][src]); x_out[src] = x; h_out[row] = w * bf16(x * rstd) ----
  const bf16* x_in;
  bf16* x_out;               // may be null
  bf16* h_out;
  const float* w;
  const float* part;         // may be null (n_part = 0)
  int n_part;
  long long part_stride;
  const int* gather;         // src row of row r, or null (identity)
  int n_rows;
  int C;
  float eps;
  // ---- folded RMSNorm (decode): the row phase between two GEMMs and its grid barrier disappear ----
  // PH_GEMM, EPI_PART with fold_reduce: once every k-split of a tile pair has parked its partial (a counter per tile), the
  //   split CTAs share the tile's tokens: x = bf16(x_in + sum_s part[s]) -> x_out, h_out = bf16(x * w) (NOT normalised),
  //   sq_out[tile][tok] = sum of x^2 over the tile's 256 features.  Uses x_in / x_out / h_out / w above.
  // PH_GEMM with sq_in: the accumulators are multiplied by rstd[tok] = rsqrt(sum_t sq_in[t][tok] / C + eps) in the epilogue
  //   (the GEMM is linear in its activation rows, so normalising after it is the same sum).
  int fold_reduce;
  float* sq_out;
  const float* sq_in;
  int n_sq;                  // tiles summed per token
};

struct Params {
  CUtensorMap wmaps[kMaxMaps];
  CUtensorMap amaps[kMaxMaps];
  Phase ph[kMaxPhases];
  int n_phases;
  int n_wmaps, n_amaps;
  int n_tok;
  unsigned long long* dbg;                    // optional [grid][32] %globaltimer stamps (phase analysis), may be null
  unsigned long long* bar;                    // [kMaxPhases] grid barrier counters, one per phase index
  unsigned long long bar_base[kMaxPhases];    // their values when this launch starts
  int* tile_ctr;                              // fold_reduce arrival counters [2 halves][kMaxPhases * kMaxTiles]: alternate
  int ctr_parity;                             //   launches use alternate halves; a launch re-arms the other half
};

template <int kBN>
struct Cfg {
  static constexpr int kActBytes = kBN * kBK * 2;
  static constexpr int kWBytes = kBM * kBK * 2;
  static constexpr int kStageBytes = kActBytes + 2 * kWBytes;
  static constexpr int kStagesRaw = (222 * 1024) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 12 ? 12 : kStagesRaw;
  static constexpr int kAccCols = 2 * kBN;                    // two weight tiles: gate | up, or two adjacent feature tiles
  static constexpr int kAccBufs = (2 * kAccCols <= 512) ? 2 : 1;   // double-buffered up to 128 token columns
  static constexpr int kTmemColsRaw = kAccBufs * kAccCols;
  static constexpr int kTmemCols = kTmemColsRaw <= 32 ? 32 : (kTmemColsRaw <= 64 ? 64 : (kTmemColsRaw <= 128 ? 128 : (kTmemColsRaw <= 256 ? 256 : 512)));
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 512 + 1024;   // + rstd[256]
  static constexpr int kEpiThreads = 256;
  static constexpr int kThreads = 128 + kEpiThreads;
  static constexpr int kEpiHalves = kBN >= 32 ? 2 : 1;
  static constexpr int kHalfCols = kBN / kEpiHalves;
  static_assert(kBN == 16 || kBN == 32 || kBN == 64 || kBN == 128 || kBN == 256, "token tile");
  static_assert(kStages >= 3, "ring too shallow");
  static_assert(kActBytes % 1024 == 0 && kWBytes % 1024 == 0, "tiles must keep 1024B alignment");
};

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void grid_wait(const Params& p, int phase_done, int G) {      // one thread
  const unsigned long long target = p.bar_base[phase_done] + static_cast<unsigned long long>(G);
  while (ld_acquire_u64(p.bar + phase_done) < target) __nanosleep(20);
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

// unit u of a GEMM phase -> (tile, k-range); CTA c owns units c, c + G, ...
__device__ __forceinline__ void unit_range(const Phase& ph, int u, int& tile, int& split, int& kb0, int& kb1) {
  tile = u / ph.splits;
  split = u - tile * ph.splits;
  kb0 = static_cast<int>(static_cast<long long>(ph.num_kb) * split / ph.splits);
  kb1 = static_cast<int>(static_cast<long long>(ph.num_kb) * (split + 1) / ph.splits);
}

// fold_reduce of one (tile pair, k-split) unit: this CTA's share of the tile's tokens (see Phase).  Kept out of line: the
// epilogue around the call site is at the register limit of a 384-thread CTA.
template <int kEpiThreads>
__device__ __noinline__ void fold_reduce_tokens(const Phase& ph, int n_tok, int tile, int split, int et) {
  const float* part = reinterpret_cast<const float*>(ph.out);
  const int lane = et & 31;
  const int t0 = n_tok * split / ph.splits, t1 = n_tok * (split + 1) / ph.splits;
  // one warp per token, a lane owns 8 consecutive features of the tile pair's 256 (16-byte loads), the squared sum is
  // one warp reduction - no block barrier
  const int f0 = tile * ph.tile_rows + lane * 8;
  const bool f_ok = f0 + 8 <= ph.n_out;                      // n_out is a multiple of 8 (checked on the host)
  const size_t pstride = static_cast<size_t>(n_tok) * ph.n_out;
  float wv[8];
  {
    const float4 w0 = f_ok ? __ldg(reinterpret_cast<const float4*>(ph.w + f0)) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 w1 = f_ok ? __ldg(reinterpret_cast<const float4*>(ph.w + f0 + 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
    wv[0] = w0.x; wv[1] = w0.y; wv[2] = w0.z; wv[3] = w0.w; wv[4] = w1.x; wv[5] = w1.y; wv[6] = w1.z; wv[7] = w1.w;
  }
#pragma unroll 1
  for (int t = t0 + (et >> 5); t < t1; t += kEpiThreads / 32) {
    const size_t o = static_cast<size_t>(t) * ph.n_out + f0;
    float a8[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a8[j] = 0.f;
    float sq = 0.f;
    if (f_ok) {
      const uint4 raw = *reinterpret_cast<const uint4*>(ph.x_in + o);
#pragma unroll 1
      for (int sp0 = 0; sp0 < ph.splits; sp0 += 4) {         // four splits' loads in flight (eight measured slower
        float4 pa[4], pb[4];                                 // at 64 rows: 77.1 vs 75.9 ms per step); split order
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          if (sp0 + q4 < ph.splits) {
            const float4* pp = reinterpret_cast<const float4*>(part + (sp0 + q4) * pstride + o);
            pa[q4] = __ldcg(pp); pb[q4] = __ldcg(pp + 1);
          }
        }
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          if (sp0 + q4 < ph.splits) {
            a8[0] += pa[q4].x; a8[1] += pa[q4].y; a8[2] += pa[q4].z; a8[3] += pa[q4].w;
            a8[4] += pb[q4].x; a8[5] += pb[q4].y; a8[6] += pb[q4].z; a8[7] += pb[q4].w;
          }
        }
      }
      const uint32_t ru[4] = {raw.x, raw.y, raw.z, raw.w};
      uint32_t nu[4], hu[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f2 = unpack_bf16(ru[j]);
        const float x0 = bf16_round(a8[2 * j] + f2.x), x1 = bf16_round(a8[2 * j + 1] + f2.y);
        nu[j] = pack_bf16(x0, x1);
        hu[j] = pack_bf16(x0 * wv[2 * j], x1 * wv[2 * j + 1]);
        sq += x0 * x0 + x1 * x1;
      }
      if (ph.x_out) *reinterpret_cast<uint4*>(ph.x_out + o) = make_uint4(nu[0], nu[1], nu[2], nu[3]);
      *reinterpret_cast<uint4*>(ph.h_out + o) = make_uint4(hu[0], hu[1], hu[2], hu[3]);
    }
    sq = warp_sum(sq);
    if (lane == 0) ph.sq_out[static_cast<size_t>(tile) * n_tok + t] = sq;
  }
}

// kFold = false: GEMM phases + RMSNorm row phases (the head chain, chunk-prefill, `chain_fold` = 0).
// kFold = true : GEMM phases with the folded RMSNorm (fold_reduce / sq_in) and NO row-phase code - the layer chains of a
//                decode forward.  Two instantiations keep either variant inside the register budget of a 384-thread CTA.
template <int kBN, bool kFold>
__global__ void __launch_bounds__(384, 1)
decode_chain_kernel(const __grid_constant__ Params p) {
  using C = Cfg<kBN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes);
  uint64_t* empty_bar = full_bar + C::kStages;
  uint64_t* tfull_bar = empty_bar + C::kStages;     // [2] accumulator ready
  uint64_t* tempty_bar = tfull_bar + 2;             // [2] accumulator drained
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* red = reinterpret_cast<float*>(tmem_ptr_smem + 2);   // [16] row-phase reduction scratch (+ [1] result)
  volatile int* rows_done = reinterpret_cast<volatile int*>(red + 20);   // 1 + index of the last row phase this CTA finished
  float* rstd_s = reinterpret_cast<float*>(smem + C::kStages * C::kStageBytes + 512);   // [256] per-token 1 / rms (sq_in phases)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int G = gridDim.x, cta = blockIdx.x;
  pdl_launch_dependents();
  if (p.dbg && threadIdx.x == 0) p.dbg[cta * 32] = gtimer();

  uint32_t tmem_base = 0;
  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < p.n_wmaps; ++i)
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.wmaps[i])) : "memory");
      for (int i = 0; i < p.n_amaps; ++i)
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.amaps[i])) : "memory");
      for (int s = 0; s < C::kStages; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(&tfull_bar[a], 1);
        mbar_init(&tempty_bar[a], C::kEpiThreads);
      }
      *rows_done = 0;
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("bar.arrive 2, %0;" ::"n"(C::kThreads) : "memory");
  } else {
    if (warp == 2) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                   "r"(static_cast<uint32_t>(C::kTmemCols))
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    asm volatile("bar.sync 2, %0;" ::"n"(C::kThreads) : "memory");
    tcgen05_fence_after();
    tmem_base = *tmem_ptr_smem;
  }

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      bool first_dep = true;                          // nothing of the predecessor KERNEL has been waited for yet
      for (int pi = 0; pi < p.n_phases; ++pi) {
        const Phase& ph = p.ph[pi];
        if (ph.kind != PH_GEMM) continue;
        const CUtensorMap* wm = &p.wmaps[ph.wmap];
        const CUtensorMap* am = &p.amaps[ph.amap];
        const int units = ph.tiles * ph.splits;
        // Weights do not depend on earlier phases: weight tiles are requested as soon as ring stages free up; the
        // activation tiles of those stages follow once the phase before this one is complete everywhere.
        // (An L2 look-ahead beyond the ring - cp.async.bulk.prefetch.tensor for the next 8-64 tiles while waiting at
        // the barrier - was measured: no gain, 64 tiles slower; phases are bound by SM ingest, not by HBM idling.)
        bool dep_ok = false;
        int n_pend = 0;
        int pend_stage[C::kStages], pend_k0[C::kStages];
        // a CTA that owns rows of the row phase before this one leaves the SM's ingest path to them: refilling the ring
        // with 160 KB of weights at the same time stretched the slowest rows from ~2.5 to ~6 us (the phase waits for them)
        if (pi > 0 && p.ph[pi - 1].kind == PH_ROWS && cta < p.ph[pi - 1].n_rows)
          while (*rows_done < pi) __nanosleep(32);
        auto resolve = [&]() {
          if (pi == 0 || first_dep) pdl_wait();
          first_dep = false;
          if (pi > 0) grid_wait(p, pi - 1, G);
          fence_proxy_async_all();                    // other CTAs' generic-proxy stores -> visible to TMA reads
          for (int i = 0; i < n_pend; ++i)
            tma_load_4d(smem + pend_stage[i] * C::kStageBytes, am, &full_bar[pend_stage[i]], pend_k0[i], 0, 0, 0);
          n_pend = 0;
          dep_ok = true;
          if (p.dbg) p.dbg[cta * 32 + 1 + 3 * pi] = gtimer();          // dependency of this phase resolved, activations requested
        };
        for (int u = cta; u < units; u += G) {
          int tile, split, kb0, kb1;
          unit_range(ph, u, tile, split, kb0, kb1);
          const int w_row0 = tile * ph.tile_rows;
          for (int kb = kb0; kb < kb1; ++kb) {
            if (!dep_ok && n_pend == C::kStages) resolve();
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * C::kStageBytes;
            mbar_expect_tx(&full_bar[stage], C::kStageBytes);
            tma_load_2d(sa + C::kActBytes, wm, &full_bar[stage], kb * kBK, w_row0);
            tma_load_2d(sa + C::kActBytes + C::kWBytes, wm, &full_bar[stage], kb * kBK, w_row0 + ph.sub_off);
            if (dep_ok) {
              tma_load_4d(sa, am, &full_bar[stage], kb * kBK, 0, 0, 0);
            } else {
              pend_stage[n_pend] = stage; pend_k0[n_pend] = kb * kBK;
              ++n_pend;
            }
            if (++stage == C::kStages) { stage = 0; phase ^= 1; }
          }
        }
        if (!dep_ok && n_pend > 0) resolve();
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    constexpr uint32_t idesc = make_idesc(kBM, kBN);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int pi = 0; pi < p.n_phases; ++pi) {
      const Phase& ph = p.ph[pi];
      if (ph.kind != PH_GEMM) continue;
      const int units = ph.tiles * ph.splits;
      for (int u = cta; u < units; u += G) {
        int tile, split, kb0, kb1;
        unit_range(ph, u, tile, split, kb0, kb1);
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t tacc = tmem_base + acc * C::kAccCols;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          if (lane == 0) {
            const uint32_t sa = smem_u32(smem + stage * C::kStageBytes);
            const uint64_t d_act = make_smem_desc(sa);
            const uint64_t d_w0 = make_smem_desc(sa + C::kActBytes);
            const uint64_t d_w1 = make_smem_desc(sa + C::kActBytes + C::kWBytes);
#pragma unroll
            for (int k = 0; k < kBK / kUmmaK; ++k) {
              const uint64_t koff = static_cast<uint64_t>((k * kUmmaK * 2) >> 4);
              const uint32_t accum = (kb > kb0 || k > 0) ? 1u : 0u;
              umma_bf16(tacc, d_w0 + koff, d_act + koff, idesc, accum);
              umma_bf16(tacc + kBN, d_w1 + koff, d_act + koff, idesc, accum);
            }
            umma_commit(&empty_bar[stage]);
            if (kb == kb1 - 1) umma_commit(&tfull_bar[acc]);
          }
          __syncwarp();
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
        if (++acc == C::kAccBufs) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ================= epilogue + row phases =================
    pdl_wait();                                   // every buffer written below belongs to earlier kernels until now
    const int q = warp & 3;                       // TMEM lane quarter accessible to this warp
    const int half = (warp - 4) >> 2;             // token-column half handled by this warp
    const int r = q * 32 + lane;                  // TMEM lane == feature row of the tile
    const int et = threadIdx.x - 128;             // 0..255
    const bool works = half < C::kEpiHalves;
    const int hc0 = half * C::kHalfCols;
    int* ctr = p.tile_ctr + (p.ctr_parity ? kMaxPhases * kMaxTiles : 0);
    if (kFold) {
      for (int i = cta * C::kEpiThreads + et; i < kMaxPhases * kMaxTiles; i += G * C::kEpiThreads)
        p.tile_ctr[(p.ctr_parity ? 0 : kMaxPhases * kMaxTiles) + i] = 0;     // the next folded launch's half (its last user is complete)
    }
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int pi = 0; pi < p.n_phases; ++pi) {
      const Phase& ph = p.ph[pi];
      if (ph.kind == PH_GEMM) {
        const int units = ph.tiles * ph.splits;
        const bool scaled = kFold && ph.sq_in != nullptr;
        if (scaled) {
          // 1 / rms of every token row from the squared sums the phase before left per feature tile
          if (et == 0) grid_wait(p, pi - 1, G);
          asm volatile("bar.sync 1, %0;" ::"n"(C::kEpiThreads) : "memory");
          if (et < p.n_tok) {
            float qs = 0.f;
            for (int t = 0; t < ph.n_sq; ++t) qs += __ldcg(ph.sq_in + static_cast<size_t>(t) * p.n_tok + et);
            rstd_s[et] = rsqrtf(qs / ph.C + ph.eps);
          }
          asm volatile("bar.sync 1, %0;" ::"n"(C::kEpiThreads) : "memory");
        }
        for (int u = cta; u < units; u += G) {
          int tile, split, kb0, kb1;
          unit_range(ph, u, tile, split, kb0, kb1);
          mbar_wait(&tfull_bar[acc], acc_phase);
          tcgen05_fence_after();
          if (p.dbg && et == 0 && u == cta) p.dbg[cta * 32 + 2 + 3 * pi] = gtimer();   // first accumulator of the phase complete
          const uint32_t taddr = tmem_base + acc * C::kAccCols + (static_cast<uint32_t>(q * 32) << 16);
          const int n_tok = p.n_tok - hc0;         // valid token columns from hc0 on
          // tcgen05.ld is warp-collective (.sync.aligned): the condition around it must be warp-uniform (a partial
          // last tile, e.g. the 7 valid rows of lm_head's tile 1002, guards the STORES per lane instead)
          if (ph.epi == EPI_SILU) {
            const int f = tile * kBM + r;
            if (works && tile * kBM + q * 32 < ph.n_out) {
              const bool row_ok = f < ph.n_out;
              bf16* dst = reinterpret_cast<bf16*>(ph.out) + static_cast<size_t>(hc0) * ph.n_out + f;
              // 32 token columns per step: the four TMEM loads of a step are in flight together (one wait)
              constexpr int kStep = C::kHalfCols >= 32 ? 32 : 16;
#pragma unroll 1
              for (int c = 0; c < C::kHalfCols; c += kStep) {
                uint32_t g[kStep], up[kStep];
#pragma unroll
                for (int j = 0; j < kStep; j += 16) {
                  tmem_ld16_issue(taddr + hc0 + c + j, g + j);
                  tmem_ld16_issue(taddr + kBN + hc0 + c + j, up + j);
                }
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < kStep; ++i) {
                  const float rs = scaled ? rstd_s[hc0 + c + i] : 1.0f;
                  const float gv = __uint_as_float(g[i]) * rs;
                  const float x = __fdividef(gv, 1.0f + __expf(-gv)) * (__uint_as_float(up[i]) * rs);
                  if (row_ok && c + i < n_tok) dst[static_cast<size_t>(c + i) * ph.n_out] = __float2bfloat16_rn(x);
                }
              }
            }
          } else {
            // EPI_PART: fp32 partial of this k-range [split][tok][feature]; EPI_F32: fp32 logits [tok][feature];
            // the two accumulators are two adjacent feature tiles
            // both tiles' columns of a step are loaded from TMEM together (one wait), then stored
            const int fa = tile * ph.tile_rows, fb = fa + ph.sub_off;
            const bool warp_a = works && fa + q * 32 < ph.n_out, warp_b = works && fb + q * 32 < ph.n_out;
            if (warp_a) {
              const bool ok_a = fa + r < ph.n_out, ok_b = fb + r < ph.n_out;
              float* dst_a = reinterpret_cast<float*>(ph.out) +
                             (static_cast<size_t>(ph.epi == EPI_PART ? split : 0) * p.n_tok + hc0) * ph.n_out + fa + r;
              float* dst_b = dst_a + ph.sub_off;
              constexpr int kStep = C::kHalfCols >= 32 ? 32 : 16;
#pragma unroll 1
              for (int c = 0; c < C::kHalfCols; c += kStep) {
                uint32_t va[kStep], vb[kStep];
#pragma unroll
                for (int j = 0; j < kStep; j += 16) {
                  tmem_ld16_issue(taddr + hc0 + c + j, va + j);
                  if (warp_b) tmem_ld16_issue(taddr + kBN + hc0 + c + j, vb + j);
                }
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < kStep; ++i)
                  if (ok_a && c + i < n_tok)
                    dst_a[static_cast<size_t>(c + i) * ph.n_out] = scaled ? __uint_as_float(va[i]) * rstd_s[hc0 + c + i] : __uint_as_float(va[i]);
                if (warp_b) {
#pragma unroll
                  for (int i = 0; i < kStep; ++i)
                    if (ok_b && c + i < n_tok)
                      dst_b[static_cast<size_t>(c + i) * ph.n_out] = scaled ? __uint_as_float(vb[i]) * rstd_s[hc0 + c + i] : __uint_as_float(vb[i]);
                }
              }
            }
          }
          tcgen05_fence_before();
          mbar_arrive(&tempty_bar[acc]);
          if (++acc == C::kAccBufs) { acc = 0; acc_phase ^= 1; }
        }
        if (kFold && ph.fold_reduce) {
          // ---- folded residual + RMSNorm statistics: the k-split CTAs of a tile pair share its tokens ----
          for (int u = cta; u < units; u += G) {
            int tile, split, kb0, kb1;
            unit_range(ph, u, tile, split, kb0, kb1);
            __threadfence();
            asm volatile("bar.sync 1, %0;" ::"n"(C::kEpiThreads) : "memory");      // this CTA's partials are out
            if (et == 0) {
              atomicAdd(&ctr[pi * kMaxTiles + tile], 1);
              while (ld_acquire(&ctr[pi * kMaxTiles + tile]) < ph.splits) __nanosleep(20);
            }
            asm volatile("bar.sync 1, %0;" ::"n"(C::kEpiThreads) : "memory");
            fold_reduce_tokens<C::kEpiThreads>(ph, p.n_tok, tile, split, et);
          }
        }
      } else if (!kFold) {
        // ---- row phase: one row per CTA (rows cta, cta + G, ...) ----
        if (pi > 0) {
          if (et == 0) grid_wait(p, pi - 1, G);
          asm volatile("bar.sync 1, %0;" ::"n"(C::kEpiThreads) : "memory");
        }
        if (p.dbg && et == 0) p.dbg[cta * 32 + 2 + 3 * pi] = gtimer();                 // row phase may start
        const int Cc = ph.C;
        const int groups = Cc >> 3;                // 16-byte column groups; thread et owns groups et and et + 256
        const int nvw = (groups + 31) >> 5;        // "virtual warps" of the one-group-per-thread layout (norm_rows_kernel)
        for (int row = cta; row < ph.n_rows; row += G) {
          const size_t src = ph.gather ? static_cast<size_t>(ph.gather[row]) : static_cast<size_t>(row);
          const bf16* x = ph.x_in + src * Cc;
          float v[2][8];
          float4 wq[2][2];                           // norm weights: independent of the row, fetched ahead of the reduction
          float sq[2] = {0.f, 0.f};
#pragma unroll
          for (int it = 0; it < 2; ++it) {
            const int gi = it * C::kEpiThreads + et;
            if (gi < groups) {
              wq[it][0] = __ldg(reinterpret_cast<const float4*>(ph.w + gi * 8));
              wq[it][1] = __ldg(reinterpret_cast<const float4*>(ph.w + gi * 8 + 4));
            }
          }
#pragma unroll
          for (int it = 0; it < 2; ++it) {
            const int gi = it * C::kEpiThreads + et;
            if (gi < groups) {
              const int c0 = gi * 8;
              uint4 raw = *reinterpret_cast<const uint4*>(x + c0);
              if (ph.n_part > 0) {
                float a8[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) a8[j] = 0.f;
                const float* p0 = ph.part + src * Cc + c0;
                float4 pa[8], pb[8];
#pragma unroll
                for (int sp = 0; sp < 8; ++sp) {
                  if (sp < ph.n_part) {
                    const float4* pp = reinterpret_cast<const float4*>(p0 + sp * ph.part_stride);
                    pa[sp] = __ldcg(pp); pb[sp] = __ldcg(pp + 1);
                  }
                }
#pragma unroll
                for (int sp = 0; sp < 8; ++sp) {
                  if (sp < ph.n_part) {
                    a8[0] += pa[sp].x; a8[1] += pa[sp].y; a8[2] += pa[sp].z; a8[3] += pa[sp].w;
                    a8[4] += pb[sp].x; a8[5] += pb[sp].y; a8[6] += pb[sp].z; a8[7] += pb[sp].w;
                  }
                }
                const uint32_t ru[4] = {raw.x, raw.y, raw.z, raw.w};
                uint32_t nu[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float2 f2 = unpack_bf16(ru[j]);
                  nu[j] = pack_bf16(a8[2 * j] + f2.x, a8[2 * j + 1] + f2.y);
                }
                raw = make_uint4(nu[0], nu[1], nu[2], nu[3]);
                if (ph.x_out) *reinterpret_cast<uint4*>(ph.x_out + src * Cc + c0) = raw;
              }
              const uint32_t uu[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 f2 = unpack_bf16(uu[j]);
                v[it][2 * j] = f2.x; v[it][2 * j + 1] = f2.y;
                sq[it] += f2.x * f2.x + f2.y * f2.y;
              }
            }
          }
          // reduction in the order of norm_rows_kernel<.., 1>: warp tree per 32 groups, then the warps in sequence
          const float s0 = warp_sum(sq[0]), s1 = warp_sum(sq[1]);
          if (lane == 0) { red[et >> 5] = s0; red[8 + (et >> 5)] = s1; }
          asm volatile("bar.sync 1, %0;" ::"n"(C::kEpiThreads) : "memory");
          if (et == 0) {
            float qsum = 0.f;
            for (int i = 0; i < nvw; ++i) qsum += red[i];
            red[16] = rsqrtf(qsum / Cc + ph.eps);
          }
          asm volatile("bar.sync 1, %0;" ::"n"(C::kEpiThreads) : "memory");
          const float rstd = red[16];
          bf16* o = ph.h_out + static_cast<size_t>(row) * Cc;
#pragma unroll
          for (int it = 0; it < 2; ++it) {
            const int gi = it * C::kEpiThreads + et;
            if (gi < groups) {
              const int c0 = gi * 8;
              const float4 w0 = wq[it][0], w1 = wq[it][1];
              const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
              uint32_t pk[4];
#pragma unroll
              for (int j = 0; j < 4; ++j)
                pk[j] = pack_bf16(wv[2 * j] * bf16_round(v[it][2 * j] * rstd), wv[2 * j + 1] * bf16_round(v[it][2 * j + 1] * rstd));
              *reinterpret_cast<uint4*>(o + c0) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
          }
          asm volatile("bar.sync 1, %0;" ::"n"(C::kEpiThreads) : "memory");      // `red` is reused by the next row
        }
        if (et == 0) *rows_done = pi + 1;
      }
      // ---- this CTA is done with phase pi ----
      if (p.dbg && et == 0) p.dbg[cta * 32 + 3 + 3 * pi] = gtimer();                   // this CTA's part of the phase is done
      if (pi + 1 < p.n_phases) {
        asm volatile("bar.sync 1, %0;" ::"n"(C::kEpiThreads) : "memory");
        if (et == 0) {
          __threadfence();
          atomicAdd(p.bar + pi, 1ULL);
        }
      }
    }
    tcgen05_fence_before();
  }
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(static_cast<uint32_t>(C::kTmemCols))
                 : "memory");
  }
}

}  // namespace chain
}  // namespace isst

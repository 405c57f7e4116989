// tcgen05 / TMEM / TMA GEMM for sm_100a:  out[tok, feat] = act[tok, K] . W[feat, K]^T  (+ fused epilogue)
//
// Replaces the cuBLAS Linear / cuDNN Conv1d call sites of the reference's per-chunk step
// (E1 conv1-6, E3, E6, E10, E11, E13, E14, L2, L7, L8, L9).
//
// Both operands are K-major (activations [tok, K] row-major, nn.Linear weights [feat, K]
// row-major), staged by TMA into 128B-swizzled shared memory and consumed by
// tcgen05.mma.cta_group::1.kind::f16 (bf16 x bf16 -> fp32 in TMEM).
//
//   kSwap = false : UMMA M (128 TMEM lanes) = tokens,   UMMA N = kBN features   (prefill / batched / encoder)
//   kSwap = true  : UMMA M (128 TMEM lanes) = features, UMMA N = kBN tokens     (few tokens: weight streaming,
//                   HBM-bound; the 128-row operand is the weight so no tensor-core rows are wasted on padding)
//   kDual         : two weight tiles per k-block (rows f and f + dual_off) -> two accumulators;
//                   epilogue writes silu(acc0) * acc1  (Llama gate/up)
//
// Warp roles: warp0 = TMA producer, warp1 = MMA issuer (one elected lane), warp2 = TMEM allocator,
// warps 4-11 = epilogue (TMEM lane quarter = warp % 4, two column halves); see gemm_sk_kernel below.
#pragma once
#include "common.cuh"

namespace isst {
namespace tc {

constexpr int kBM = 128;       // rows of the 128-lane operand
constexpr int kBK = 64;        // bf16 elements per k-block = 128 bytes = one swizzle atom row
constexpr int kUmmaK = 16;

struct GemmParams {
  int M_tok;        // token rows per batch
  int N_out;        // output features (per half when dual)
  int K;
  int batch;        // activation batches (dim 2 of the activation tensor map)
  int splits;       // split-K factor
  int dual_off;     // row offset of the second weight block (dual)
  int conv_c;       // activation view: K index = tap * conv_c + c  (plain matrices: conv_c = K, conv_s = 1)
  int conv_s;       //   row r, tap t lives at time index r * conv_s + t  (im2col-free strided Conv1d)
  void* out;
  long long ldo;                // elements
  long long out_batch_stride;   // elements
  int out_f32;
  const float* bias;            // [N_out] or null
  const bf16* resid;            // may alias out (same element read then written by the same thread)
  long long ldr;
  long long resid_batch_stride;
  int act;                      // 0 none, 1 gelu(erf)
  float* part_out;              // deferred split reduction: fp32 partials [split][M_tok][N_out] (consumer kernel sums), or null
  int part_splits;              // k-ranges per tile in that mode
  float* ws;                    // split-K workspace
  int* counters;                // split-reduction arrival counters: two halves used by alternate launches
  int counter_half;             // ints per half
  int counter_parity;           // which half this launch uses; the other half is re-armed for the next launch
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
// start>>4 [0,14) | LBO>>4 [16,30) (unused for swizzled K-major, set 1) | SBO>>4 [32,46) = 1024 B (8 rows x 128 B)
// | version=1 [46,48) | layout_type=2 (SWIZZLE_128B) [61,64).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (1<<4), a/b format BF16 (1<<7, 1<<10),
// both K-major, N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc(int umma_m, int umma_n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(umma_n >> 3) << 17) |
         (static_cast<uint32_t>(umma_m >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// tcgen05.ld of 16 columns WITHOUT the completion wait: the registers may only be read after tmem_ld_wait()
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// Final epilogue for 16 consecutive columns of one TMEM lane.
//   normal: lane = token row, columns = features;  swap: lane = feature, columns = tokens.
// Swap-mode epilogue for NC consecutive token columns of one TMEM lane (= one output feature).
template <bool kDual, int NC>
__device__ __forceinline__ void epilogue_store_swap(const GemmParams& p, int b, int f, int col0,
                                                    const float* v0, const float* v1) {
  if (f >= p.N_out) return;
  const float bias = p.bias ? p.bias[f] : 0.f;
  const int n = min(NC, p.M_tok - col0);          // valid token columns of this chunk
  // Residual values are read up front: `resid` may alias `out` (x += ...), so a load issued after a store
  // is serialised behind it (one dependent L2 round trip per token otherwise).
  float rv[NC];
  if (p.resid) {
    const bf16* r = p.resid + static_cast<long long>(b) * p.resid_batch_stride + static_cast<long long>(col0) * p.ldr + f;
#pragma unroll
    for (int i = 0; i < NC; ++i) rv[i] = i < n ? __bfloat162float(r[static_cast<long long>(i) * p.ldr]) : 0.f;
  } else {
#pragma unroll
    for (int i = 0; i < NC; ++i) rv[i] = 0.f;
  }
  const long long o0 = static_cast<long long>(b) * p.out_batch_stride + static_cast<long long>(col0) * p.ldo + f;
  float* of = reinterpret_cast<float*>(p.out) + o0;
  bf16* ob = reinterpret_cast<bf16*>(p.out) + o0;
#pragma unroll
  for (int i = 0; i < NC; ++i) {
    float x = v0[i] + bias;
    if (kDual) x = __fdividef(x, 1.0f + __expf(-x)) * v1[i];      // SiLU(gate) * up; the result is rounded to bf16
    if (p.act == 1) x = gelu_erf(x);
    x += rv[i];
    if (i < n) {
      if (p.out_f32) of[static_cast<long long>(i) * p.ldo] = x;
      else ob[static_cast<long long>(i) * p.ldo] = __float2bfloat16_rn(x);
    }
  }
}

// Non-swap (row = token) epilogue of 16 feature columns, split in two so that the caller can put the TMEM load
// between them: `epi_row_prefetch` issues every global load the chunk needs (bias as 4 x float4, residual as
// 2 x uint4 - they do not depend on the accumulator), `epi_row_finish` does the branch-free arithmetic and the two
// 16-byte stores.  Only for full, aligned chunks (`epi_row_fast`); everything else takes epilogue_store16.
struct EpiRowPre {
  float4 b[4];
  uint4 r0, r1;
};
__device__ __forceinline__ bool epi_row_fast(const GemmParams& p, int col0) {
  return col0 + 16 <= p.N_out && (p.ldo & 7) == 0 && (!p.resid || (p.ldr & 7) == 0) && !p.out_f32;
}
__device__ __forceinline__ void epi_row_prefetch(const GemmParams& p, int b, int tok, int col0, EpiRowPre& e) {
  if (p.bias) {
    const float4* bp = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
    for (int i = 0; i < 4; ++i) e.b[i] = __ldg(bp + i);
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) e.b[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  e.r0 = make_uint4(0u, 0u, 0u, 0u);
  e.r1 = e.r0;
  if (p.resid && tok < p.M_tok) {
    const bf16* r = p.resid + static_cast<long long>(b) * p.resid_batch_stride + static_cast<long long>(tok) * p.ldr + col0;
    e.r0 = *reinterpret_cast<const uint4*>(r);
    e.r1 = *reinterpret_cast<const uint4*>(r + 8);
  }
}
template <bool kDual>
__device__ __forceinline__ void epi_row_finish(const GemmParams& p, int b, int tok, int col0, const float* v0,
                                               const float* v1, const EpiRowPre& e) {
  float y[16];
  const float bz[16] = {e.b[0].x, e.b[0].y, e.b[0].z, e.b[0].w, e.b[1].x, e.b[1].y, e.b[1].z, e.b[1].w,
                        e.b[2].x, e.b[2].y, e.b[2].z, e.b[2].w, e.b[3].x, e.b[3].y, e.b[3].z, e.b[3].w};
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    float x = v0[i] + bz[i];
    if (kDual) x = __fdividef(x, 1.0f + __expf(-x)) * v1[i];
    y[i] = x;
  }
  if (p.act == 1) {
#pragma unroll
    for (int i = 0; i < 16; ++i) y[i] = gelu_erf(y[i]);
  }
  const uint32_t rr[8] = {e.r0.x, e.r0.y, e.r0.z, e.r0.w, e.r1.x, e.r1.y, e.r1.z, e.r1.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) {            // zeros when there is no residual
    const float2 f2 = unpack_bf16(rr[i]);
    y[2 * i] += f2.x;
    y[2 * i + 1] += f2.y;
  }
  if (tok < p.M_tok) {
    bf16* o = reinterpret_cast<bf16*>(p.out) + static_cast<long long>(b) * p.out_batch_stride + static_cast<long long>(tok) * p.ldo + col0;
    uint4 s0, s1;
    s0.x = pack_bf16(y[0], y[1]);   s0.y = pack_bf16(y[2], y[3]);
    s0.z = pack_bf16(y[4], y[5]);   s0.w = pack_bf16(y[6], y[7]);
    s1.x = pack_bf16(y[8], y[9]);   s1.y = pack_bf16(y[10], y[11]);
    s1.z = pack_bf16(y[12], y[13]); s1.w = pack_bf16(y[14], y[15]);
    *reinterpret_cast<uint4*>(o) = s0;
    *reinterpret_cast<uint4*>(o + 8) = s1;
  }
}

template <bool kDual, bool kSwap>
__device__ __forceinline__ void epilogue_store16(const GemmParams& p, int b, int lane_idx, int col0,
                                                 const float* v0, const float* v1) {
  const long long obase = static_cast<long long>(b) * p.out_batch_stride;
  const long long rbase = static_cast<long long>(b) * p.resid_batch_stride;
  if (!kSwap) {
    const int tok = lane_idx;
    if (tok >= p.M_tok) return;
    float y[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int f = col0 + i;
      float x = v0[i];
      if (f < p.N_out) {
        if (p.bias) x += p.bias[f];
        if (kDual) x = silu(x) * v1[i];
        if (p.act == 1) x = gelu_erf(x);
      }
      y[i] = x;
    }
    const bool full = (col0 + 16 <= p.N_out);
    if (p.resid) {
      const bf16* r = p.resid + rbase + static_cast<long long>(tok) * p.ldr + col0;
      if (full && ((p.ldr & 7) == 0)) {
        uint4 r0 = *reinterpret_cast<const uint4*>(r);
        uint4 r1 = *reinterpret_cast<const uint4*>(r + 8);
        const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float2 f2 = unpack_bf16(rr[i]);
          y[2 * i] += f2.x;
          y[2 * i + 1] += f2.y;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (col0 + i < p.N_out) y[i] += __bfloat162float(r[i]);
      }
    }
    if (p.out_f32) {
      float* o = reinterpret_cast<float*>(p.out) + obase + static_cast<long long>(tok) * p.ldo + col0;
      if (full && ((p.ldo & 3) == 0)) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          *reinterpret_cast<float4*>(o + 4 * i) = make_float4(y[4 * i], y[4 * i + 1], y[4 * i + 2], y[4 * i + 3]);
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (col0 + i < p.N_out) o[i] = y[i];
      }
    } else {
      bf16* o = reinterpret_cast<bf16*>(p.out) + obase + static_cast<long long>(tok) * p.ldo + col0;
      if (full && ((p.ldo & 7) == 0)) {
        uint4 s0, s1;
        s0.x = pack_bf16(y[0], y[1]);   s0.y = pack_bf16(y[2], y[3]);
        s0.z = pack_bf16(y[4], y[5]);   s0.w = pack_bf16(y[6], y[7]);
        s1.x = pack_bf16(y[8], y[9]);   s1.y = pack_bf16(y[10], y[11]);
        s1.z = pack_bf16(y[12], y[13]); s1.w = pack_bf16(y[14], y[15]);
        *reinterpret_cast<uint4*>(o) = s0;
        *reinterpret_cast<uint4*>(o + 8) = s1;
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (col0 + i < p.N_out) o[i] = __float2bfloat16_rn(y[i]);
      }
    }
  } else {
    epilogue_store_swap<kDual, 16>(p, b, lane_idx, col0, v0, v1);
  }
}

// =================================================================================================
// Persistent data-parallel + stream-K kernel.
//
// G CTAs (G <= #SMs, one per SM) stay resident for the whole GEMM:
//   * data-parallel part: the first floor(tiles / G) * G tiles are dealt round-robin (CTA c takes tiles
//     c, c + G, ...), whole k-range each, so the tiles in flight at any moment are neighbours (token
//     tile fastest) that share weight / activation tiles in L2;
//   * stream-K part: the remaining tiles (all of them when tiles < G, the weight-streaming decode
//     case) form a linear sequence of (tile, k-block) units cut into G equal contiguous ranges, so
//     every SM streams the same number of bytes whatever the tile count.
// The TMA ring never drains at tile boundaries and the accumulator is double-buffered in TMEM: the
// epilogue of one segment (8 warps) overlaps the MMAs of the next.  A tile whose k-range is shared by
// several CTAs is reduced through an fp32 workspace: every contributor parks its partial and bumps
// the tile's counter; once all have arrived (all CTAs are co-resident, so waiting cannot deadlock)
// each contributor reduces its own share of the columns in contributor order (deterministic).
// =================================================================================================
struct SkParams {
  int tiles_tok, tiles_feat;   // tiles per batch entry
  int num_kb;                  // k-blocks per tile
  long long tiles;             // batch * tiles_tok * tiles_feat
  long long tiles_dp;          // tiles handled data-parallel, dealt round-robin (tile i * G + cta)
  long long units_sk;          // (tiles - tiles_dp) * num_kb
  int g_sk;                    // CTAs taking part in the stream-K part (<= grid, <= units_sk)
  unsigned long long* dbg;     // optional [grid][8] globaltimer stamps (phase analysis), may be null
};


template <int kBN, bool kDual, bool kSwap>
struct SkCfg {
  static constexpr int kActRows = kSwap ? kBN : kBM;
  static constexpr int kWRows = kSwap ? kBM : kBN;
  static constexpr int kNW = kDual ? 2 : 1;
  static constexpr int kActBytes = kActRows * kBK * 2;
  static constexpr int kWBytes = kWRows * kBK * 2;
  static constexpr int kStageBytes = kActBytes + kNW * kWBytes;
  static constexpr int kStagesRaw = (224 * 1024) / kStageBytes;     // 227 KB per CTA minus barriers / alignment slack
  static constexpr int kStages = kStagesRaw > 10 ? 10 : kStagesRaw;
  static constexpr int kAccAll = kBN * kNW;                       // accumulator columns of one tile
  static constexpr int kAccBufs = (2 * kAccAll <= 512) ? 2 : 1;
  static constexpr int kTmemColsRaw = kAccBufs * kAccAll;
  static constexpr int kTmemCols = kTmemColsRaw <= 32 ? 32 : (kTmemColsRaw <= 64 ? 64 : (kTmemColsRaw <= 128 ? 128 : (kTmemColsRaw <= 256 ? 256 : 512)));
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 512 /*barriers*/;
  // warp 0 TMA, warp 1 MMA, warp 2 TMEM alloc, warps 4.. epilogue: 8 warps (2 column halves) for the big
  // tensor-bound tiles, 4 for the weight-streaming (swap) mode whose epilogue is tiny - the smaller CTA leaves
  // registers for a co-resident neighbour kernel under programmatic dependent launch
  static constexpr int kEpiWarps = 8;
  static constexpr int kEpiThreads = kEpiWarps * 32;
  static constexpr int kThreadsTotal = 128 + kEpiThreads;
  static constexpr int kEpiHalves = (kEpiWarps == 8 && kBN >= 32) ? 2 : 1;     // column halves handled by warps 4-7 / 8-11
  static constexpr int kHalfCols = kBN / kEpiHalves;
  static constexpr int kNC = (kSwap && !kDual && kHalfCols >= 32) ? 32 : 16;   // columns per epilogue step
  static_assert(kBN % 16 == 0 && kBN >= 16 && kBN <= 256, "UMMA N");
  static_assert(kTmemColsRaw <= 512, "the accumulators must fit TMEM");
  static_assert(kActBytes % 1024 == 0 && kWBytes % 1024 == 0, "tiles must keep 1024B alignment");
};

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

struct SkTile { int b, tf, tt; };
__device__ __forceinline__ SkTile sk_tile(long long tile, const SkParams& sk) {
  SkTile t;
  t.tt = static_cast<int>(tile % sk.tiles_tok);
  const long long rest = tile / sk.tiles_tok;
  t.tf = static_cast<int>(rest % sk.tiles_feat);
  t.b = static_cast<int>(rest / sk.tiles_feat);
  return t;
}
// CTA that owns stream-K unit u when CTA c starts at floor(U * c / G)
__device__ __forceinline__ int sk_cta_of(long long u, long long U, int G) {
  return static_cast<int>(((u + 1) * G - 1) / U);
}

// Walks the work of one CTA: first its data-parallel tiles (whole k-range), then its stream-K range.
// All three roles (producer, MMA, epilogue) iterate the same sequence of segments.
struct SkWalker {
  const SkParams& sk;
  int G, cta;
  long long dp_i, dp_n;        // data-parallel tiles done / total for this CTA
  long long u, u1;             // stream-K unit cursor / end (units are relative to tile tiles_dp)
  __device__ SkWalker(const SkParams& s, int G_, int cta_) : sk(s), G(G_), cta(cta_) {
    dp_i = 0;
    dp_n = (sk.tiles_dp + G - 1 - cta) / G;      // tiles i * G + cta below tiles_dp
    if (cta < sk.g_sk) { u = sk.units_sk * cta / sk.g_sk; u1 = sk.units_sk * (cta + 1) / sk.g_sk; }
    else { u = 0; u1 = 0; }
  }
  // next segment: tile, [kb_begin, kb_end); returns false when the CTA is done
  __device__ bool next(long long& tile, int& kb_begin, int& kb_end) {
    if (dp_i < dp_n) {
      tile = dp_i * G + cta;
      kb_begin = 0; kb_end = sk.num_kb;
      ++dp_i;
      return true;
    }
    if (u < u1) {
      tile = sk.tiles_dp + u / sk.num_kb;
      kb_begin = static_cast<int>(u % sk.num_kb);
      kb_end = static_cast<int>(min(static_cast<long long>(sk.num_kb), kb_begin + (u1 - u)));
      u += kb_end - kb_begin;
      return true;
    }
    return false;
  }
};

template <int kBN, bool kDual, bool kSwap>
__global__ void __launch_bounds__(384) __maxnreg__(kSwap ? 128 : 168)
gemm_sk_kernel(const __grid_constant__ CUtensorMap tm_act, const __grid_constant__ CUtensorMap tm_w,
               const GemmParams p, const SkParams sk) {
  using C = SkCfg<kBN, kDual, kSwap>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes);
  uint64_t* empty_bar = full_bar + C::kStages;
  uint64_t* tfull_bar = empty_bar + C::kStages;     // [2] accumulator ready
  uint64_t* tempty_bar = tfull_bar + 2;             // [2] accumulator drained
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int G = gridDim.x, cta = blockIdx.x;
  const int nkb = sk.num_kb;
  if (sk.dbg && threadIdx.x == 0) sk.dbg[cta * 8 + 0] = gtimer();
  pdl_launch_dependents();      // the next kernel may start its own prologue / weight prefetch

  // Start-up: the producer warp initialises the barriers, ARRIVES at named barrier 2 and starts requesting weight
  // tiles at once; everybody else waits there for the barriers and for the TMEM allocation (the producer needs
  // neither the TMEM address nor the other warps).
  uint32_t tmem_base = 0;
  if (warp == 0) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_w)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_act)) : "memory");
      for (int s = 0; s < C::kStages; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(&tfull_bar[a], 1);
        mbar_init(&tempty_bar[a], C::kEpiThreads);
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("bar.arrive 2, %0;" ::"n"(C::kThreadsTotal) : "memory");
  } else {
    if (warp == 2) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                   "r"(static_cast<uint32_t>(C::kTmemCols))
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    asm volatile("bar.sync 2, %0;" ::"n"(C::kThreadsTotal) : "memory");
    tcgen05_fence_after();
    tmem_base = *tmem_ptr_smem;
  }

  if (warp == 0) {
    // ================= TMA producer =================
    // Weights do not depend on the previous kernel: the first ring-full of weight tiles is requested BEFORE
    // pdl_wait(), so it streams from HBM while the predecessor is still finishing; the activation tiles of
    // those stages follow once the dependency is resolved.
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      SkWalker w(sk, G, cta);
      long long tile;
      int kb_begin, kb_end;
      bool first = true;
      int pre = 0;                                   // stages whose weights were requested ahead of the wait
      int pre_row[C::kStages], pre_k0[C::kStages], pre_b[C::kStages];
      bool waited = false;
      while (w.next(tile, kb_begin, kb_end)) {
        const SkTile t = sk_tile(tile, sk);
        const int act_row0 = t.tt * C::kActRows;
        const int w_row0 = t.tf * C::kWRows;
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          if (!waited && pre == C::kStages) {
            // ring full of prefetched weights: resolve the dependency, then feed the activations
            pdl_wait();
            waited = true;
            for (int s2 = 0; s2 < pre; ++s2) {
              const int k0p = pre_k0[s2];
              const int tapp = k0p / p.conv_c;
              tma_load_4d(smem + s2 * C::kStageBytes, &tm_act, &full_bar[s2], k0p - tapp * p.conv_c, tapp % p.conv_s,
                          pre_row[s2] + tapp / p.conv_s, pre_b[s2]);
            }
          }
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * C::kStageBytes;
          uint8_t* sw = sa + C::kActBytes;
          mbar_expect_tx(&full_bar[stage], C::kStageBytes);
          const int k0 = kb * kBK;
          const int tap = k0 / p.conv_c;
          if (waited) {
            tma_load_4d(sa, &tm_act, &full_bar[stage], k0 - tap * p.conv_c, tap % p.conv_s, act_row0 + tap / p.conv_s, t.b);
          } else {
            pre_row[pre] = act_row0; pre_k0[pre] = k0; pre_b[pre] = t.b;
            ++pre;
          }
#pragma unroll
          for (int h = 0; h < C::kWRows / kBM; ++h)      // the weight map's box is 128 rows
            tma_load_2d(sw + h * (kBM * kBK * 2), &tm_w, &full_bar[stage], k0, w_row0 + h * kBM);
          if (kDual) tma_load_2d(sw + C::kWBytes, &tm_w, &full_bar[stage], k0, w_row0 + p.dual_off);
          if (first && sk.dbg) sk.dbg[cta * 8 + 1] = gtimer();
          first = false;
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
      }
      if (!waited) {                                 // fewer units than stages
        pdl_wait();
        for (int s2 = 0; s2 < pre; ++s2) {
          const int k0p = pre_k0[s2];
          const int tapp = k0p / p.conv_c;
          tma_load_4d(smem + s2 * C::kStageBytes, &tm_act, &full_bar[s2], k0p - tapp * p.conv_c, tapp % p.conv_s,
                      pre_row[s2] + tapp / p.conv_s, pre_b[s2]);
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    constexpr uint32_t idesc = make_idesc(kBM, kBN);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    SkWalker w(sk, G, cta);
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
          const uint32_t sa = smem_u32(smem + stage * C::kStageBytes);
          const uint32_t sw = sa + C::kActBytes;
          const uint64_t d_act = make_smem_desc(sa);
          const uint64_t d_w0 = make_smem_desc(sw);
          const uint64_t d_w1 = make_smem_desc(sw + C::kWBytes);
#pragma unroll
          for (int k = 0; k < kBK / kUmmaK; ++k) {
            const uint64_t koff = static_cast<uint64_t>((k * kUmmaK * 2) >> 4);
            const uint32_t accum = (kb > kb_begin || k > 0) ? 1u : 0u;
            if (!kSwap) {
              umma_bf16(tacc, d_act + koff, d_w0 + koff, idesc, accum);
              if (kDual) umma_bf16(tacc + kBN, d_act + koff, d_w1 + koff, idesc, accum);
            } else {
              umma_bf16(tacc, d_w0 + koff, d_act + koff, idesc, accum);
              if (kDual) umma_bf16(tacc + kBN, d_w1 + koff, d_act + koff, idesc, accum);
            }
          }
          umma_commit(&empty_bar[stage]);
          if (kb == kb_end - 1) umma_commit(&tfull_bar[acc]);
        }
        __syncwarp();
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      }
      if (++acc == C::kAccBufs) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 4) {
    // ================= epilogue: 8 warps = 4 TMEM lane quarters x 2 column halves =================
    pdl_wait();                                   // residual / workspace / output buffers belong to earlier kernels until now
    const int q = warp & 3;                       // TMEM lane quarter accessible to this warp
    const int half = (warp - 4) >> 2;             // column half of the tile
    const int r = q * 32 + lane;                  // TMEM lane == row of the 128-row operand
    const int et = threadIdx.x - 128;             // 0..kEpiThreads-1
    const bool works = half < C::kEpiHalves;
    const int* counters = p.counters + (p.counter_parity ? p.counter_half : 0);
    // the other half of the counter array belongs to the next GEMM launch (whatever its grid size): re-arm it now
    // (its previous user completed before pdl_wait() returned)
    for (int i = cta * C::kEpiThreads + et; i < p.counter_half; i += G * C::kEpiThreads)
      p.counters[(p.counter_parity ? 0 : p.counter_half) + i] = 0;
    const int hc0 = half * C::kHalfCols;          // first column of this warp's half
    constexpr int NC = C::kNC;
    constexpr size_t kSlot = static_cast<size_t>(C::kAccAll) * kBM;
    int acc = 0;
    uint32_t acc_phase = 0;
    SkWalker w(sk, G, cta);
    long long tile;
    int kb_begin, kb_end;
    long long part_tile[2];                       // partial tiles of this CTA (at most its first and last stream-K tile)
    int n_part = 0;
    bool first_seg = true;
    while (w.next(tile, kb_begin, kb_end)) {
      const SkTile t = sk_tile(tile, sk);
      const int col_base = (kSwap ? t.tt : t.tf) * kBN;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tcgen05_fence_after();
      if (first_seg && sk.dbg && et == 0) sk.dbg[cta * 8 + 2] = gtimer();
      first_seg = false;
      const uint32_t taddr0 = tmem_base + acc * C::kAccAll + (static_cast<uint32_t>(q * 32) << 16);
      const int lane_idx = (kSwap ? t.tf * kBM : t.tt * C::kActRows) + r;
      const uint32_t taddr = taddr0;
      if (kb_begin == 0 && kb_end == nkb) {
        // ---- the whole k-range of this tile was accumulated here: final epilogue straight from TMEM ----
        if (works) {
#pragma unroll 1
          for (int c = hc0; c < hc0 + C::kHalfCols; c += (kSwap ? NC : 32)) {
            if (!kSwap) {
              // two 16-column chunks per step: their global loads and both TMEM loads are in flight together
              constexpr int W = (C::kHalfCols % 32 == 0) ? 2 : 1;
              const bool fast = epi_row_fast(p, col_base + c + (W - 1) * 16);      // warp-uniform
              if (fast) {
                EpiRowPre pre[W];
                uint32_t r0[W][16], r1[W][16];
#pragma unroll
                for (int h = 0; h < W; ++h) epi_row_prefetch(p, t.b, lane_idx, col_base + c + 16 * h, pre[h]);
#pragma unroll
                for (int h = 0; h < W; ++h) {
                  tmem_ld16_issue(taddr + c + 16 * h, r0[h]);
                  if (kDual) tmem_ld16_issue(taddr + kBN + c + 16 * h, r1[h]);
                }
                tmem_ld_wait();
#pragma unroll
                for (int h = 0; h < W; ++h) {
                  float v0[16], v1[16];
#pragma unroll
                  for (int i = 0; i < 16; ++i) { v0[i] = __uint_as_float(r0[h][i]); v1[i] = kDual ? __uint_as_float(r1[h][i]) : 0.f; }
                  epi_row_finish<kDual>(p, t.b, lane_idx, col_base + c + 16 * h, v0, v1, pre[h]);
                }
              } else {
#pragma unroll 1
                for (int h = 0; h < W; ++h) {
                  float v0[16], v1[16];
                  tmem_ld16(taddr + c + 16 * h, v0);
                  if (kDual) tmem_ld16(taddr + kBN + c + 16 * h, v1);
                  epilogue_store16<kDual, false>(p, t.b, lane_idx, col_base + c + 16 * h, v0, v1);
                }
              }
              if (W == 1) c -= 16;          // (kHalfCols is a multiple of 32 for every non-swap tile shape in use)
            } else {
              float v0[NC], v1[NC];
#pragma unroll
              for (int j = 0; j < NC; j += 16) {
                tmem_ld16(taddr + c + j, v0 + j);
                if (kDual) tmem_ld16(taddr + kBN + c + j, v1 + j);
              }
              epilogue_store_swap<kDual, NC>(p, t.b, lane_idx, col_base + c, v0, v1);
            }
          }
        }
        tcgen05_fence_before();
        mbar_arrive(&tempty_bar[acc]);
      } else {
        if (kSwap && p.part_out) {
          // ---- deferred split reduction: this CTA's fp32 partial goes to part_out[split][tok][feature]; the
          //      consumer row kernel (norm / RoPE-append) sums the splits - no fence, counter or wait here ----
          const int split = cta % p.part_splits;
          if (works && lane_idx - lane < p.N_out) {              // warp-uniform: tcgen05.ld is warp-collective
            const bool row_ok = lane_idx < p.N_out;
            float* dst = p.part_out + (static_cast<size_t>(split) * p.M_tok + col_base + hc0) * p.N_out + lane_idx;
            const int n_tok = p.M_tok - col_base - hc0;          // valid token columns from hc0 on
#pragma unroll 1
            for (int c = 0; c < C::kHalfCols; c += 16) {
              float v[16];
              tmem_ld16(taddr + hc0 + c, v);
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (row_ok && c + i < n_tok) dst[static_cast<size_t>(c + i) * p.N_out] = v[i];
            }
          }
          tcgen05_fence_before();
          mbar_arrive(&tempty_bar[acc]);
          if (++acc == C::kAccBufs) { acc = 0; acc_phase ^= 1; }
          continue;
        }
        // ---- partial tile: park the fp32 partial and announce it; the reduction happens after the walk ----
        const long long ut = (tile - sk.tiles_dp) * nkb;
        const int c_first = sk_cta_of(ut, sk.units_sk, sk.g_sk);
        float* mine = p.ws + static_cast<size_t>(cta == c_first ? G + cta : cta) * kSlot;
        if (works) {
#pragma unroll 1
          for (int cc = 0; cc < C::kNW; ++cc)                  // accumulator blocks of kBN columns: gate | up
#pragma unroll 1
            for (int c = hc0; c < hc0 + C::kHalfCols; c += 16) {
              float v[16];
              tmem_ld16(taddr0 + cc * kBN + c, v);
#pragma unroll
              for (int i = 0; i < 16; ++i) mine[(cc * kBN + c + i) * kBM + r] = v[i];
            }
        }
        tcgen05_fence_before();
        mbar_arrive(&tempty_bar[acc]);            // the accumulator is free again: MMAs of the next segment go on
        __threadfence();
        asm volatile("bar.sync 1, %0;" ::"n"(C::kEpiThreads) : "memory");
        if (et == 0) atomicAdd(const_cast<int*>(&counters[c_first]), 1);
        part_tile[n_part++] = tile;
      }
      if (++acc == C::kAccBufs) { acc = 0; acc_phase ^= 1; }
    }
    if (sk.dbg && et == 0) sk.dbg[cta * 8 + 3] = gtimer();
    // ---- reduce the shared tiles: every contributor takes its share of the 16-column chunks ----
    for (int pi = 0; pi < n_part; ++pi) {
      const long long tl = part_tile[pi];
      const long long ut = (tl - sk.tiles_dp) * nkb;
      const int c_first = sk_cta_of(ut, sk.units_sk, sk.g_sk);
      const int c_last = sk_cta_of(ut + nkb - 1, sk.units_sk, sk.g_sk);
      const int nc = c_last - c_first + 1;
      if (et == 0) {
        while (ld_acquire(&counters[c_first]) < nc) __nanosleep(32);
      }
      asm volatile("bar.sync 1, %0;" ::"n"(C::kEpiThreads) : "memory");
      const SkTile t = sk_tile(tl, sk);
      const int col_base = (kSwap ? t.tt : t.tf) * kBN;
      // chunk list: kBN / 16 chunks of 16 columns dealt to (contributor, half) pairs
      constexpr int kChunks = kBN / 16;
      const int workers = nc * C::kEpiHalves;
      const int me = (cta - c_first) * C::kEpiHalves + half;
      const int ch0 = kChunks * me / workers, ch1 = kChunks * (me + 1) / workers;
#pragma unroll 1
      for (int ch = ch0; ch < ch1; ++ch) {
        const int c = ch * 16;
        const int lane_idx = (kSwap ? t.tf * kBM : t.tt * C::kActRows) + r;
        float v0[16], v1[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) { v0[i] = 0.f; v1[i] = 0.f; }
#pragma unroll 4
        for (int cc = c_first; cc <= c_last; ++cc) {
          const float* src = p.ws + static_cast<size_t>(cc == c_first ? G + cc : cc) * kSlot;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            v0[i] += __ldcg(&src[(c + i) * kBM + r]);
            if (kDual) v1[i] += __ldcg(&src[(kBN + c + i) * kBM + r]);
          }
        }
        if (kSwap) epilogue_store_swap<kDual, 16>(p, t.b, lane_idx, col_base + c, v0, v1);
        else epilogue_store16<kDual, false>(p, t.b, lane_idx, col_base + c, v0, v1);
      }
    }
    if (sk.dbg && et == 0) sk.dbg[cta * 8 + 5] = gtimer();
    tcgen05_fence_before();
  }
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(static_cast<uint32_t>(C::kTmemCols))
                 : "memory");
  }
  if (sk.dbg && threadIdx.x == 0) sk.dbg[cta * 8 + 4] = gtimer();
}

}  // namespace tc
}  // namespace isst

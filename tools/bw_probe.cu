// Read-bandwidth probe for the weight-streaming GEMMs (not product code): how fast can N CTAs (one per SM) pull a
// K-major bf16 weight matrix out of HBM into shared memory, as a function of
//   * the copy mechanism: 2-D TMA boxes of 128 rows x 128 B (what the GEMM uses today: 128 scattered 128-byte
//     segments per box, row pitch 8 KB), 1-D bulk copies of contiguous 16 KB (tile-packed weights), plain LDG.128;
//   * bytes in flight per SM (ring depth);
//   * CTA count (112 vs 148).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bw_probe tools/bw_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE;\nbra WAIT;\nDONE:\n}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
               "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

constexpr int kTileBytes = 16384;   // 128 rows x 64 bf16

// mode 0: 2-D TMA box (64 cols x 128 rows) of a [rows][K] matrix; mode 1: contiguous 16 KB bulk copies.
// Each CTA walks `tiles_per_cta` tiles: row tile = cta-major so that a CTA streams whole weight rows like the GEMM.
template <int kMode>
__global__ void __launch_bounds__(128) ring_kernel(const __grid_constant__ CUtensorMap tm, const uint8_t* base, int K, int num_kb,
                                                   long long tiles_total, int stages, unsigned long long* sink) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + stages * kTileBytes);
  uint64_t* empty = full + stages;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // contiguous unit range per CTA (stream-K style): unit u = (row_tile, kb)
  const long long per = (tiles_total + gridDim.x - 1) / gridDim.x;
  const long long u0 = static_cast<long long>(blockIdx.x) * per;
  const long long u1 = min(tiles_total, u0 + per);
  if (threadIdx.x == 0) {          // producer
    int stage = 0, phase = 0;
    for (long long u = u0; u < u1; ++u) {
      mbar_wait(&empty[stage], phase ^ 1);
      mbar_expect_tx(&full[stage], kTileBytes);
      const int rt = static_cast<int>(u / num_kb), kb = static_cast<int>(u % num_kb);
      if (kMode == 0) tma_load_2d(smem + stage * kTileBytes, &tm, &full[stage], kb * 64, rt * 128);
      else bulk_load_1d(smem + stage * kTileBytes, base + static_cast<size_t>(u) * kTileBytes, kTileBytes, &full[stage]);
      if (++stage == stages) { stage = 0; phase ^= 1; }
    }
  } else if (threadIdx.x == 32) {  // consumer: touch one word per tile, release the slot
    int stage = 0, phase = 0;
    unsigned long long acc = 0;
    for (long long u = u0; u < u1; ++u) {
      mbar_wait(&full[stage], phase);
      acc += *reinterpret_cast<volatile uint32_t*>(smem + stage * kTileBytes + 64);
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[stage])) : "memory");
      if (++stage == stages) { stage = 0; phase ^= 1; }
    }
    if (acc == 0x123456789ull) *sink = acc;
  }
}

// GEMM-like stage: two 16 KB weight boxes (gate rows, up rows) + optionally one 8 KB activation box [64 tokens x 64 k]
// of a [64][K] matrix that is either shared by all CTAs (what the GEMM does: every CTA asks for the same lines at
// about the same time) or private per CTA.  act_mode: 0 none, 1 shared, 2 private.
__global__ void __launch_bounds__(128) gemm_like_kernel(const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap tm_act,
                                                        int num_kb, int tiles, int up_off_rows, int stages, int act_mode,
                                                        unsigned long long* sink, int blocked = 0) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stage_bytes = 2 * kTileBytes + 8192;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + stages * stage_bytes);
  uint64_t* empty = full + stages;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int stage = 0, phase = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x)
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_expect_tx(&full[stage], 2 * kTileBytes + (act_mode ? 8192 : 0));
        uint8_t* s = smem + stage * stage_bytes;
        if (act_mode) tma_load_2d(s, &tm_act, &full[stage], kb * 64, act_mode == 2 ? blockIdx.x * 64 : 0);
        if (!blocked) {
          tma_load_2d(s + 8192, &tm, &full[stage], kb * 64, tile * 128);
          tma_load_2d(s + 8192 + kTileBytes, &tm, &full[stage], kb * 64, tile * 128 + up_off_rows);
        } else {
          // pre-tiled weights [row tile][k-block][128 rows][64]: every box is one contiguous 16 KB block
          tma_load_2d(s + 8192, &tm, &full[stage], 0, (tile * num_kb + kb) * 128);
          tma_load_2d(s + 8192 + kTileBytes, &tm, &full[stage], 0, ((tile + up_off_rows / 128) * num_kb + kb) * 128);
        }
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
  } else if (threadIdx.x == 32) {
    int stage = 0, phase = 0;
    unsigned long long acc = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x)
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full[stage], phase);
        acc += *reinterpret_cast<volatile uint32_t*>(smem + stage * stage_bytes + 64);
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[stage])) : "memory");
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
    if (acc == 0x123456789ull) *sink = acc;
  }
}

// Decode-attention-like ring: a producer WARP fills stages of 64 K rows + 64 V rows (256 B each, padded to 272 B in
// shared memory) with 128 row-sized bulk copies per stage (4 per lane); one consumer thread releases the stages.
// Tests whether 256-byte bulk copies sustain HBM bandwidth (2 CTAs per SM, 3 stages).
__global__ void __launch_bounds__(160, 2) rows_ring_kernel(const uint8_t* base, long long n_rows_total, int tiles_per_cta, int stages,
                                                           unsigned long long* sink) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  constexpr int kStage = 2 * 64 * 272;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + stages * kStage);
  uint64_t* empty = full + stages;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row0 = static_cast<long long>(blockIdx.x) * tiles_per_cta * 128;
  if (warp == 4) {
    int stage = 0, phase = 0;
    for (int t = 0; t < tiles_per_cta; ++t) {
      mbar_wait(&empty[stage], phase ^ 1);
      if (lane == 0) mbar_expect_tx(&full[stage], 128 * 256);
      __syncwarp();
      uint8_t* s = smem + stage * kStage;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int r = q * 32 + lane;                         // 0..127: 64 K rows then 64 V rows
        const long long gr = (row0 + static_cast<long long>(t) * 128 + r) % n_rows_total;
        bulk_load_1d(s + r * 272, base + gr * 256, 256, &full[stage]);
      }
      if (++stage == stages) { stage = 0; phase ^= 1; }
    }
  } else if (threadIdx.x == 0) {
    int stage = 0, phase = 0;
    unsigned long long acc = 0;
    for (int t = 0; t < tiles_per_cta; ++t) {
      mbar_wait(&full[stage], phase);
      acc += *reinterpret_cast<volatile uint32_t*>(smem + stage * kStage + 64);
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[stage])) : "memory");
      if (++stage == stages) { stage = 0; phase ^= 1; }
    }
    if (acc == 0x123456789ull) *sink = acc;
  }
}

__global__ void __launch_bounds__(512) ldg_kernel(const uint4* __restrict__ p, long long n_vec, unsigned long long* sink) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  uint32_t acc = 0;
  for (; i + 7 * stride < n_vec; i += 8 * stride) {
    uint4 v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __ldcs(p + i + j * stride);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc += v[j].x ^ v[j].w;
  }
  if (acc == 0x12345678u) *sink = acc;
}

int main() {
  CK(cudaSetDevice(0));
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  auto encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  const int K = 4096, rows = 28672 * 8;                       // 8 gate/up matrices = 1.88 GB (>> L2)
  const size_t bytes = static_cast<size_t>(rows) * K * 2;
  uint8_t* w;
  CK(cudaMalloc(&w, bytes));
  CK(cudaMemset(w, 1, bytes));
  unsigned long long* sink;
  CK(cudaMalloc(&sink, 8));
  CUtensorMap tm;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(K) * 2};
  cuuint32_t box[2] = {64, 128};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  CK(cudaFuncSetAttribute(ring_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
  CK(cudaFuncSetAttribute(ring_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int num_kb = K / 64;
  // (a) long streams: whole 1.88 GB
  const long long tiles_all = static_cast<long long>(rows / 128) * num_kb;
  printf("# long stream: %.2f GB per launch\n", bytes / 1e9);
  for (int mode = 0; mode < 2; ++mode)
    for (int ctas : {112, 148})
      for (int stages : {4, 8, 12}) {
        const size_t smem = static_cast<size_t>(stages) * kTileBytes + 1024 + 256;
        float best = 1e9f;
        for (int it = 0; it < 3; ++it) {
          CK(cudaEventRecord(e0));
          if (mode == 0) ring_kernel<0><<<ctas, 128, smem>>>(tm, w, K, num_kb, tiles_all, stages, sink);
          else ring_kernel<1><<<ctas, 128, smem>>>(tm, w, K, num_kb, tiles_all, stages, sink);
          CK(cudaEventRecord(e1));
          CK(cudaEventSynchronize(e1));
          float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
          best = ms < best ? ms : best;
        }
        printf("%s ctas=%3d stages=%2d (%3d KB in flight/SM): %7.1f us  %6.2f TB/s  %5.1f GB/s per CTA\n",
               mode == 0 ? "tma2d 128x128B" : "bulk1d 16KB   ", ctas, stages, stages * 16, best * 1e3, bytes / best / 1e9,
               bytes / best / 1e6 / ctas);
      }
  for (int ctas : {148, 296, 592}) {
    float best = 1e9f;
    for (int it = 0; it < 3; ++it) {
      CK(cudaEventRecord(e0));
      ldg_kernel<<<ctas, 512>>>(reinterpret_cast<const uint4*>(w), static_cast<long long>(bytes / 16), sink);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      best = ms < best ? ms : best;
    }
    printf("ldg.128 ctas=%3d x512 thr, 8 loads in flight/thread: %7.1f us  %6.2f TB/s\n", ctas, best * 1e3, bytes / best / 1e9);
  }
  // (b) short streams: one gate/up matrix (235 MB) per launch, rotating over the 8 copies (cold L2)
  const long long tiles_one = tiles_all / 8;
  printf("# short stream: %.1f MB per launch (one gate/up matrix), rotating over 8 copies\n", bytes / 8 / 1e6);
  for (int mode = 0; mode < 2; ++mode)
    for (int ctas : {112, 148})
      for (int stages : {4, 8, 12}) {
        const size_t smem = static_cast<size_t>(stages) * kTileBytes + 1024 + 256;
        const int iters = 16;
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        for (int it = 0; it < iters; ++it) {
          const uint8_t* base = w + static_cast<size_t>(it % 8) * (bytes / 8);
          // 2-D map covers the whole buffer: offset the unit range instead of the pointer
          if (mode == 0) {
            CUtensorMap tmi;
            cuuint64_t d2[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(rows / 8)};
            encode(&tmi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<uint8_t*>(base), d2, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            ring_kernel<0><<<ctas, 128, smem>>>(tmi, base, K, num_kb, tiles_one, stages, sink);
          } else {
            ring_kernel<1><<<ctas, 128, smem>>>(tm, base, K, num_kb, tiles_one, stages, sink);
          }
        }
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        const double us = ms * 1e3 / iters;
        printf("%s ctas=%3d stages=%2d: %7.1f us per launch  %6.2f TB/s\n", mode == 0 ? "tma2d 128x128B" : "bulk1d 16KB   ", ctas, stages, us,
               bytes / 8 / us / 1e6);
      }

  // (c) GEMM-like stages: gate/up of one matrix (112 dual tiles), 4 or 5 stages of 40 KB, activation variants
  {
    uint8_t* act;
    CK(cudaMalloc(&act, static_cast<size_t>(148) * 64 * K * 2));
    CK(cudaMemset(act, 1, static_cast<size_t>(148) * 64 * K * 2));
    CUtensorMap tma;
    cuuint64_t da[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(148 * 64)};
    cuuint32_t boxa[2] = {64, 64};
    r = encode(&tma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, act, da, strides, boxa, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode act failed %d\n", (int)r); return 1; }
    CK(cudaFuncSetAttribute(gemm_like_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
    printf("# GEMM-like: 112 dual tiles (gate+up 2 x 16 KB per stage) of one 235 MB matrix, rotating over 8 copies\n");
    const char* names[3] = {"no act   ", "act shared", "act private"};
    for (int act_mode = 0; act_mode < 3; ++act_mode)
      for (int stages : {4, 5}) {
        const size_t smem = static_cast<size_t>(stages) * (2 * kTileBytes + 8192) + 1024 + 256;
        const int iters = 16;
        std::vector<CUtensorMap> maps(8);
        for (int i = 0; i < 8; ++i) {
          cuuint64_t d2[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(rows / 8)};
          encode(&maps[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w + static_cast<size_t>(i) * (bytes / 8), d2, strides, box, estr,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        }
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        for (int it = 0; it < iters; ++it)
          gemm_like_kernel<<<112, 128, smem>>>(maps[it % 8], tma, num_kb, 112, 14336, stages, act_mode, sink);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        const double us = ms * 1e3 / iters;
        printf("%s stages=%d: %7.1f us per launch  %6.2f TB/s (weights only)\n", names[act_mode], stages, us, bytes / 8 / us / 1e6);
      }
  }
  // (c2) the same GEMM-like stream over PRE-TILED weights (each 128 x 64 box contiguous in memory)
  {
    uint8_t* act;
    CK(cudaMalloc(&act, static_cast<size_t>(148) * 64 * K * 2));
    CK(cudaMemset(act, 1, static_cast<size_t>(148) * 64 * K * 2));
    CUtensorMap tma;
    cuuint64_t da[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(148 * 64)};
    cuuint32_t boxa[2] = {64, 64};
    r = encode(&tma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, act, da, strides, boxa, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode act failed %d\n", (int)r); return 1; }
    printf("# GEMM-like over pre-tiled weights ([row tile][k-block][128][64], contiguous 16 KB boxes), act shared\n");
    cuuint64_t strides_b[1] = {128};
    for (int stages : {4, 5}) {
      const size_t smem = static_cast<size_t>(stages) * (2 * kTileBytes + 8192) + 1024 + 256;
      const int iters = 16;
      std::vector<CUtensorMap> maps(8);
      for (int i = 0; i < 8; ++i) {
        cuuint64_t d2[2] = {64, static_cast<cuuint64_t>(rows / 8) * (K / 64)};
        r = encode(&maps[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w + static_cast<size_t>(i) * (bytes / 8), d2, strides_b, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode blocked failed %d\n", (int)r); return 1; }
      }
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0));
      for (int it = 0; it < iters; ++it)
        gemm_like_kernel<<<112, 128, smem>>>(maps[it % 8], tma, num_kb, 112, 14336, stages, 1, sink, 1);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      const double us = ms * 1e3 / iters;
      printf("pre-tiled  stages=%d: %7.1f us per launch  %6.2f TB/s (weights only)\n", stages, us, bytes / 8 / us / 1e6);
    }
  }
  // (d) decode-attention-like: 512 CTAs (64 streams x 8 kv heads), 16 tiles of 64 keys each, rows of 256 B
  {
    CK(cudaFuncSetAttribute(rows_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
    printf("# decode-attention-like: 512 CTAs x 16 tiles x (64 K + 64 V rows of 256 B) = 268 MB per launch, 2 CTAs per SM\n");
    const long long n_rows = static_cast<long long>(bytes / 256);
    for (int stages : {2, 3}) {
      const size_t smem = static_cast<size_t>(stages) * 2 * 64 * 272 + 256 + 128;
      const int iters = 8;
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0));
      for (int it = 0; it < iters; ++it)
        rows_ring_kernel<<<512, 160, smem>>>(w + static_cast<size_t>(it % 7) * 268435456ull, 1048576ll, 16, stages, sink);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      const double us = ms * 1e3 / iters;
      printf("rows 256B bulk, stages=%d: %7.1f us per launch  %6.2f TB/s\n", stages, us, 512.0 * 16 * 128 * 256 / us / 1e6);
    }
  }
  CK(cudaGetLastError());
  return 0;
}

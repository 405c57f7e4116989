// Common device/host helpers for the infinisst_b200 sm_100a kernels.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>

typedef __nv_bfloat16 bf16;
typedef __nv_bfloat162 bf162;

namespace isst {

// ---- error plumbing (C-ABI returns int status, message via isst_last_error) ----
extern thread_local std::string g_last_error;
inline int set_error(const std::string& msg) {
  g_last_error = msg;
  return -1;
}
#define ISST_CUDA(call)                                                                     \
  do {                                                                                      \
    cudaError_t _e = (call);                                                                \
    if (_e != cudaSuccess) {                                                                \
      return isst::set_error(std::string(#call) + " failed: " + cudaGetErrorString(_e) +    \
                             " at " + __FILE__ + ":" + std::to_string(__LINE__));           \
    }                                                                                       \
  } while (0)
#define ISST_CHECK(cond, msg)                                                               \
  do {                                                                                      \
    if (!(cond)) return isst::set_error(std::string(msg) + " [" #cond "] at " + __FILE__ +  \
                                        ":" + std::to_string(__LINE__));                    \
  } while (0)
#define ISST_TRY(expr)                                                                      \
  do {                                                                                      \
    int _s = (expr);                                                                        \
    if (_s != 0) return _s;                                                                 \
  } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---- device helpers ----
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float silu(float x) { return x / (1.0f + __expf(-x)); }

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  bf162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t u) {
  bf162 v = *reinterpret_cast<bf162*>(&u);
  return __bfloat1622float2(v);
}
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
// 2^x for softmax probabilities: one MUFU.EX2 (ex2.approx.ftz) instead of exp2f's denormal-safe sequence; results
// below 2^-126 flush to zero, -inf gives 0
__device__ __forceinline__ float exp2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Programmatic dependent launch (PDL): a kernel launched with the programmatic-stream-serialization attribute may
// start while its predecessor in the stream is still running; everything it does before pdl_wait() overlaps the
// predecessor's tail (launch latency, barrier / TMEM setup, weight prefetch).  pdl_wait() returns once the
// predecessor grid has completed and its writes are visible.  Both are no-ops for ordinary launches.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// 16-byte streaming load (read-once data: weights / KV), bypassing L1 allocation.
__device__ __forceinline__ uint4 ld_nc_u4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ uint2 ld_nc_u2(const void* p) {
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}

}  // namespace isst

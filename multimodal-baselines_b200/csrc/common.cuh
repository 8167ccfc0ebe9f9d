// Shared helpers for libmmb_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "mmb_b200.h"

namespace mmb {

// Thread-local last-error text behind mmb_last_error().
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
int sm_count();

#define MMB_CUDA(call)                                           \
  do {                                                           \
    cudaError_t e__ = (call);                                    \
    if (e__ != cudaSuccess) return ::mmb::cuda_fail(e__, #call); \
  } while (0)

#define MMB_LAUNCH_CHECK(name)                                    \
  do {                                                            \
    cudaError_t e__ = cudaGetLastError();                         \
    if (e__ != cudaSuccess) return ::mmb::cuda_fail(e__, name);   \
  } while (0)

#define MMB_REQUIRE(cond, msg)             \
  do {                                     \
    if (!(cond)) {                         \
      ::mmb::set_error("%s: %s", __func__, msg); \
      return MMB_E_INVALID;                \
    }                                      \
  } while (0)

static inline cudaStream_t as_stream(mmb_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Streaming (evict-first) global accesses for data touched exactly once.
__device__ __forceinline__ float4 ld_stream(const float4* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(float4* p, float4 v) { __stcs(p, v); }

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace mmb

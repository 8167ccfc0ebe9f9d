// Projection subtraction (reference: sif_functions.py:69-81 remove_pc):
//   npc == 1: XX = X - (X pc^T) * pc          (line 78)
//   npc  > 1: XX = X - (X pc^T) pc            (line 80; all projections taken from X itself)
// HBM-bound streaming pass: one warp per row, the row lives in registers as float4 (read
// once, written once = 2*d*4 bytes per utterance), the components sit in shared memory.
#include "common.cuh"

namespace mmb {

constexpr int kRmWarps = 8;
constexpr int kRmMaxPc = 8;  // components kept in shared memory per pass

template <int NCH>
__global__ void __launch_bounds__(kRmWarps * 32)
    remove_pc_kernel(const float4* __restrict__ X4, int64_t N, int d4, const float4* __restrict__ pc4,
                     int npc, float4* __restrict__ out4) {
  extern __shared__ float4 spc[];  // npc x d4
  for (int k = threadIdx.x; k < npc * d4; k += blockDim.x) spc[k] = pc4[k];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kRmWarps + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kRmWarps;
  for (int64_t i = warp0; i < N; i += nwarps) {
    float4 x[NCH], r[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int k = lane + 32 * c;
      x[c] = k < d4 ? ld_stream(X4 + (size_t)i * d4 + k) : make_float4(0.f, 0.f, 0.f, 0.f);
      r[c] = x[c];
    }
    for (int p = 0; p < npc; ++p) {
      float dot = 0.f;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int k = lane + 32 * c;
        if (k < d4) {
          const float4 v = spc[p * d4 + k];
          dot = fmaf(x[c].x, v.x, dot);
          dot = fmaf(x[c].y, v.y, dot);
          dot = fmaf(x[c].z, v.z, dot);
          dot = fmaf(x[c].w, v.w, dot);
        }
      }
      dot = warp_sum(dot);
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int k = lane + 32 * c;
        if (k < d4) {
          const float4 v = spc[p * d4 + k];
          r[c].x = fmaf(-dot, v.x, r[c].x);
          r[c].y = fmaf(-dot, v.y, r[c].y);
          r[c].z = fmaf(-dot, v.z, r[c].z);
          r[c].w = fmaf(-dot, v.w, r[c].w);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int k = lane + 32 * c;
      if (k < d4) st_stream(out4 + (size_t)i * d4 + k, r[c]);
    }
  }
}

}  // namespace mmb

using namespace mmb;

extern "C" int mmb_remove_pc(const float* X, int64_t N, int d, const float* pc, int npc, float* out,
                             mmb_stream_t stream) {
  MMB_REQUIRE(X && pc && out, "null pointer");
  MMB_REQUIRE(d > 0 && d % 4 == 0 && d <= 1024, "d must be a multiple of 4, <= 1024");
  MMB_REQUIRE(npc > 0 && npc <= 32, "npc must be in [1, 32]");
  MMB_REQUIRE(N >= 0, "negative size");
  if (N == 0) return MMB_OK;
  const int d4 = d / 4, nch = (d4 + 31) / 32;
  int64_t blocks = ceil_div(N, kRmWarps);
  int64_t cap = (int64_t)sm_count() * 8;
  const int grid = (int)(blocks < cap ? blocks : cap);
  const size_t smem = (size_t)npc * d4 * sizeof(float4);
  cudaStream_t st = as_stream(stream);
#define RM_LAUNCH(NCH)                                                                          \
  remove_pc_kernel<NCH><<<grid, kRmWarps * 32, smem, st>>>((const float4*)X, N, d4, (const float4*)pc, \
                                                          npc, (float4*)out)
  switch (nch) {
    case 1: RM_LAUNCH(1); break;
    case 2: RM_LAUNCH(2); break;
    case 3: RM_LAUNCH(3); break;
    case 4: RM_LAUNCH(4); break;
    default: {
      if (smem > 48 * 1024)
        MMB_CUDA(cudaFuncSetAttribute(remove_pc_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
      RM_LAUNCH(8);
    }
  }
#undef RM_LAUNCH
  MMB_LAUNCH_CHECK("remove_pc");
  return MMB_OK;
}

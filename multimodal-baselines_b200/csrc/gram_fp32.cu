// G = X^T X on CUDA cores in FP32 (reference: sif_functions.py:58-67 -- the Gram is what
// sklearn's randomized SVD sees of X).
//
// Exact-FP32 path: used for small N (MOSI / POM splits, where a tensor-core pipeline cannot
// be filled) and as the on-device cross-check of the tcgen05 3xTF32 kernel (gram_tc.cu).
// Only the upper-triangular 64x64 tiles are computed (G is symmetric); K = N is split
// across CTAs and the partial tiles are summed in a fixed order by a second kernel, so the
// result is deterministic and independent of the grid.
#include "common.cuh"

namespace mmb {

constexpr int kGT = 64;      // tile edge
constexpr int kGK = 32;      // rows of X per shared-memory chunk
constexpr int kGThreads = 256;

__global__ void __launch_bounds__(kGThreads)
    gram_fp32_kernel(const float* __restrict__ X, int64_t N, int d, int nt, int64_t rows_per_split,
                     float* __restrict__ partial) {
  __shared__ __align__(16) float sA[kGK][kGT];
  __shared__ __align__(16) float sB[kGK][kGT];
  // upper-triangular tile pair (ti <= tj) from the linear block index
  int pair = blockIdx.x, ti = 0;
  while (pair >= nt - ti) { pair -= nt - ti; ++ti; }
  const int tj = ti + pair;
  const bool diag = (ti == tj);
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t k_begin = (int64_t)blockIdx.y * rows_per_split;
  int64_t k_end = k_begin + rows_per_split;
  if (k_end > N) k_end = N;

  // each thread stages 2 float4 per tile: rows r0 and r0+16, float4 column c4
  const int c4 = tid & 15, r0 = tid >> 4;
  const int colA = ti * kGT + c4 * 4, colB = tj * kGT + c4 * 4;
  const bool okA = colA < d, okB = colB < d;  // d % 4 == 0
  float4 pa[2], pb[2];
  auto fetch = [&](int64_t k0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int64_t r = k0 + r0 + 16 * h;
      const bool in = r < k_end;
      pa[h] = (in && okA) ? __ldg((const float4*)(X + r * d + colA)) : make_float4(0, 0, 0, 0);
      if (!diag) pb[h] = (in && okB) ? __ldg((const float4*)(X + r * d + colB)) : make_float4(0, 0, 0, 0);
    }
  };
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  if (k_begin < k_end) fetch(k_begin);
  for (int64_t k0 = k_begin; k0 < k_end; k0 += kGK) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      *(float4*)&sA[r0 + 16 * h][c4 * 4] = pa[h];
      if (!diag) *(float4*)&sB[r0 + 16 * h][c4 * 4] = pb[h];
    }
    __syncthreads();
    if (k0 + kGK < k_end) fetch(k0 + kGK);  // next chunk in flight during the FMAs
    const float(*B)[kGT] = diag ? sA : sB;
#pragma unroll
    for (int k = 0; k < kGK; ++k) {
      const float4 a = *(const float4*)&sA[k][ty * 4];
      const float4 b = *(const float4*)&B[k][tx * 4];
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* out = partial + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * (kGT * kGT);
#pragma unroll
  for (int i = 0; i < 4; ++i)
    *(float4*)&out[(ty * 4 + i) * kGT + tx * 4] = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
}

// Sum the split-K partials in split order and mirror the upper triangle.
__global__ void __launch_bounds__(256)
    gram_reduce_kernel(const float* __restrict__ partial, int npairs, int nsplit, int nt, int d,
                       float* __restrict__ G) {
  const int pair_id = blockIdx.x;
  int pair = pair_id, ti = 0;
  while (pair >= nt - ti) { pair -= nt - ti; ++ti; }
  const int tj = ti + pair;
  for (int e = threadIdx.x; e < kGT * kGT; e += blockDim.x) {
    float s = 0.f;
    for (int sp = 0; sp < nsplit; ++sp) s += partial[((size_t)sp * npairs + pair_id) * (kGT * kGT) + e];
    const int gi = ti * kGT + e / kGT, gj = tj * kGT + e % kGT;
    if (gi < d && gj < d) {
      G[(size_t)gi * d + gj] = s;
      G[(size_t)gj * d + gi] = s;
    }
  }
}

static void fp32_plan(int64_t N, int d, int* nt, int* npairs, int* nsplit, int64_t* rows_per_split) {
  *nt = (d + kGT - 1) / kGT;
  *npairs = *nt * (*nt + 1) / 2;
  int64_t chunks = ceil_div(N > 0 ? N : 1, kGK);
  int64_t want = ceil_div((int64_t)sm_count() * 2, *npairs);
  int64_t s = chunks < want ? chunks : want;
  if (s < 1) s = 1;
  int64_t cps = ceil_div(chunks, s);  // chunks per split
  *rows_per_split = cps * kGK;
  *nsplit = (int)ceil_div(chunks, cps);
}

size_t gram_fp32_workspace_bytes(int64_t N, int d) {
  int nt, npairs, nsplit;
  int64_t rps;
  fp32_plan(N, d, &nt, &npairs, &nsplit, &rps);
  return (size_t)nsplit * npairs * kGT * kGT * sizeof(float);
}

int gram_fp32(const float* X, int64_t N, int d, float* G, void* ws, size_t ws_bytes, cudaStream_t st) {
  int nt, npairs, nsplit;
  int64_t rps;
  fp32_plan(N, d, &nt, &npairs, &nsplit, &rps);
  MMB_REQUIRE(ws_bytes >= gram_fp32_workspace_bytes(N, d), "workspace too small");
  gram_fp32_kernel<<<dim3(npairs, nsplit), kGThreads, 0, st>>>(X, N, d, nt, rps, (float*)ws);
  MMB_LAUNCH_CHECK("gram_fp32");
  gram_reduce_kernel<<<npairs, 256, 0, st>>>((const float*)ws, npairs, nsplit, nt, d, G);
  MMB_LAUNCH_CHECK("gram_reduce");
  return MMB_OK;
}

}  // namespace mmb

// MMB likelihood step:
//   * generator heads   -- reference models.py:187-202  (mu = z W^T + b, sigma = exp(z W'^T + b'))
//   * masked Gaussian    -- reference losses.py:13-34    (get_normal_log_prob), value + gradients
//   * angular word term  -- reference losses.py:68-95    (get_word_log_prob_angular2), value + d/d latents
// One step works on a minibatch of B = 64 utterances (512 for valid/test), so every kernel
// here is launch/latency bound rather than bandwidth bound; the design goal is FEW launches
// that read each input once (all heads in one launch, all modalities in one launch, no
// torch.cat of the base tensors, no (B, V, d) broadcast temporaries), FP32 throughout.
#include <math.h>

#include "common.cuh"

namespace mmb {

constexpr int kMaxHeads = 16;
constexpr int kMaxMods = 8;
constexpr int kMaxSegs = 4;

// --------------------------------------------------------------------------------------
// 64x64 output tile of C = A * B with arbitrary element strides, FP32 FMA, 256 threads,
// 4x4 register block per thread.  A(m,k) = A[m*a_rs + k*a_cs], B(k,n) = Bm[k*b_rs + n*b_cs].
// Operand tiles go through shared memory transposed to [k][m] / [k][n] so that the inner
// product reads two float4 per k.  Small-matrix helper: the problems here are <= 512 x 3016 x 425.
// --------------------------------------------------------------------------------------
constexpr int kTM = 64, kTN = 64, kTK = 32;
constexpr int kStageElems = kTM * kTK / 256;   // operand elements each thread stages per k-tile

struct TileSmem {
  float a[kTK][kTM + 4];
  float b[kTK][kTN + 4];
};

// The problems are tiny (one to a few CTAs per SM, K of a few hundred), so a k-tile's global loads are pure
// exposed latency unless they are in flight while the previous tile is multiplied: the next tile is
// prefetched into registers right after the barrier that publishes the current one (software pipeline,
// one stage).  The k order of every output element is unchanged (0 .. K-1), so results are bit-identical
// to the unpipelined loop.
__device__ __forceinline__ void tile_gemm_accum(TileSmem& sm, float (&acc)[4][4], const float* __restrict__ A,
                                                int64_t a_rs, int64_t a_cs, const float* __restrict__ Bm,
                                                int64_t b_rs, int64_t b_cs, int m0, int n0, int M, int N, int K) {
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  // make the fastest-varying thread index follow the smaller stride of each operand
  const bool a_kfast = a_cs <= a_rs, b_kfast = b_rs <= b_cs;
  float ra[kStageElems], rb[kStageElems];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int e = 0; e < kStageElems; ++e) {
      const int idx = tid + 256 * e;
      const int mm = a_kfast ? idx / kTK : idx % kTM, ka = a_kfast ? idx % kTK : idx / kTM;
      const int nn = b_kfast ? idx / kTK : idx % kTN, kb = b_kfast ? idx % kTK : idx / kTN;
      const int m = m0 + mm, n = n0 + nn;
      ra[e] = (m < M && k0 + ka < K) ? __ldg(A + m * a_rs + (k0 + ka) * a_cs) : 0.f;
      rb[e] = (n < N && k0 + kb < K) ? __ldg(Bm + (k0 + kb) * b_rs + n * b_cs) : 0.f;
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < K; k0 += kTK) {
#pragma unroll
    for (int e = 0; e < kStageElems; ++e) {
      const int idx = tid + 256 * e;
      const int mm = a_kfast ? idx / kTK : idx % kTM, ka = a_kfast ? idx % kTK : idx / kTM;
      const int nn = b_kfast ? idx / kTK : idx % kTN, kb = b_kfast ? idx % kTK : idx / kTN;
      sm.a[ka][mm] = ra[e];
      sm.b[kb][nn] = rb[e];
    }
    __syncthreads();
    if (k0 + kTK < K) fetch(k0 + kTK);
#pragma unroll
    for (int kk = 0; kk < kTK; ++kk) {
      const float4 a = *(const float4*)&sm.a[kk][ty * 4];
      const float4 b = *(const float4*)&sm.b[kk][tx * 4];
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
}

__device__ __forceinline__ void zero_acc(float (&acc)[4][4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
}

// out[i] = sum_s part[s][i] in split order (deterministic): second stage of the split-K products
// below, which spread a (64 x K) x (K x d) product with K in the thousands over ~100 CTAs instead
// of d / 64 = 5.
__global__ void __launch_bounds__(256)
    splitk_reduce_kernel(const float* __restrict__ part, int n_split, int n, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int k = 0; k < n_split; ++k) s += part[(size_t)k * n + i];
  out[i] = s;
}

constexpr int kMaxDzItems = 96;
struct DzItems {          // split-K work items of heads_bwd_dz: rows [k0, k0 + klen) of head h
  unsigned char h[kMaxDzItems];
  short k0[kMaxDzItems];
  short klen[kMaxDzItems];
  int n_items;
};

// ---- multi-tensor helpers: the step is launch-bound, so chains of per-modality element-wise
// ---- torch ops (12 + 6 multiplies in the backward pass, 8 batch gathers) become one launch each.
constexpr int kMaxMulti = 16;
struct ScaleArgs {
  const float* in[kMaxMulti];     // (B, D[i])
  const float* row[kMaxMulti];    // (B) row scale or null
  const float* elem[kMaxMulti];   // (B, D[i]) element-wise factor or null
  float* out[kMaxMulti];
  int D[kMaxMulti];
  int n;
};
// out[i][b][f] = in[i][b][f] * (row[i] ? row[i][b] : 1) * (elem[i] ? elem[i][b][f] : 1); grid (B, n)
__global__ void __launch_bounds__(128) scale_multi_kernel(const __grid_constant__ ScaleArgs a) {
  const int b = blockIdx.x, i = blockIdx.y, D = a.D[i];
  const float r = a.row[i] ? __ldg(a.row[i] + b) : 1.f;
  const float* in = a.in[i] + (size_t)b * D;
  const float* el = a.elem[i] ? a.elem[i] + (size_t)b * D : nullptr;
  float* out = a.out[i] + (size_t)b * D;
  for (int f = threadIdx.x; f < D; f += blockDim.x) out[f] = in[f] * r * (el ? __ldg(el + f) : 1.f);
}

// The combination step of get_log_prob_matrix (reference losses.py:267-272) in one launch each way:
//   out[b]      = other_w * sum_m lp[m][b] + word_w * wlp[b]
//   g_lp[m][b]  = other_w * g[b],   g_wlp[b] = word_w * g[b]
// (torch: sum, two multiplies, an add and their four backward nodes.)  other_w / word_w may come from
// device scalars (ow_dev / ww_dev non-NULL): a captured graph then serves every grid point's weights.
__global__ void __launch_bounds__(256)
    combine_lp_kernel(const float* __restrict__ lp, const float* __restrict__ wlp, int M, int B, float ow, float ww,
                      const float* __restrict__ ow_dev, const float* __restrict__ ww_dev, float* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  if (ow_dev) ow = __ldg(ow_dev);
  if (ww_dev) ww = __ldg(ww_dev);
  float s = 0.f;
  for (int m = 0; m < M; ++m) s += lp[(size_t)m * B + b];      // modality order, as torch's sum(0)
  out[b] = __fadd_rn(__fmul_rn(s, ow), __fmul_rn(ww, wlp[b]));   // torch's two products and their sum, unfused
}
__global__ void __launch_bounds__(256)
    combine_lp_bwd_kernel(const float* __restrict__ g, int M, int B, float ow, float ww,
                          const float* __restrict__ ow_dev, const float* __restrict__ ww_dev,
                          float* __restrict__ g_lp, float* __restrict__ g_wlp) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  if (ow_dev) ow = __ldg(ow_dev);
  if (ww_dev) ww = __ldg(ww_dev);
  const float gb = g[b];
  for (int m = 0; m < M; ++m) g_lp[(size_t)m * B + b] = ow * gb;
  g_wlp[b] = ww * gb;
}

struct GatherArgs {
  const float* src[kMaxMulti];    // (N_i, W[i]) row-major
  float* dst[kMaxMulti];          // (B, W[i])
  int64_t W[kMaxMulti];           // row width in floats (trailing dims flattened)
  int n;
};
// dst[i][b][:] = src[i][idx[b]][:]; grid (B, n)
__global__ void __launch_bounds__(256)
    gather_multi_kernel(const __grid_constant__ GatherArgs a, const int64_t* __restrict__ idx) {
  const int b = blockIdx.x, i = blockIdx.y;
  const int64_t W = a.W[i];
  const float* s = a.src[i] + (size_t)idx[b] * W;
  float* d = a.dst[i] + (size_t)b * W;
  if ((W & 3) == 0 && (((uintptr_t)s | (uintptr_t)d) & 15) == 0) {
    for (int64_t k = threadIdx.x; k < (W >> 2); k += blockDim.x) ((float4*)d)[k] = __ldg((const float4*)s + k);
  } else {
    for (int64_t k = threadIdx.x; k < W; k += blockDim.x) d[k] = __ldg(s + k);
  }
}

// ---------------------------------------------------------------- heads (A6) ---------
struct HeadsArgs {
  const float* W[kMaxHeads];
  const float* b[kMaxHeads];
  float* out[kMaxHeads];
  const float* gout[kMaxHeads];
  float* dW[kMaxHeads];
  float* db[kMaxHeads];
  int D[kMaxHeads];
  int is_ls[kMaxHeads];
  int n_heads;
};

// out[h][m][n] = act(sum_k z[m][k] W[h][n][k] + b[h][n]); grid (n tiles, m tiles, heads)
__global__ void __launch_bounds__(256)
    heads_fwd_kernel(const float* __restrict__ z, int B, int d, const __grid_constant__ HeadsArgs args) {
  __shared__ TileSmem sm;
  const int h = blockIdx.z, D = args.D[h];
  const int n0 = blockIdx.x * kTN, m0 = blockIdx.y * kTM;
  if (n0 >= D) return;
  float acc[4][4];
  zero_acc(acc);
  tile_gemm_accum(sm, acc, z, d, 1, args.W[h], 1, d, m0, n0, B, D, d);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const bool ls = args.is_ls[h] != 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= B) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= D) continue;
      float v = acc[i][j] + __ldg(args.b[h] + n);
      args.out[h][(size_t)m * D + n] = ls ? expf(v) : v;
    }
  }
}

// dz[m][n] = sum_h sum_k gout[h][m][k] W[h][k][n]; grid (n tiles over d, m tiles, split-K items):
// item z handles rows [k0, k0 + klen) of head h and writes its partial tile to part[z][m][n].
__global__ void __launch_bounds__(256)
    heads_bwd_dz_kernel(float* __restrict__ part, int B, int d, const __grid_constant__ HeadsArgs args,
                        const __grid_constant__ DzItems items) {
  __shared__ TileSmem sm;
  const int n0 = blockIdx.x * kTN, m0 = blockIdx.y * kTM;
  const int it = blockIdx.z, h = items.h[it], k0 = items.k0[it], klen = items.klen[it];
  float acc[4][4];
  zero_acc(acc);
  tile_gemm_accum(sm, acc, args.gout[h] + k0, args.D[h], 1, args.W[h] + (size_t)k0 * d, d, 1, m0, n0, B, d, klen);
  float* out = part + (size_t)it * B * d;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= B) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < d) out[(size_t)m * d + n] = acc[i][j];
    }
  }
}

// dW[h][m][n] = sum_k gout[h][k][m] z[k][n]  (m over D[h], n over d, k over B); db[h][m] = sum_k gout[h][k][m]
__global__ void __launch_bounds__(256)
    heads_bwd_dw_kernel(const float* __restrict__ z, int B, int d, const __grid_constant__ HeadsArgs args) {
  __shared__ TileSmem sm;
  const int h = blockIdx.z, D = args.D[h];
  const int n0 = blockIdx.x * kTN, m0 = blockIdx.y * kTM;
  if (m0 >= D) return;
  float acc[4][4];
  zero_acc(acc);
  tile_gemm_accum(sm, acc, args.gout[h], 1, D, z, d, 1, m0, n0, D, d, B);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= D) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < d) args.dW[h][(size_t)m * d + n] = acc[i][j];
    }
  }
  if (blockIdx.x == 0 && args.db[h]) {   // bias gradient: column sums of gout[h], fixed order
    for (int m = m0 + threadIdx.x; m < D && m < m0 + kTM; m += 256) {
      float s = 0.f;
      for (int k = 0; k < B; ++k) s += args.gout[h][(size_t)k * D + m];
      args.db[h][m] = s;
    }
  }
}

// ---------------------------------------------------------------- Gaussian (A7) ------
struct GaussArgs {
  // modality m is the concatenation (like torch.cat(dim=-1)) of n_seg[m] segments
  const float* val[kMaxMods][kMaxSegs];
  const float* msk[kMaxMods][kMaxSegs];
  int F[kMaxMods][kMaxSegs];
  int n_seg[kMaxMods];
  const float* mu[kMaxMods];
  const float* sigma[kMaxMods];
  float* dmu[kMaxMods];
  float* dsigma[kMaxMods];
  int D[kMaxMods];
  int n_mod;
};

// One CTA per (utterance b, modality m).  Thread f walks the T time steps of feature f:
//   S0 = sum m, S1 = sum m (x - mu), S2 = sum m (x - mu)^2          (coalesced across f)
//   lp_f = -0.5 log(2 pi sigma^2) S0 - S2 / (2 sigma^2)
//   d lp/d mu = S1 / sigma^2,   d lp/d sigma = -S0 / sigma + S2 / sigma^3
__global__ void __launch_bounds__(128)
    gauss_ll_kernel(const __grid_constant__ GaussArgs args, int B, int T, float* __restrict__ lp,
                    int* __restrict__ status) {
  __shared__ float red[4];
  const int b = blockIdx.x, m = blockIdx.y;
  const int D = args.D[m];
  float lp_acc = 0.f;
  for (int f = threadIdx.x; f < D; f += blockDim.x) {
    int seg = 0, fl = f;
    while (seg + 1 < args.n_seg[m] && fl >= args.F[m][seg]) { fl -= args.F[m][seg]; ++seg; }
    const int F = args.F[m][seg];
    const float* x = args.val[m][seg] + (size_t)b * T * F + fl;
    const float* k = args.msk[m][seg] + (size_t)b * T * F + fl;
    const float mu = __ldg(args.mu[m] + (size_t)b * D + f);
    const float sg = __ldg(args.sigma[m] + (size_t)b * D + f);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    int t = 0;
    for (; t + 4 <= T; t += 4) {   // 8 independent loads in flight, then the (ordered) sums
      float mk[4], xv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        mk[u] = __ldg(k + (size_t)(t + u) * F);
        xv[u] = __ldg(x + (size_t)(t + u) * F);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float df = xv[u] - mu;
        s0 += mk[u];
        s1 = fmaf(mk[u], df, s1);
        s2 = fmaf(mk[u] * df, df, s2);
      }
    }
    for (; t < T; ++t) {
      const float mk = __ldg(k + (size_t)t * F);
      const float df = __ldg(x + (size_t)t * F) - mu;
      s0 += mk;
      s1 = fmaf(mk, df, s1);
      s2 = fmaf(mk * df, df, s2);
    }
    const float var = sg * sg;
    const float inv_var = 1.f / var;
    lp_acc += -0.5f * logf(6.283185307179586f * var) * s0 - 0.5f * s2 * inv_var;
    if (args.dmu[m]) args.dmu[m][(size_t)b * D + f] = s1 * inv_var;
    if (args.dsigma[m]) args.dsigma[m][(size_t)b * D + f] = (s2 * inv_var - s0) / sg;
  }
  lp_acc = warp_sum(lp_acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = lp_acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    const float s = red[0] + red[1] + red[2] + red[3];
    lp[(size_t)m * B + b] = s;
    if (!isfinite(s)) atomicOr(status, MMB_STATUS_NONFINITE);
  }
}

// ---- the same likelihood from per-utterance moments -------
// The data of an utterance never changes between optimisation steps; only mu and sigma do.  With the
// masked moments over time of each base feature,
//     S0 = sum_t m,   mean = sum_t m x / S0,   M2 = sum_t m (x - mean)^2        (weights m = the float mask)
// the masked squared distance to ANY mu is  sum_t m (x - mu)^2 = M2 + S0 (mean - mu)^2  (exact algebra, and
// in this centred form free of the cancellation the raw-moment form S2 - 2 mu S1 + mu^2 S0 suffers), so
//     lp_f = -0.5 log(2 pi sigma^2) S0 - (M2 + S0 (mean - mu)^2) / (2 sigma^2)
//     d lp/d mu = S0 (mean - mu) / sigma^2,   d lp/d sigma = ((M2 + S0 (mean - mu)^2) / sigma^2 - S0) / sigma.
// mmb_gauss_moments runs ONCE per dataset; every step then reads 3 floats per (utterance, feature)
// instead of 2 T, and the (B, T, F) batch gathers of values and masks disappear from the step.
//
// stats layout: (N, 3, F) -- [S0 | mean | M2] rows of F floats per utterance.
__global__ void __launch_bounds__(128)
    gauss_moments_kernel(const float* __restrict__ val, const float* __restrict__ msk, int T, int F,
                         float* __restrict__ stats) {
  const int n = blockIdx.x;
  const int f = blockIdx.y * blockDim.x + threadIdx.x;
  if (f >= F) return;
  const float* x = val + (size_t)n * T * F + f;
  const float* k = msk + (size_t)n * T * F + f;
  float s0 = 0.f, s1 = 0.f;
  for (int t = 0; t < T; ++t) {
    const float mk = __ldg(k + (size_t)t * F);
    s0 += mk;
    s1 = fmaf(mk, __ldg(x + (size_t)t * F), s1);
  }
  const float mean = s0 != 0.f ? s1 / s0 : 0.f;
  float m2 = 0.f;
  for (int t = 0; t < T; ++t) {             // second pass over lines this thread just pulled into L1/L2
    const float mk = __ldg(k + (size_t)t * F);
    const float df = __ldg(x + (size_t)t * F) - mean;
    m2 = fmaf(mk * df, df, m2);
  }
  float* o = stats + (size_t)n * 3 * F + f;
  o[0] = s0;
  o[F] = mean;
  o[2 * F] = m2;
}

struct GaussStatArgs {
  const float* st[kMaxMods][kMaxSegs];   // (B, 3, F) moments of the batch rows
  int F[kMaxMods][kMaxSegs];
  int n_seg[kMaxMods];
  const float* mu[kMaxMods];
  const float* sigma[kMaxMods];
  float* dmu[kMaxMods];
  float* dsigma[kMaxMods];
  int D[kMaxMods];
  int n_mod;
};

__global__ void __launch_bounds__(128)
    gauss_ll_stats_kernel(const __grid_constant__ GaussStatArgs args, int B, float* __restrict__ lp,
                          int* __restrict__ status) {
  __shared__ float red[4];
  const int b = blockIdx.x, m = blockIdx.y;
  const int D = args.D[m];
  float lp_acc = 0.f;
  for (int f = threadIdx.x; f < D; f += blockDim.x) {
    int seg = 0, fl = f;
    while (seg + 1 < args.n_seg[m] && fl >= args.F[m][seg]) { fl -= args.F[m][seg]; ++seg; }
    const int F = args.F[m][seg];
    const float* st = args.st[m][seg] + (size_t)b * 3 * F + fl;
    const float s0 = __ldg(st), mean = __ldg(st + F), m2 = __ldg(st + 2 * F);
    const float mu = __ldg(args.mu[m] + (size_t)b * D + f);
    const float sg = __ldg(args.sigma[m] + (size_t)b * D + f);
    const float df = mean - mu;
    const float q = fmaf(s0 * df, df, m2);
    const float var = sg * sg;
    const float inv_var = 1.f / var;
    lp_acc += -0.5f * logf(6.283185307179586f * var) * s0 - 0.5f * q * inv_var;
    if (args.dmu[m]) args.dmu[m][(size_t)b * D + f] = s0 * df * inv_var;
    if (args.dsigma[m]) args.dsigma[m][(size_t)b * D + f] = (q * inv_var - s0) / sg;
  }
  lp_acc = warp_sum(lp_acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = lp_acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    const float s = red[0] + red[1] + red[2] + red[3];
    lp[(size_t)m * B + b] = s;
    if (!isfinite(s)) atomicOr(status, MMB_STATUS_NONFINITE);
  }
}

// ---------------------------------------------------------------- word term (A8) -----
__global__ void row_inv_norm_kernel(const float* __restrict__ X, int64_t n, int d, float* __restrict__ out) {
  // one warp per row: 1 / max(||x||, 1e-8)   (torch CosineSimilarity eps)
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  float s = 0.f;
  for (int k = lane; k < d; k += 32) {
    const float v = __ldg(X + row * d + k);
    s = fmaf(v, v, s);
  }
  s = warp_sum(s);
  if (lane == 0) out[row] = 1.f / fmaxf(sqrtf(s), 1e-8f);
}

constexpr float kInvPi = 0.3183098861837907f;

// C tile = (e . w_v) * ie_b * iw_v;  writes S[b][v] = 1 - acos(c)/pi and
// Hc[b][v] = h(c) * iw_v with h(c) = 1 / (pi sqrt(1 - c^2)), Q[b][v] = h(c) * c.
__global__ void __launch_bounds__(256)
    word_cos_kernel(const float* __restrict__ e, const float* __restrict__ ie, int B, int d,
                    const float* __restrict__ W, const float* __restrict__ iw, int V, float* __restrict__ S,
                    float* __restrict__ Hw, float* __restrict__ Q, float* __restrict__ Cos = nullptr) {
  __shared__ TileSmem sm;
  const int n0 = blockIdx.x * kTN, m0 = blockIdx.y * kTM;
  float acc[4][4];
  zero_acc(acc);
  tile_gemm_accum(sm, acc, e, d, 1, W, 1, d, m0, n0, B, V, d);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= B) continue;
    const float iem = ie[m];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= V) continue;
      const float iwn = iw[n];
      const float c = acc[i][j] * iem * iwn;
      const float h = kInvPi / sqrtf(1.f - c * c);
      S[(size_t)m * V + n] = 1.f - acosf(c) * kInvPi;
      Hw[(size_t)m * V + n] = h * iwn;
      Q[(size_t)m * V + n] = h * c;
      if (Cos) Cos[(size_t)m * V + n] = c;
    }
  }
}

// Hsum[b][:] = sum_v Hw[b][v] * W[v][:]   ((B x V) x (V x d)), split over V: CTA z handles
// rows [z * vchunk, (z + 1) * vchunk) of the table and writes part[z][b][:].
__global__ void __launch_bounds__(256)
    word_h_kernel(const float* __restrict__ Hw, const float* __restrict__ W, int B, int V, int d, int vchunk,
                  float* __restrict__ part) {
  __shared__ TileSmem sm;
  const int n0 = blockIdx.x * kTN, m0 = blockIdx.y * kTM;
  const int v0 = blockIdx.z * vchunk;
  const int vlen = (V - v0) < vchunk ? (V - v0) : vchunk;
  float acc[4][4];
  zero_acc(acc);
  tile_gemm_accum(sm, acc, Hw + v0, V, 1, W + (size_t)v0 * d, d, 1, m0, n0, B, d, vlen);
  float* out = part + (size_t)blockIdx.z * B * d;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= B) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < d) out[(size_t)m * d + n] = acc[i][j];
    }
  }
}

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
  return s;
}

// One CTA (256 threads) per utterance: partition sum, per-token terms, lp and d lp / d e.
__global__ void __launch_bounds__(256)
    word_finish_kernel(const float* __restrict__ e, const float* __restrict__ ie, int B, int d, int V,
                       const float* __restrict__ S, const float* __restrict__ Q,
                       const float* __restrict__ Hsum, const float* __restrict__ sent, int64_t s_sb,
                       int64_t s_st, const float* __restrict__ word_w, const float* __restrict__ tmask,
                       int64_t m_sb, int64_t m_st, int L, float a, float* __restrict__ lp,
                       float* __restrict__ grad, int* __restrict__ status) {
  extern __shared__ float dyn[];   // L floats: r_t, then L floats: c_t
  __shared__ float red[8];
  __shared__ float bc[4];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* r_t = dyn;
  float* c_t = dyn + L;
  // Z = sum_v S, hc = sum_v Q   (fixed order per thread, then block tree)
  float z = 0.f, hc = 0.f;
  for (int v = tid; v < V; v += 256) {
    z += S[(size_t)b * V + v];
    hc += Q[(size_t)b * V + v];
  }
  const float Z = block_sum(z, red);
  const float HC = block_sum(hc, red);
  const float inv_e = ie[b];
  const float alpha = 1.f / (Z * a + 1.f);
  // per-token cosine: one warp per token
  for (int t = warp; t < L; t += 8) {
    const float* s = sent + (size_t)b * s_sb + (size_t)t * s_st;
    float dot = 0.f, nn = 0.f;
    for (int k = lane; k < d; k += 32) {
      const float sv = __ldg(s + k);
      dot = fmaf(sv, __ldg(e + (size_t)b * d + k), dot);
      nn = fmaf(sv, sv, nn);
    }
    dot = warp_sum(dot);
    nn = warp_sum(nn);
    if (lane == 0) {
      const float is = 1.f / fmaxf(sqrtf(nn), 1e-8f);
      c_t[t] = dot * inv_e * is;
      r_t[t] = is;   // temporarily the token's inverse norm
    }
  }
  __syncthreads();
  float lp_acc = 0.f, dz_acc = 0.f, rc_acc = 0.f;
  for (int t = tid; t < L; t += 256) {
    const float c = c_t[t], is = r_t[t];
    const float st = 1.f - acosf(c) * kInvPi;
    const float w = word_w[(size_t)b * L + t];
    const float mk = tmask[(size_t)b * m_sb + (size_t)t * m_st];
    const float p = alpha * w + (1.f - alpha) * st / Z;
    lp_acc += logf(p) * mk;
    const float dlp_dp = mk / p;
    const float dp_dZ = (-a * alpha * alpha) * (w - st / Z) - (1.f - alpha) * st / (Z * Z);
    dz_acc += dlp_dp * dp_dZ;
    const float rr = dlp_dp * (1.f - alpha) / Z * (kInvPi / sqrtf(1.f - c * c));
    rc_acc += rr * c;
    r_t[t] = rr * is;       // weight of the raw token vector in the gradient
  }
  const float LP = block_sum(lp_acc, red);
  const float DZ = block_sum(dz_acc, red);
  const float RC = block_sum(rc_acc, red);
  if (tid == 0) {
    lp[b] = LP;
    if (!isfinite(LP)) atomicOr(status, MMB_STATUS_NONFINITE);
    bc[0] = DZ;
  }
  __syncthreads();
  // d lp/d e = [ DZ * (Hsum - HC * ehat) + sum_t r_t shat_t - RC * ehat ] / ||e||
  for (int k = tid; k < d; k += 256) {
    const float eh = e[(size_t)b * d + k] * inv_e;
    float g = DZ * (Hsum[(size_t)b * d + k] - HC * eh) - RC * eh;
    float tok = 0.f;
    for (int t = 0; t < L; ++t) tok = fmaf(r_t[t], __ldg(sent + (size_t)b * s_sb + (size_t)t * s_st + k), tok);
    grad[(size_t)b * d + k] = (g + tok) * inv_e;
  }
}

// ---- word term from token IDS -------------------------------------------------
// When the token vectors are rows of the word table itself (reference simplesif.py:319-340 builds
// `text = word_embeddings[ids]`), everything the per-token part needs is already in the (B, V) matrices
// of the partition term: the token's cosine is Cos[b][id], its inverse norm is iw[id], and its share of
// the gradient, r_t * w_hat_{id}, is one more entry of the (B, V) coefficient matrix that multiplies the
// table in the gradient product.  No (B, L, d) tensor is read or even exists; the cost no longer
// depends on L * d (POM: L = 1357).
//
// One CTA per utterance.  M[b][v] = DZ_b * Hw[b][v] + sum_{t: id_t = v} r_t * iw[v] is assembled in shared
// memory; the token contributions are added by ONE warp in token order, duplicates inside a 32-token
// chunk merged with match.any and a fixed butterfly -> deterministic.  aux[b] = DZ_b * HC_b + RC_b.
__global__ void __launch_bounds__(256)
    word_token_ids_kernel(int B, int V, const float* __restrict__ S, const float* __restrict__ Q,
                          const float* __restrict__ Cos, float* __restrict__ HwM, const float* __restrict__ iw,
                          const int64_t* __restrict__ ids, int64_t ids_sb, const float* __restrict__ word_w,
                          const float* __restrict__ tmask, int64_t m_sb, int64_t m_st, int L, float a,
                          float* __restrict__ lp, float* __restrict__ aux, int* __restrict__ status) {
  extern __shared__ float dyn[];   // V floats: row of M; then L floats: r_t * iw
  __shared__ float red[8];
  float* Mrow = dyn;
  float* r_t = dyn + V;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float z = 0.f, hc = 0.f;
  for (int v = tid; v < V; v += 256) {
    z += S[(size_t)b * V + v];
    hc += Q[(size_t)b * V + v];
  }
  const float Z = block_sum(z, red);
  const float HC = block_sum(hc, red);
  const float alpha = 1.f / (Z * a + 1.f);
  float lp_acc = 0.f, dz_acc = 0.f, rc_acc = 0.f;
  bool bad = false;
  for (int t = tid; t < L; t += 256) {
    const int64_t id = ids[(size_t)b * ids_sb + t];
    float rw = 0.f;
    if (id >= 0 && id < V) {
      const float c = Cos[(size_t)b * V + id];
      const float st = S[(size_t)b * V + id];
      const float w = word_w[(size_t)b * L + t];
      const float mk = tmask ? tmask[(size_t)b * m_sb + (size_t)t * m_st] : (id != 0 ? 1.f : 0.f);
      const float p = alpha * w + (1.f - alpha) * st / Z;
      lp_acc += logf(p) * mk;
      const float dlp_dp = mk / p;
      const float dp_dZ = (-a * alpha * alpha) * (w - st / Z) - (1.f - alpha) * st / (Z * Z);
      dz_acc += dlp_dp * dp_dZ;
      const float rr = dlp_dp * (1.f - alpha) / Z * (kInvPi / sqrtf(1.f - c * c));
      rc_acc += rr * c;
      rw = rr * iw[id];
    } else {
      bad = true;
    }
    r_t[t] = rw;
  }
  const float LP = block_sum(lp_acc, red);
  const float DZ = block_sum(dz_acc, red);
  const float RC = block_sum(rc_acc, red);
  for (int v = tid; v < V; v += 256) Mrow[v] = DZ * HwM[(size_t)b * V + v];
  __syncthreads();
  if (warp == 0) {
    for (int t0 = 0; t0 < L; t0 += 32) {
      const int t = t0 + lane;
      int64_t id = -1;
      float val = 0.f;
      if (t < L) {
        id = ids[(size_t)b * ids_sb + t];
        if (id < 0 || id >= V) id = -1;
        else val = r_t[t];
      }
      const int key = (int)id;
      const unsigned grp = __match_any_sync(0xffffffffu, key);
      const bool head = key >= 0 && lane == __ffs(grp) - 1;
      unsigned multi = __ballot_sync(0xffffffffu, head && (grp & (grp - 1u)));
      while (multi) {
        const int j = __ffs(multi) - 1;
        multi &= multi - 1u;
        const unsigned gj = __shfl_sync(0xffffffffu, grp, j);
        const float sum = warp_sum(((gj >> lane) & 1u) ? val : 0.f);
        if (lane == j) val = sum;
      }
      if (head) Mrow[key] += val;
      __syncwarp();
    }
  }
  __syncthreads();
  for (int v = tid; v < V; v += 256) HwM[(size_t)b * V + v] = Mrow[v];
  if (tid == 0) {
    lp[b] = LP;
    aux[b] = DZ * HC + RC;
    if (!isfinite(LP)) atomicOr(status, MMB_STATUS_NONFINITE);
  }
  if (bad) atomicOr(status, MMB_STATUS_BAD_INDEX);
}

// grad[b][k] = (G[b][k] - aux[b] * e[b][k] * ie[b]) * ie[b]
__global__ void __launch_bounds__(256)
    word_grad_finish_kernel(const float* __restrict__ G, const float* __restrict__ e, const float* __restrict__ ie,
                            const float* __restrict__ aux, int B, int d, float* __restrict__ grad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * d) return;
  const int b = i / d;
  const float inv_e = ie[b];
  grad[i] = (G[i] - aux[b] * e[i] * inv_e) * inv_e;
}

}  // namespace mmb

using namespace mmb;

extern "C" int mmb_heads_forward(const float* z, int B, int d, int n_heads, const float* const* W,
                                 const float* const* b, const int* D, const int* is_log_sigma,
                                 float* const* out, mmb_stream_t stream) {
  MMB_REQUIRE(z && W && b && D && is_log_sigma && out, "null pointer");
  MMB_REQUIRE(n_heads > 0 && n_heads <= kMaxHeads, "1..16 heads");
  MMB_REQUIRE(B > 0 && d > 0, "bad size");
  HeadsArgs a = {};
  int max_D = 0;
  for (int h = 0; h < n_heads; ++h) {
    MMB_REQUIRE(W[h] && b[h] && out[h] && D[h] > 0, "null head");
    a.W[h] = W[h]; a.b[h] = b[h]; a.out[h] = out[h]; a.D[h] = D[h]; a.is_ls[h] = is_log_sigma[h];
    if (D[h] > max_D) max_D = D[h];
  }
  a.n_heads = n_heads;
  dim3 grid((max_D + kTN - 1) / kTN, (B + kTM - 1) / kTM, n_heads);
  heads_fwd_kernel<<<grid, 256, 0, as_stream(stream)>>>(z, B, d, a);
  MMB_LAUNCH_CHECK("heads_fwd");
  return MMB_OK;
}

static const int kDzChunk = 128;   // rows of W per split-K item

static int dz_items(int n_heads, const int* D, DzItems* it) {
  int n = 0;
  for (int h = 0; h < n_heads; ++h)
    for (int k0 = 0; k0 < D[h]; k0 += kDzChunk) {
      if (n >= kMaxDzItems) return -1;
      if (it) {
        it->h[n] = (unsigned char)h;
        it->k0[n] = (short)k0;
        it->klen[n] = (short)((D[h] - k0) < kDzChunk ? (D[h] - k0) : kDzChunk);
      }
      ++n;
    }
  if (it) it->n_items = n;
  return n;
}

extern "C" size_t mmb_heads_backward_workspace_bytes(int B, int d, int n_heads, const int* D) {
  if (!D || n_heads <= 0) return 0;
  const int n = dz_items(n_heads, D, nullptr);
  return n > 0 ? (size_t)n * B * d * sizeof(float) : 0;
}

extern "C" int mmb_heads_backward(const float* z, int B, int d, int n_heads, const float* const* W,
                                  const int* D, const float* const* gout, float* dz, float* const* dW,
                                  float* const* db, void* ws, size_t ws_bytes, mmb_stream_t stream) {
  MMB_REQUIRE(z && W && D && gout, "null pointer");
  MMB_REQUIRE(n_heads > 0 && n_heads <= kMaxHeads, "1..16 heads");
  MMB_REQUIRE(B > 0 && d > 0, "bad size");
  HeadsArgs a = {};
  int max_D = 0;
  for (int h = 0; h < n_heads; ++h) {
    MMB_REQUIRE(W[h] && gout[h] && D[h] > 0 && D[h] < 32768, "null head / head too wide");
    a.W[h] = W[h]; a.gout[h] = gout[h]; a.D[h] = D[h];
    a.dW[h] = dW ? dW[h] : nullptr;
    a.db[h] = db ? db[h] : nullptr;
    if (D[h] > max_D) max_D = D[h];
  }
  a.n_heads = n_heads;
  cudaStream_t st = as_stream(stream);
  if (dz) {
    DzItems items = {};
    MMB_REQUIRE(dz_items(n_heads, D, &items) > 0, "too many split-K items (sum of ceil(D / 128) must be <= 96)");
    MMB_REQUIRE(ws && ws_bytes >= mmb_heads_backward_workspace_bytes(B, d, n_heads, D), "workspace too small");
    dim3 grid((d + kTN - 1) / kTN, (B + kTM - 1) / kTM, items.n_items);
    heads_bwd_dz_kernel<<<grid, 256, 0, st>>>((float*)ws, B, d, a, items);
    MMB_LAUNCH_CHECK("heads_bwd_dz");
    splitk_reduce_kernel<<<(B * d + 255) / 256, 256, 0, st>>>((const float*)ws, items.n_items, B * d, dz);
    MMB_LAUNCH_CHECK("heads_bwd_dz_reduce");
  }
  if (dW) {
    for (int h = 0; h < n_heads; ++h) MMB_REQUIRE(dW[h], "null dW");
    dim3 grid((d + kTN - 1) / kTN, (max_D + kTM - 1) / kTM, n_heads);
    heads_bwd_dw_kernel<<<grid, 256, 0, st>>>(z, B, d, a);
    MMB_LAUNCH_CHECK("heads_bwd_dw");
  }
  return MMB_OK;
}

extern "C" int mmb_scale_multi(int n, int B, const int* D, const float* const* in, const float* const* row,
                               const float* const* elem, float* const* out, mmb_stream_t stream) {
  MMB_REQUIRE(D && in && out, "null pointer");
  MMB_REQUIRE(n > 0 && n <= kMaxMulti && B > 0, "1..16 tensors");
  ScaleArgs a = {};
  for (int i = 0; i < n; ++i) {
    MMB_REQUIRE(in[i] && out[i] && D[i] > 0, "null tensor");
    a.in[i] = in[i]; a.out[i] = out[i]; a.D[i] = D[i];
    a.row[i] = row ? row[i] : nullptr;
    a.elem[i] = elem ? elem[i] : nullptr;
  }
  a.n = n;
  scale_multi_kernel<<<dim3(B, n), 128, 0, as_stream(stream)>>>(a);
  MMB_LAUNCH_CHECK("scale_multi");
  return MMB_OK;
}

extern "C" int mmb_combine_lp(const float* lp, const float* wlp, int M, int B, float other_w, float word_w,
                              const float* other_w_dev, const float* word_w_dev, float* out, mmb_stream_t stream) {
  MMB_REQUIRE(lp && wlp && out && M > 0 && B > 0, "bad argument");
  combine_lp_kernel<<<(B + 255) / 256, 256, 0, as_stream(stream)>>>(lp, wlp, M, B, other_w, word_w, other_w_dev,
                                                                     word_w_dev, out);
  MMB_LAUNCH_CHECK("combine_lp");
  return MMB_OK;
}

extern "C" int mmb_combine_lp_backward(const float* g, int M, int B, float other_w, float word_w,
                                       const float* other_w_dev, const float* word_w_dev, float* g_lp, float* g_wlp,
                                       mmb_stream_t stream) {
  MMB_REQUIRE(g && g_lp && g_wlp && M > 0 && B > 0, "bad argument");
  combine_lp_bwd_kernel<<<(B + 255) / 256, 256, 0, as_stream(stream)>>>(g, M, B, other_w, word_w, other_w_dev,
                                                                         word_w_dev, g_lp, g_wlp);
  MMB_LAUNCH_CHECK("combine_lp_bwd");
  return MMB_OK;
}

extern "C" int mmb_gather_multi(int n, int B, const int64_t* idx, const float* const* src, const int64_t* W,
                                float* const* dst, mmb_stream_t stream) {
  MMB_REQUIRE(idx && src && W && dst, "null pointer");
  MMB_REQUIRE(n > 0 && n <= kMaxMulti && B > 0, "1..16 tensors");
  GatherArgs a = {};
  for (int i = 0; i < n; ++i) {
    MMB_REQUIRE(src[i] && dst[i] && W[i] > 0, "null tensor");
    a.src[i] = src[i]; a.dst[i] = dst[i]; a.W[i] = W[i];
  }
  a.n = n;
  gather_multi_kernel<<<dim3(B, n), 256, 0, as_stream(stream)>>>(a, idx);
  MMB_LAUNCH_CHECK("gather_multi");
  return MMB_OK;
}

extern "C" int mmb_gauss_ll(int B, int T, int n_mod, const int* n_seg, const float* const* seg_val,
                            const float* const* seg_mask, const int* seg_F, const float* const* mu,
                            const float* const* sigma, float* lp, float* const* dmu, float* const* dsigma,
                            int* status, mmb_stream_t stream) {
  MMB_REQUIRE(n_seg && seg_val && seg_mask && seg_F && mu && sigma && lp && status, "null pointer");
  MMB_REQUIRE(n_mod > 0 && n_mod <= kMaxMods, "1..8 modalities");
  MMB_REQUIRE(B > 0 && T > 0, "bad size");
  GaussArgs a = {};
  int s = 0;
  for (int m = 0; m < n_mod; ++m) {
    MMB_REQUIRE(n_seg[m] > 0 && n_seg[m] <= kMaxSegs, "1..4 segments per modality");
    a.n_seg[m] = n_seg[m];
    int D = 0;
    for (int g = 0; g < n_seg[m]; ++g, ++s) {
      MMB_REQUIRE(seg_val[s] && seg_mask[s] && seg_F[s] > 0, "null segment");
      a.val[m][g] = seg_val[s]; a.msk[m][g] = seg_mask[s]; a.F[m][g] = seg_F[s];
      D += seg_F[s];
    }
    MMB_REQUIRE(mu[m] && sigma[m], "null mu/sigma");
    a.mu[m] = mu[m]; a.sigma[m] = sigma[m]; a.D[m] = D;
    a.dmu[m] = dmu ? dmu[m] : nullptr;
    a.dsigma[m] = dsigma ? dsigma[m] : nullptr;
  }
  a.n_mod = n_mod;
  gauss_ll_kernel<<<dim3(B, n_mod), 128, 0, as_stream(stream)>>>(a, B, T, lp, status);
  MMB_LAUNCH_CHECK("gauss_ll");
  return MMB_OK;
}

extern "C" int mmb_row_inv_norm(const float* X, int64_t n, int d, float* inv_norm, mmb_stream_t stream) {
  MMB_REQUIRE(X && inv_norm, "null pointer");
  MMB_REQUIRE(n >= 0 && d > 0, "bad size");
  if (n == 0) return MMB_OK;
  row_inv_norm_kernel<<<(unsigned)((n + 7) / 8), 256, 0, as_stream(stream)>>>(X, n, d, inv_norm);
  MMB_LAUNCH_CHECK("row_inv_norm");
  return MMB_OK;
}

static int word_h_splits(int64_t V) {   // ~192 table rows per split-K slice, at most 64 slices
  int64_t n = (V + 191) / 192;
  return (int)(n < 1 ? 1 : (n > 64 ? 64 : n));
}

extern "C" size_t mmb_word_ll_workspace_bytes(int B, int64_t V, int d) {
  // S, Hw, Q: (B, V) each; Hsum: (B, d); ie: (B); split-K partials of Hsum: (splits, B, d)
  return ((size_t)3 * B * V + (size_t)B * d + B + (size_t)word_h_splits(V) * B * d) * sizeof(float) + 256;
}

extern "C" int mmb_word_ll(const float* latents, int B, int d, const float* table, const float* inv_norm,
                           int64_t V, const float* sent, int64_t sent_stride_b, int64_t sent_stride_t,
                           const float* word_w, const float* tmask, int64_t tmask_stride_b,
                           int64_t tmask_stride_t, int L, float a, float* lp, float* grad, void* ws,
                           size_t ws_bytes, int* status, mmb_stream_t stream) {
  MMB_REQUIRE(latents && table && inv_norm && sent && word_w && tmask && lp && grad && ws && status,
              "null pointer");
  MMB_REQUIRE(B > 0 && d > 0 && V > 0 && L > 0 && V < (1 << 30), "bad size");
  MMB_REQUIRE(ws_bytes >= mmb_word_ll_workspace_bytes(B, V, d), "workspace too small");
  MMB_REQUIRE((size_t)2 * L * sizeof(float) <= 48 * 1024, "L too large for the per-utterance kernel");
  cudaStream_t st = as_stream(stream);
  float* S = (float*)ws;
  float* Hw = S + (size_t)B * V;
  float* Q = Hw + (size_t)B * V;
  float* Hsum = Q + (size_t)B * V;
  float* ie = Hsum + (size_t)B * d;
  float* hpart = ie + B;
  row_inv_norm_kernel<<<(B + 7) / 8, 256, 0, st>>>(latents, B, d, ie);
  MMB_LAUNCH_CHECK("row_inv_norm(latents)");
  dim3 g1((unsigned)((V + kTN - 1) / kTN), (B + kTM - 1) / kTM);
  word_cos_kernel<<<g1, 256, 0, st>>>(latents, ie, B, d, table, inv_norm, (int)V, S, Hw, Q);
  MMB_LAUNCH_CHECK("word_cos");
  const int splits = word_h_splits(V);
  const int vchunk = (int)(((V + splits - 1) / splits + kTK - 1) / kTK * kTK);
  dim3 g2((d + kTN - 1) / kTN, (B + kTM - 1) / kTM, (unsigned)((V + vchunk - 1) / vchunk));
  word_h_kernel<<<g2, 256, 0, st>>>(Hw, table, B, (int)V, d, vchunk, hpart);
  MMB_LAUNCH_CHECK("word_h");
  splitk_reduce_kernel<<<(B * d + 255) / 256, 256, 0, st>>>(hpart, (int)g2.z, B * d, Hsum);
  MMB_LAUNCH_CHECK("word_h_reduce");
  word_finish_kernel<<<B, 256, 2 * L * sizeof(float), st>>>(latents, ie, B, d, (int)V, S, Q, Hsum, sent,
                                                          sent_stride_b, sent_stride_t, word_w, tmask,
                                                          tmask_stride_b, tmask_stride_t, L, a, lp, grad, status);
  MMB_LAUNCH_CHECK("word_finish");
  return MMB_OK;
}

extern "C" size_t mmb_word_ll_ids_workspace_bytes(int B, int64_t V, int d) {
  // S, Hw/M, Q, Cos: (B, V) each; G: (B, d); ie, aux: (B) each; split-K partials of G: (splits, B, d)
  return ((size_t)4 * B * V + (size_t)B * d + 2 * (size_t)B + (size_t)word_h_splits(V) * B * d) * sizeof(float) + 256;
}

extern "C" int mmb_word_ll_ids(const float* latents, int B, int d, const float* table, const float* inv_norm,
                               int64_t V, const int64_t* ids, int64_t ids_stride_b, const float* word_w,
                               const float* tmask, int64_t tmask_stride_b, int64_t tmask_stride_t, int L,
                               float a, float* lp, float* grad, void* ws, size_t ws_bytes, int* status,
                               mmb_stream_t stream) {
  MMB_REQUIRE(latents && table && inv_norm && ids && word_w && lp && grad && ws && status, "null pointer");
  MMB_REQUIRE(B > 0 && d > 0 && V > 0 && L > 0 && V < (1 << 30), "bad size");
  MMB_REQUIRE(ws_bytes >= mmb_word_ll_ids_workspace_bytes(B, V, d), "workspace too small");
  const size_t smem = ((size_t)V + (size_t)L) * sizeof(float);
  if (smem > 200 * 1024) {
    set_error("mmb_word_ll_ids: vocabulary row + tokens (%zu bytes) exceed shared memory; use mmb_word_ll", smem);
    return MMB_E_UNSUPPORTED;
  }
  cudaStream_t st = as_stream(stream);
  float* S = (float*)ws;
  float* Hw = S + (size_t)B * V;
  float* Q = Hw + (size_t)B * V;
  float* Cos = Q + (size_t)B * V;
  float* G = Cos + (size_t)B * V;
  float* ie = G + (size_t)B * d;
  float* aux = ie + B;
  float* hpart = aux + B;
  row_inv_norm_kernel<<<(B + 7) / 8, 256, 0, st>>>(latents, B, d, ie);
  MMB_LAUNCH_CHECK("row_inv_norm(latents)");
  dim3 g1((unsigned)((V + kTN - 1) / kTN), (B + kTM - 1) / kTM);
  word_cos_kernel<<<g1, 256, 0, st>>>(latents, ie, B, d, table, inv_norm, (int)V, S, Hw, Q, Cos);
  MMB_LAUNCH_CHECK("word_cos");
  static bool attr_set = false;
  if (!attr_set) {
    MMB_CUDA(cudaFuncSetAttribute(word_token_ids_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  word_token_ids_kernel<<<B, 256, smem, st>>>(B, (int)V, S, Q, Cos, Hw, inv_norm, ids, ids_stride_b, word_w, tmask,
                                              tmask_stride_b, tmask_stride_t, L, a, lp, aux, status);
  MMB_LAUNCH_CHECK("word_token_ids");
  const int splits = word_h_splits(V);
  const int vchunk = (int)(((V + splits - 1) / splits + kTK - 1) / kTK * kTK);
  dim3 g2((d + kTN - 1) / kTN, (B + kTM - 1) / kTM, (unsigned)((V + vchunk - 1) / vchunk));
  word_h_kernel<<<g2, 256, 0, st>>>(Hw, table, B, (int)V, d, vchunk, hpart);
  MMB_LAUNCH_CHECK("word_h");
  splitk_reduce_kernel<<<(B * d + 255) / 256, 256, 0, st>>>(hpart, (int)g2.z, B * d, G);
  MMB_LAUNCH_CHECK("word_h_reduce");
  word_grad_finish_kernel<<<(B * d + 255) / 256, 256, 0, st>>>(G, latents, ie, aux, B, d, grad);
  MMB_LAUNCH_CHECK("word_grad_finish");
  return MMB_OK;
}

extern "C" int mmb_gauss_moments(const float* val, const float* mask, int64_t N, int T, int F, float* stats,
                                 mmb_stream_t stream) {
  MMB_REQUIRE(N >= 0 && T > 0 && F > 0 && N < ((int64_t)1 << 31), "bad size");
  if (N == 0) return MMB_OK;
  MMB_REQUIRE(val && mask && stats, "null pointer");
  gauss_moments_kernel<<<dim3((unsigned)N, (F + 127) / 128), 128, 0, as_stream(stream)>>>(val, mask, T, F, stats);
  MMB_LAUNCH_CHECK("gauss_moments");
  return MMB_OK;
}

extern "C" int mmb_gauss_ll_stats(int B, int n_mod, const int* n_seg, const float* const* seg_stats,
                                  const int* seg_F, const float* const* mu, const float* const* sigma, float* lp,
                                  float* const* dmu, float* const* dsigma, int* status, mmb_stream_t stream) {
  MMB_REQUIRE(n_seg && seg_stats && seg_F && mu && sigma && lp && status, "null pointer");
  MMB_REQUIRE(n_mod > 0 && n_mod <= kMaxMods, "1..8 modalities");
  MMB_REQUIRE(B > 0, "bad size");
  GaussStatArgs a = {};
  int s = 0;
  for (int m = 0; m < n_mod; ++m) {
    MMB_REQUIRE(n_seg[m] > 0 && n_seg[m] <= kMaxSegs, "1..4 segments per modality");
    a.n_seg[m] = n_seg[m];
    int D = 0;
    for (int g = 0; g < n_seg[m]; ++g, ++s) {
      MMB_REQUIRE(seg_stats[s] && seg_F[s] > 0, "null segment");
      a.st[m][g] = seg_stats[s];
      a.F[m][g] = seg_F[s];
      D += seg_F[s];
    }
    MMB_REQUIRE(mu[m] && sigma[m], "null mu/sigma");
    a.mu[m] = mu[m];
    a.sigma[m] = sigma[m];
    a.dmu[m] = dmu ? dmu[m] : nullptr;
    a.dsigma[m] = dsigma ? dsigma[m] : nullptr;
    a.D[m] = D;
  }
  a.n_mod = n_mod;
  gauss_ll_stats_kernel<<<dim3(B, n_mod), 128, 0, as_stream(stream)>>>(a, B, lp, status);
  MMB_LAUNCH_CHECK("gauss_ll_stats");
  return MMB_OK;
}

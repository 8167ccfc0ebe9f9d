// Device-side preprocessing of the audio / visual feature tensors:
// what the reference does in NumPy on the host before any torch code runs
//   normalize_data             utils.py:155-191   constant-feature drop, masks from exact zeros,
//                                                 (x + min) * 2 / (max - min) - 1 with the split's
//                                                 own min / max (the reference ADDS the minimum),
//                                                 padding -> -10
//   add_positional_embeddings  utils.py:130-153   pos_embed_dim extra columns; quirk kept: the
//                                                 sin / cos transform indexes the FIRST axis, so
//                                                 only data points 0 .. pos_embed_dim-1 get
//                                                 sinusoids, every other point the raw position
//   mask extension             simplesif.py:369-375  mask columns of ones for the new features
// in two launches (column min / max, then one fused pass that writes values and float masks in
// the layout MMData keeps on the device), instead of five host passes plus an H2D copy of the
// expanded arrays.
#include <math.h>

#include "common.cuh"

namespace mmb {

// Partial column min / max: CTA b scans rows [b * rows_per, ...), thread f owns column f.
__global__ void __launch_bounds__(256)
    feature_minmax_partial_kernel(const float* __restrict__ x, int64_t rows, int F, int64_t rows_per,
                                  float* __restrict__ pmin, float* __restrict__ pmax) {
  const int64_t r0 = (int64_t)blockIdx.x * rows_per;
  int64_t r1 = r0 + rows_per;
  if (r1 > rows) r1 = rows;
  for (int f = threadIdx.x; f < F; f += blockDim.x) {
    float mn = INFINITY, mx = -INFINITY;
    for (int64_t r = r0; r < r1; ++r) {
      const float v = __ldg(x + r * F + f);
      mn = fminf(mn, v);
      mx = fmaxf(mx, v);
    }
    pmin[(size_t)blockIdx.x * F + f] = mn;
    pmax[(size_t)blockIdx.x * F + f] = mx;
  }
}

__global__ void feature_minmax_final_kernel(const float* __restrict__ pmin, const float* __restrict__ pmax,
                                            int parts, int F, float* __restrict__ mn, float* __restrict__ mx) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  float a = INFINITY, b = -INFINITY;
  for (int p = 0; p < parts; ++p) {
    a = fminf(a, pmin[(size_t)p * F + f]);
    b = fmaxf(b, pmax[(size_t)p * F + f]);
  }
  mn[f] = a;
  mx[f] = b;
}

// out / mask: (N, T, F_out + P).  keep[f] = source column of output column f (the non-constant
// features, ascending); mn / mx are indexed by SOURCE column.
__global__ void __launch_bounds__(256)
    prep_features_kernel(const float* __restrict__ x, int64_t N, int T, int F_in, const int* __restrict__ keep,
                         int F_out, int P, const float* __restrict__ mn, const float* __restrict__ mx,
                         float* __restrict__ out, float* __restrict__ mask) {
  const int W = F_out + P;
  const int64_t total = N * T * W;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c = (int)(i % W);
    const int64_t row = i / W;
    float v, m;
    if (c < F_out) {
      const int src = keep[c];
      const float raw = __ldg(x + row * F_in + src);
      const bool pad = raw == 0.f;
      const float lo = mn[src], hi = mx[src];
      v = pad ? -10.f : (raw + lo) * 2.f / (hi - lo) - 1.f;
      m = pad ? 0.f : 1.f;
    } else {
      const int64_t n = row / T;
      const float t = (float)(row % T);
      const int half = P / 2;
      if (n < 2 * half) {
        // data point n = 2 i (+1): sin (cos) of t / 10000^(2 i / P), in every new column
        const int ii = (int)(n >> 1);
        const float scale = powf(10000.f, 2.f * ii / (float)P);
        v = (n & 1) ? cosf(t / scale) : sinf(t / scale);
      } else {
        v = t;
      }
      m = 1.f;
    }
    out[i] = v;
    mask[i] = m;
  }
}

}  // namespace mmb

using namespace mmb;

static int minmax_parts(int64_t rows) {
  int64_t p = (rows + 255) / 256;
  const int cap = sm_count() * 4;
  return (int)(p < 1 ? 1 : (p > cap ? cap : p));
}

extern "C" size_t mmb_feature_minmax_workspace_bytes(int64_t rows, int F) {
  return (size_t)2 * minmax_parts(rows) * (F > 0 ? F : 1) * sizeof(float);
}

extern "C" int mmb_feature_minmax(const float* x, int64_t rows, int F, float* mn, float* mx, void* ws,
                                  size_t ws_bytes, mmb_stream_t stream) {
  MMB_REQUIRE(x && mn && mx && ws, "null pointer");
  MMB_REQUIRE(rows > 0 && F > 0, "bad size");
  MMB_REQUIRE(ws_bytes >= mmb_feature_minmax_workspace_bytes(rows, F), "workspace too small");
  const int parts = minmax_parts(rows);
  const int64_t rows_per = ceil_div(rows, parts);
  float* pmin = (float*)ws;
  float* pmax = pmin + (size_t)parts * F;
  feature_minmax_partial_kernel<<<parts, 256, 0, as_stream(stream)>>>(x, rows, F, rows_per, pmin, pmax);
  MMB_LAUNCH_CHECK("feature_minmax_partial");
  feature_minmax_final_kernel<<<(F + 127) / 128, 128, 0, as_stream(stream)>>>(pmin, pmax, parts, F, mn, mx);
  MMB_LAUNCH_CHECK("feature_minmax_final");
  return MMB_OK;
}

extern "C" int mmb_prep_features(const float* x, int64_t N, int T, int F_in, const int* keep, int F_out,
                                 int pos_embed_dim, const float* mn, const float* mx, float* out, float* mask,
                                 mmb_stream_t stream) {
  MMB_REQUIRE(x && keep && mn && mx && out && mask, "null pointer");
  MMB_REQUIRE(N > 0 && T > 0 && F_in > 0 && F_out > 0 && F_out <= F_in && pos_embed_dim >= 0, "bad size");
  const int64_t total = N * T * (F_out + pos_embed_dim);
  int64_t blocks = ceil_div(total, 256 * 4);
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  prep_features_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(x, N, T, F_in, keep, F_out, pos_embed_dim, mn,
                                                                  mx, out, mask);
  MMB_LAUNCH_CHECK("prep_features");
  return MMB_OK;
}

// ---- masks (reference simplesif.py:36-47), SURVEY.md 8f N3 ------------------------------------------------
namespace mmb {
// update_masks: text mask = (ids != 0); the reference broadcasts it to (N, L, embedding_dim) ints on the host
// (1 GB at POM) -- here it is the (N, L) float vector, expanded as a stride-0 view by the caller.
__global__ void token_mask_kernel(const int64_t* __restrict__ ids, int64_t n, float* __restrict__ mask) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    mask[i] = __ldcs(ids + i) != 0 ? 1.f : 0.f;
}
// update_masks_vect: a time step is valid iff NO feature of it is exactly 0.  Warp per (n, t) row.
__global__ void __launch_bounds__(256)
    step_mask_kernel(const float* __restrict__ x, int64_t rows, int F, float* __restrict__ mask) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r = warp0; r < rows; r += nwarps) {
    bool ok = true;
    for (int f = lane; f < F; f += 32) ok = ok && (__ldcs(x + r * F + f) != 0.f);
    ok = __all_sync(0xffffffffu, ok);
    if (lane == 0) mask[r] = ok ? 1.f : 0.f;
  }
}
}  // namespace mmb

extern "C" int mmb_token_mask(const int64_t* ids, int64_t n, float* mask, mmb_stream_t stream) {
  MMB_REQUIRE(n >= 0, "negative size");
  if (n == 0) return MMB_OK;
  MMB_REQUIRE(ids && mask, "null pointer");
  const int64_t blocks = mmb::ceil_div(n, 1024);
  const int grid = (int)(blocks < (int64_t)mmb::sm_count() * 8 ? blocks : (int64_t)mmb::sm_count() * 8);
  mmb::token_mask_kernel<<<grid, 256, 0, mmb::as_stream(stream)>>>(ids, n, mask);
  MMB_LAUNCH_CHECK("token_mask");
  return MMB_OK;
}

extern "C" int mmb_step_mask(const float* x, int64_t rows, int F, float* mask, mmb_stream_t stream) {
  MMB_REQUIRE(rows >= 0 && F > 0, "bad size");
  if (rows == 0) return MMB_OK;
  MMB_REQUIRE(x && mask, "null pointer");
  const int64_t blocks = mmb::ceil_div(rows, 8);
  const int grid = (int)(blocks < (int64_t)mmb::sm_count() * 8 ? blocks : (int64_t)mmb::sm_count() * 8);
  mmb::step_mask_kernel<<<grid, 256, 0, mmb::as_stream(stream)>>>(x, rows, F, mask);
  MMB_LAUNCH_CHECK("step_mask");
  return MMB_OK;
}

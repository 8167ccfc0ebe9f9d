// Closed-form multimodal latent estimate (reference: sif2.py:103-114 calc_weights and 164-208
// estimate_embedding_overall_gpu2; call site simplesif.py:808-880).  SURVEY.md section 8f, N1.
//
// For every modality k (data x of shape (N, T, D_k), head biases b_mu, b_ls and weights W_mu,
// W_ls of shape (D_k, d)) the reference forms
//     q_mean  = (x - b_mu) / exp(2 b_ls)            q_sigma = (x - b_mu)^2 / exp(2 b_ls) - 1
// and returns the L2-normalised rows of
//     cs_n = [ sum_t w[n,t] E[n,t,:] + sum_k ( (sum_t q_mean_k[n,t,:]) W_mu_k
//                                            + (sum_t q_sigma_k[n,t,:]) W_ls_k ) ] / tw_n,
//     tw_n = sum_t w[n,t] + sum_k sum_{t,f} (q_mean_k + q_sigma_k)              (masks unused).
// The data enter only through their sums over time, so the path is:
//   closed_form_stats_kernel : one pass over the BASE tensors (the torch.cat'ed modalities of
//                              the call site are never materialised): S1_k, S2_k (N, D_k) and the
//                              per-(k, n) contribution to tw;
//   mmb_heads_backward (dz)  : the (N x sum D) x (sum D x d) product, split-K, deterministic;
//   closed_form_finish_kernel: weighted text sum + product, divide by tw, L2-normalise.
#include "common.cuh"

namespace mmb {

constexpr int kCfMaxMods = 8;
constexpr int kCfMaxSegs = 4;

struct CfArgs {
  const float* val[kCfMaxMods][kCfMaxSegs];
  int F[kCfMaxMods][kCfMaxSegs];
  int n_seg[kCfMaxMods];
  const float* b_mu[kCfMaxMods];
  const float* b_ls[kCfMaxMods];
  float* S1[kCfMaxMods];
  float* S2[kCfMaxMods];
  int D[kCfMaxMods];
  int n_mod;
};

// One CTA per (utterance n, modality k); thread f walks the T time steps of feature f
// (coalesced across f).  tw_part[k][n] = sum_f (S1 + S2), block-reduced in a fixed order.
__global__ void __launch_bounds__(128)
    closed_form_stats_kernel(const __grid_constant__ CfArgs args, int N, int T, float* __restrict__ tw_part) {
  __shared__ float red[4];
  const int n = blockIdx.x, k = blockIdx.y;
  const int D = args.D[k];
  float tot = 0.f;
  for (int f = threadIdx.x; f < D; f += blockDim.x) {
    int seg = 0, fl = f;
    while (seg + 1 < args.n_seg[k] && fl >= args.F[k][seg]) { fl -= args.F[k][seg]; ++seg; }
    const int F = args.F[k][seg];
    const float* x = args.val[k][seg] + (size_t)n * T * F + fl;
    const float bm = __ldg(args.b_mu[k] + f);
    const float inv = expf(-2.f * __ldg(args.b_ls[k] + f));   // 1 / exp(2 b_ls)
    float s1 = 0.f, s2 = 0.f;
    for (int t = 0; t < T; ++t) {
      const float df = __ldg(x + (size_t)t * F) - bm;
      const float a = df * inv;
      s1 += a;
      s2 += fmaf(df, a, -1.f);
    }
    args.S1[k][(size_t)n * D + f] = s1;
    args.S2[k][(size_t)n * D + f] = s2;
    tot += s1 + s2;
  }
  tot = warp_sum(tot);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = tot;
  __syncthreads();
  if (threadIdx.x == 0) tw_part[(size_t)k * N + n] = (red[0] + red[1]) + (red[2] + red[3]);
}

// One CTA per utterance: cs = (sum_t w_t E_t + prod) / tw, then cs / ||cs||.
__global__ void __launch_bounds__(128)
    closed_form_finish_kernel(int N, int L, int d, const float* __restrict__ sent_w,
                              const float* __restrict__ emb, const float* __restrict__ prod,
                              const float* __restrict__ tw_part, int n_mod, float* __restrict__ out) {
  __shared__ float red[4];
  __shared__ float bc;
  const int n = blockIdx.x;
  if (threadIdx.x == 0) {
    float tw = 0.f;
    for (int t = 0; t < L; ++t) tw += sent_w[(size_t)n * L + t];
    for (int k = 0; k < n_mod; ++k) tw += tw_part[(size_t)k * N + n];
    bc = tw;
  }
  __syncthreads();
  const float tw = bc;
  float nn = 0.f;
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    float s = 0.f;
    for (int t = 0; t < L; ++t) s = fmaf(sent_w[(size_t)n * L + t], __ldg(emb + ((size_t)n * L + t) * d + c), s);
    const float v = (s + prod[(size_t)n * d + c]) / tw;
    out[(size_t)n * d + c] = v;
    nn = fmaf(v, v, nn);
  }
  nn = warp_sum(nn);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = nn;
  __syncthreads();
  const float norm = sqrtf((red[0] + red[1]) + (red[2] + red[3]));
  for (int c = threadIdx.x; c < d; c += blockDim.x) out[(size_t)n * d + c] /= norm;
}

}  // namespace mmb

using namespace mmb;

extern "C" int mmb_closed_form_stats(int N, int T, int n_mod, const int* n_seg, const float* const* seg_val,
                                     const int* seg_F, const float* const* b_mu, const float* const* b_ls,
                                     float* const* S1, float* const* S2, float* tw_part, mmb_stream_t stream) {
  MMB_REQUIRE(n_seg && seg_val && seg_F && b_mu && b_ls && S1 && S2 && tw_part, "null pointer");
  MMB_REQUIRE(n_mod > 0 && n_mod <= kCfMaxMods, "1..8 modalities");
  MMB_REQUIRE(N > 0 && T > 0, "bad size");
  CfArgs a = {};
  int s = 0;
  for (int m = 0; m < n_mod; ++m) {
    MMB_REQUIRE(n_seg[m] > 0 && n_seg[m] <= kCfMaxSegs, "1..4 segments per modality");
    a.n_seg[m] = n_seg[m];
    int D = 0;
    for (int g = 0; g < n_seg[m]; ++g, ++s) {
      MMB_REQUIRE(seg_val[s] && seg_F[s] > 0, "null segment");
      a.val[m][g] = seg_val[s];
      a.F[m][g] = seg_F[s];
      D += seg_F[s];
    }
    MMB_REQUIRE(b_mu[m] && b_ls[m] && S1[m] && S2[m], "null modality buffer");
    a.b_mu[m] = b_mu[m]; a.b_ls[m] = b_ls[m]; a.S1[m] = S1[m]; a.S2[m] = S2[m]; a.D[m] = D;
  }
  a.n_mod = n_mod;
  closed_form_stats_kernel<<<dim3(N, n_mod), 128, 0, as_stream(stream)>>>(a, N, T, tw_part);
  MMB_LAUNCH_CHECK("closed_form_stats");
  return MMB_OK;
}

extern "C" int mmb_closed_form_finish(int N, int L, int d, const float* sent_w, const float* emb,
                                      const float* prod, const float* tw_part, int n_mod, float* out,
                                      mmb_stream_t stream) {
  MMB_REQUIRE(sent_w && emb && prod && tw_part && out, "null pointer");
  MMB_REQUIRE(N > 0 && L > 0 && d > 0 && n_mod > 0, "bad size");
  closed_form_finish_kernel<<<N, 128, 0, as_stream(stream)>>>(N, L, d, sent_w, emb, prod, tw_part, n_mod, out);
  MMB_LAUNCH_CHECK("closed_form_finish");
  return MMB_OK;
}

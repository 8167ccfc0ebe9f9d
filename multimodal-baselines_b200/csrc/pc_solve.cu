// Principal components from the Gram (reference: sif_functions.py:58-67 compute_pc, i.e.
// scikit-learn TruncatedSVD(n_components=npc, n_iter=7, random_state=0) -> randomized SVD).
//
// sklearn's output is a function of G = X^T X and the seeded start block only (SURVEY.md
// section 7 H1; oracle/sif_oracle.py:compute_pc_from_gram):
//   Q = orth(S0); repeat n_iter: Q = orth(G Q);  Y = G Q;  T = Q^T Y
//   N >= d : components = top eigenvectors of Y T^-1 Y^T = left singular vectors of
//            W = Y R^-1 with T = R^T R            (sklearn: right singular vectors of Q_x^T X)
//   N <  d : components = Q * top eigenvectors of T   (sklearn works on X^T: Rayleigh-Ritz)
//   rows unit-normalised, largest-|entry| of each row made positive (svd_flip, v-based).
// The problem is d x k with d = 300 and k = npc + 10, so it runs as ONE CTA in FP64 (thread
// i owns row i), entirely on the device: no host round trip between the Gram and the
// projection pass.  orth() is CholeskyQR2 with rank-deficient columns dropped;
// the k x k symmetric eigenproblems use cyclic Jacobi on one warp.
#include "common.cuh"

namespace mmb {

constexpr int kPcMaxK = 32;

struct PcSmem {
  double S[kPcMaxK][kPcMaxK];   // small symmetric matrix / Cholesky factor R (upper)
  double Vv[kPcMaxK][kPcMaxK];  // Jacobi eigenvectors
  double lam[kPcMaxK];
  double red[32];
  int drop[kPcMaxK];
  int order[kPcMaxK];
  int flag;
};

// S[a][b] = sum_i A[i][a] * B[i][b] for a,b < k (A, B are d x k in global memory). Warps take
// (a,b) pairs round-robin; lanes stride over rows; fixed order -> deterministic.
__device__ void small_gram(PcSmem& sm, const double* A, const double* B, int d, int k, bool symmetric) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int p = warp; p < k * k; p += nwarps) {
    const int a = p / k, b = p % k;
    if (symmetric && b < a) continue;
    double s = 0.0;
    for (int i = lane; i < d; i += 32) s = fma(A[(size_t)i * k + a], B[(size_t)i * k + b], s);
    s = warp_sum(s);
    if (lane == 0) {
      sm.S[a][b] = s;
      if (symmetric) sm.S[b][a] = s;
    }
  }
  __syncthreads();
}

// In-place upper Cholesky S = R^T R (right-looking, warp 0, lane j owns column j); columns
// whose pivot is not positive relative to the largest diagonal entry are dropped (they lie in
// the span of earlier columns).  sm.lam[c] receives 1 / R[c][c] for apply_rinv.
__device__ void cholesky_drop(PcSmem& sm, int k) {
  if (threadIdx.x < 32) {
    const int j = threadIdx.x;
    double dmax = 0.0;
    for (int c = 0; c < k; ++c) dmax = fmax(dmax, sm.S[c][c]);
    for (int c = 0; c < k; ++c) {
      const double piv = sm.S[c][c];
      const bool bad = !(piv > 1e-26 * dmax) || !isfinite(piv);
      const double rinv = bad ? 0.0 : rsqrt(piv);
      __syncwarp();
      if (j == c) {
        sm.drop[c] = bad;
        sm.S[c][c] = bad ? 1.0 : piv * rinv;   // sqrt(piv)
        sm.lam[c] = bad ? 0.0 : rinv;
      }
      double rcj = 0.0;
      if (j > c && j < k) {
        rcj = sm.S[c][j] * rinv;               // row c of R (0 for a dropped column)
        sm.S[c][j] = rcj;
      }
      __syncwarp();
      if (j > c && j < k) {                    // trailing update of column j, rows c+1..j
        for (int i = c + 1; i <= j; ++i) sm.S[i][j] -= sm.S[c][i] * rcj;
      }
      __syncwarp();
    }
  }
  __syncthreads();
}

// row <- row * R^-1 (forward substitution on this thread's row); dropped columns become 0.
template <int KMAX>
__device__ void apply_rinv(const PcSmem& sm, double (&q)[KMAX], int k) {
#pragma unroll
  for (int c = 0; c < KMAX; ++c) {
    if (c < k) {
      double acc = q[c];
#pragma unroll
      for (int a = 0; a < KMAX; ++a)
        if (a < c) acc -= q[a] * sm.S[a][c];
      q[c] = acc * sm.lam[c];   // lam[c] = 1 / R[c][c], 0 for a dropped column
    }
  }
}

template <int KMAX>
__device__ void store_row(double* M, const double (&q)[KMAX], int row, int d, int k) {
  if (row < d) {
#pragma unroll
    for (int c = 0; c < KMAX; ++c)
      if (c < k) M[(size_t)row * k + c] = q[c];
  }
  __syncthreads();
}

// CholeskyQR2 on the rows held in registers; Qg (global, d x k) is scratch and ends up
// holding the orthonormalised block.
template <int KMAX>
__device__ void orth(PcSmem& sm, double (&q)[KMAX], double* Qg, int row, int d, int k, int passes) {
  for (int pass = 0; pass < passes; ++pass) {
    store_row<KMAX>(Qg, q, row, d, k);
    small_gram(sm, Qg, Qg, d, k, true);
    cholesky_drop(sm, k);
    apply_rinv<KMAX>(sm, q, k);
    __syncthreads();
  }
  store_row<KMAX>(Qg, q, row, d, k);
}

// y = (G Q)[row, :], G symmetric float32 (read column `row`, i.e. coalesced across threads).
template <int KMAX>
__device__ void gram_times(const float* __restrict__ G, const double* Qg, double (&y)[KMAX], int row,
                           int d, int k) {
#pragma unroll
  for (int c = 0; c < KMAX; ++c) y[c] = 0.0;
  if (row < d) {
    for (int j = 0; j < d; ++j) {
      const double g = (double)__ldg(G + (size_t)j * d + row);
#pragma unroll
      for (int c = 0; c < KMAX; ++c)
        if (c < k) y[c] = fma(g, Qg[(size_t)j * k + c], y[c]);
    }
  }
}

// Cyclic Jacobi on sm.S (k x k symmetric, k <= 32) by warp 0; eigenvalues -> sm.lam,
// eigenvectors (columns) -> sm.Vv; sm.order = indices by descending eigenvalue.
__device__ void jacobi_eig(PcSmem& sm, int k) {
  if (threadIdx.x < 32) {
    const int l = threadIdx.x;
    for (int c = 0; c < k; ++c)
      if (l < k) sm.Vv[l][c] = (l == c) ? 1.0 : 0.0;
    __syncwarp();
    for (int sweep = 0; sweep < 16; ++sweep) {
      double off = 0.0, dg = 0.0;
      if (l < k) {
        for (int c = 0; c < k; ++c) {
          const double v = sm.S[l][c];
          if (c > l) off += v * v;
          if (c == l) dg += v * v;
        }
      }
      off = warp_sum(off);
      dg = warp_sum(dg);
      if (!(off > 1e-27 * dg)) break;   // off-diagonal mass below ~3e-14 of the diagonal: converged
      for (int p = 0; p < k - 1; ++p) {
        for (int q = p + 1; q < k; ++q) {
          const double apq = sm.S[p][q];
          if (fabs(apq) > 1e-17 * sqrt(fabs(sm.S[p][p] * sm.S[q][q])) && fabs(apq) > 1e-300) {
            const double app = sm.S[p][p], aqq = sm.S[q][q];
            const double theta = (aqq - app) / (2.0 * apq);
            const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
            __syncwarp();
            if (l < k) {
              const double vlp = sm.Vv[l][p], vlq = sm.Vv[l][q];
              sm.Vv[l][p] = c * vlp - s * vlq;
              sm.Vv[l][q] = s * vlp + c * vlq;
              if (l != p && l != q) {
                const double alp = sm.S[l][p], alq = sm.S[l][q];
                const double nlp = c * alp - s * alq, nlq = s * alp + c * alq;
                sm.S[l][p] = nlp; sm.S[p][l] = nlp;
                sm.S[l][q] = nlq; sm.S[q][l] = nlq;
              }
            }
            if (l == 0) {
              sm.S[p][p] = app - t * apq;
              sm.S[q][q] = aqq + t * apq;
              sm.S[p][q] = 0.0;
              sm.S[q][p] = 0.0;
            }
            __syncwarp();
          }
        }
      }
    }
    if (l < k) sm.lam[l] = sm.S[l][l];
    __syncwarp();
    if (l == 0) {  // selection sort, descending, stable in index
      bool used[kPcMaxK];
      for (int c = 0; c < k; ++c) used[c] = false;
      for (int r = 0; r < k; ++r) {
        int best = -1;
        for (int c = 0; c < k; ++c)
          if (!used[c] && (best < 0 || sm.lam[c] > sm.lam[best])) best = c;
        used[best] = true;
        sm.order[r] = best;
      }
    }
  }
  __syncthreads();
}

template <int KMAX>
__global__ void __launch_bounds__(512)
    pc_from_gram_kernel(const float* __restrict__ G, int d, const double* __restrict__ S0, int k,
                        int npc, int transposed, int n_iter, float* __restrict__ pc, double* ws) {
  __shared__ PcSmem sm;
  double* Qg = ws;                     // d x k
  double* Yg = ws + (size_t)d * k;     // d x k
  double* cbuf = Yg + (size_t)d * k;   // d
  const int row = threadIdx.x;
  double q[KMAX], y[KMAX];
#pragma unroll
  for (int c = 0; c < KMAX; ++c) q[c] = (row < d && c < k) ? S0[(size_t)row * k + c] : 0.0;
  if (threadIdx.x < kPcMaxK) sm.drop[threadIdx.x] = 0;
  __syncthreads();

  // Only the span matters between multiplications, so one CholeskyQR pass keeps the block
  // well enough conditioned there; the last one is repeated (CholeskyQR2) for orthonormality.
  orth<KMAX>(sm, q, Qg, row, d, k, n_iter == 0 ? 2 : 1);
  for (int it = 0; it < n_iter; ++it) {
    gram_times<KMAX>(G, Qg, y, row, d, k);
    __syncthreads();  // everyone has finished reading Qg
#pragma unroll
    for (int c = 0; c < KMAX; ++c) q[c] = y[c];
    orth<KMAX>(sm, q, Qg, row, d, k, it + 1 == n_iter ? 2 : 1);
  }
  gram_times<KMAX>(G, Qg, y, row, d, k);
  store_row<KMAX>(Yg, y, row, d, k);
  small_gram(sm, Qg, Yg, d, k, false);   // T = Q^T Y
  if (threadIdx.x < k) {                 // symmetrise
    const int a = threadIdx.x;
    for (int b = a + 1; b < k; ++b) {
      const double m = 0.5 * (sm.S[a][b] + sm.S[b][a]);
      sm.S[a][b] = m;
      sm.S[b][a] = m;
    }
  }
  __syncthreads();

  // y becomes this thread's row of the basis the components are expressed in
  if (transposed) {
#pragma unroll
    for (int c = 0; c < KMAX; ++c) y[c] = q[c];
    jacobi_eig(sm, k);
  } else {
    cholesky_drop(sm, k);
    apply_rinv<KMAX>(sm, y, k);
    __syncthreads();
    store_row<KMAX>(Yg, y, row, d, k);
    small_gram(sm, Yg, Yg, d, k, true);  // W^T W
    jacobi_eig(sm, k);
  }

  for (int c = 0; c < npc; ++c) {
    const int sel = sm.order[c];
    double v = 0.0;
#pragma unroll
    for (int a = 0; a < KMAX; ++a)
      if (a < k) v = fma(y[a], sm.Vv[a][sel], v);
    if (row < d) cbuf[row] = v;
    __syncthreads();
    if (threadIdx.x == 0) {  // norm + sklearn's sign rule (first arg-max of |v|)
      double n2 = 0.0, best = -1.0;
      int arg = 0;
      for (int i = 0; i < d; ++i) {
        const double x = cbuf[i];
        n2 = fma(x, x, n2);
        if (fabs(x) > best) { best = fabs(x); arg = i; }
      }
      const double sgn = cbuf[arg] < 0.0 ? -1.0 : 1.0;
      sm.red[0] = sgn / sqrt(n2);
    }
    __syncthreads();
    if (row < d) pc[(size_t)c * d + row] = (float)(v * sm.red[0]);
    __syncthreads();
  }
}

// S0 = X^T Omega for the N < d case (N tiny): thread per (column of X, column of Omega).
__global__ void start_block_xt_kernel(const float* __restrict__ X, int64_t N, int d,
                                      const double* __restrict__ Omega, int k, double* __restrict__ S0) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= d * k) return;
  const int c = idx % k, j = idx / k;
  double s = 0.0;
  for (int64_t n = 0; n < N; ++n) s = fma((double)X[n * d + j], Omega[n * k + c], s);
  S0[(size_t)j * k + c] = s;
}

}  // namespace mmb

using namespace mmb;

extern "C" size_t mmb_pc_workspace_bytes(int d, int k) {
  return ((size_t)2 * d * k + d) * sizeof(double);
}

extern "C" int mmb_pc_from_gram(const float* G, int d, const double* S0, int k, int npc,
                                int transposed, int n_iter, float* pc, void* ws, size_t ws_bytes,
                                mmb_stream_t stream) {
  MMB_REQUIRE(G && S0 && pc && ws, "null pointer");
  MMB_REQUIRE(d > 0 && d <= 512, "d must be in [1, 512]");
  MMB_REQUIRE(k > 0 && k <= kPcMaxK && npc > 0 && npc <= k, "need 0 < npc <= k <= 32");
  MMB_REQUIRE(ws_bytes >= mmb_pc_workspace_bytes(d, k), "workspace too small");
  MMB_REQUIRE(n_iter >= 0, "n_iter < 0");
  const int threads = ((d + 31) / 32) * 32;
  if (k <= 16)
    pc_from_gram_kernel<16><<<1, threads, 0, as_stream(stream)>>>(G, d, S0, k, npc, transposed, n_iter,
                                                                 pc, (double*)ws);
  else
    pc_from_gram_kernel<32><<<1, threads, 0, as_stream(stream)>>>(G, d, S0, k, npc, transposed, n_iter,
                                                                 pc, (double*)ws);
  MMB_LAUNCH_CHECK("pc_from_gram");
  return MMB_OK;
}

extern "C" int mmb_start_block_xt(const float* X, int64_t N, int d, const double* Omega, int k,
                                  double* S0, mmb_stream_t stream) {
  MMB_REQUIRE(X && Omega && S0, "null pointer");
  MMB_REQUIRE(N > 0 && d > 0 && k > 0, "bad size");
  const int n = d * k;
  start_block_xt_kernel<<<(n + 127) / 128, 128, 0, as_stream(stream)>>>(X, N, d, Omega, k, S0);
  MMB_LAUNCH_CHECK("start_block_xt");
  return MMB_OK;
}

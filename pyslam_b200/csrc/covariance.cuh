// covariance.cuh -- dense covariance of the estimate, H^-1, replacing
// `splinalg.inv(precision.tocsc()).toarray()` (pyslam/problem.py:196-203).
// Not on the per-iteration hot path.  Uses the factors the solver
// already has: with S = L L^T the reduced system, G = L^-1,
//     Sigma_cc = G^T G,   Sigma_pc = -B_p Sigma_cc,   Sigma_pp' = d_pp' V_p^-1 - Sigma_pc B_p'^T,
// where B_p = V_p^-1 W_p^T scattered to the columns of p's poses (3 x n_pad).
#pragma once
#include "cholesky.cuh"

namespace bs {

// shared tile <- global tile, transposed: s[c][r] = g[r][c]
BS_D void tile_load_t(double* __restrict__ s, const double* __restrict__ gsrc, int ld) {
  for (int e = threadIdx.x; e < kNB * kNB; e += kCholThreads) {
    const int r = e / kNB, c = e % kNB;
    s[c * kLd + r] = __ldcg(gsrc + (size_t)r * ld + c);
  }
}

// G = L^-1, tile column c per CTA:  G_kc = X_kk (d_kc I - sum_{m=c}^{k-1} L_km G_mc)
__global__ void __launch_bounds__(kCholThreads, 1)
cov_linv_kernel(const double* __restrict__ S, int ld, const double* __restrict__ Linv, int nt,
                const unsigned char* __restrict__ mask, double* __restrict__ G) {
  extern __shared__ double smem[];
  double* sA = smem;
  double* sB = smem + kNB * kLd;
  const int c = blockIdx.x;
  for (int k = c; k < nt; ++k) {
    TileAcc acc;
    acc_zero(acc);
    for (int m = c; m < k; ++m) {
      if (!mask[(size_t)k * nt + m]) continue;
      __syncthreads();
      tile_load(sA, S + (size_t)k * kNB * ld + (size_t)m * kNB, ld, kNB);
      tile_load_t(sB, G + (size_t)m * kNB * ld + (size_t)c * kNB, ld);
      __syncthreads();
      tile_mma_abt(sA, sB, acc);
    }
    __syncthreads();
    // sB[n][r] = (d_kc I - acc)[r][n]
    acc_foreach([&](int i, int j, int r, int n) {
      sB[n * kLd + r] = ((k == c && r == n) ? 1.0 : 0.0) - acc.c[i][j][0];
      sB[(n + 1) * kLd + r] = ((k == c && r == n + 1) ? 1.0 : 0.0) - acc.c[i][j][1];
    });
    tile_load(sA, Linv + (size_t)k * kNB * kNB, kNB, kNB);
    __syncthreads();
    acc_zero(acc);
    tile_mma_abt(sA, sB, acc);
    double* Gk = G + (size_t)k * kNB * ld + (size_t)c * kNB;
    acc_foreach([&](int i, int j, int r, int n) {
      *reinterpret_cast<double2*>(Gk + (size_t)r * ld + n) = make_double2(acc.c[i][j][0], acc.c[i][j][1]);
    });
    __threadfence();
    __syncthreads();
  }
}

// X = G^T G, one CTA per tile (a, b), a >= b; written symmetrically into cov (leading dimension D)
__global__ void __launch_bounds__(kCholThreads, 1)
cov_gtg_kernel(const double* __restrict__ G, int ld, int nt, double* __restrict__ cov, size_t D) {
  const int a = blockIdx.y, b = blockIdx.x;
  if (b > a) return;
  extern __shared__ double smem[];
  double* sA = smem;
  double* sB = smem + kNB * kLd;
  TileAcc acc;
  acc_zero(acc);
  for (int k = a; k < nt; ++k) {
    __syncthreads();
    tile_load_t(sA, G + (size_t)k * kNB * ld + (size_t)a * kNB, ld);
    if (a != b) tile_load_t(sB, G + (size_t)k * kNB * ld + (size_t)b * kNB, ld);
    __syncthreads();
    tile_mma_abt(sA, a != b ? sB : sA, acc);
  }
  acc_foreach([&](int i, int j, int r, int n) {
    const size_t R = (size_t)a * kNB + r, Cc = (size_t)b * kNB + n;
    cov[R * D + Cc] = acc.c[i][j][0];
    cov[R * D + Cc + 1] = acc.c[i][j][1];
    cov[Cc * D + R] = acc.c[i][j][0];
    cov[(Cc + 1) * D + R] = acc.c[i][j][1];
  });
}

struct CovLmArgs {
  int n_lm, n_obs, n_pad;
  size_t D;
  const int* __restrict__ obs_pose;
  const int* __restrict__ lm_start;
  const int* __restrict__ lm_obs;    // [N] CSR position -> observation index
  const int* __restrict__ pose_off;
  const double* __restrict__ W;      // tiled, see w_index()
  const double* __restrict__ Vinv;   // [n_lm][6]
  double* __restrict__ cov;          // [D][D]
};

BS_D double sym6(const double* v, int a, int b) {
  const int idx[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
  return v[idx[a][b]];
}

// B_p[a][off_j + r] = sum_c Vinv_p[a][c] W_j[r][c]
BS_D double cov_B(const CovLmArgs& A, int p, int j, int a, int r) {
  const double* vi = A.Vinv + 6 * (size_t)p;
  return sym6(vi, a, 0) * A.W[w_index(j, 3 * r)] + sym6(vi, a, 1) * A.W[w_index(j, 3 * r + 1)] +
         sym6(vi, a, 2) * A.W[w_index(j, 3 * r + 2)];
}

// Sigma_pc = -B_p Sigma_cc : grid (n_pad columns / 128, 3 * n_lm rows)
__global__ void __launch_bounds__(128) cov_lm_pose_kernel(const CovLmArgs A) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  const int p = blockIdx.y / 3, a = blockIdx.y % 3;
  if (col >= A.n_pad) return;
  double s = 0.0;
  for (int jj = A.lm_start[p]; jj < A.lm_start[p + 1]; ++jj) {
    const int j = A.lm_obs[jj];
    const int off = A.pose_off[A.obs_pose[j]];
    if (off < 0) continue;
    for (int r = 0; r < 6; ++r) s += cov_B(A, p, j, a, r) * A.cov[(size_t)(off + r) * A.D + col];
  }
  const size_t row = (size_t)A.n_pad + 3 * (size_t)p + a;
  A.cov[row * A.D + col] = -s;
  A.cov[(size_t)col * A.D + row] = -s;
}

// Sigma_pp' = d_pp' V_p^-1 - Sigma_pc B_p'^T : grid (n_lm (p') / 128, 3 * n_lm rows (p, a)); 3 columns each
__global__ void __launch_bounds__(128) cov_lm_lm_kernel(const CovLmArgs A) {
  const int p2 = blockIdx.x * blockDim.x + threadIdx.x;
  const int p = blockIdx.y / 3, a = blockIdx.y % 3;
  if (p2 >= A.n_lm) return;
  const size_t row = (size_t)A.n_pad + 3 * (size_t)p + a;
  double out[3] = {0.0, 0.0, 0.0};
  for (int jj = A.lm_start[p2]; jj < A.lm_start[p2 + 1]; ++jj) {
    const int j = A.lm_obs[jj];
    const int off = A.pose_off[A.obs_pose[j]];
    if (off < 0) continue;
    for (int r = 0; r < 6; ++r) {
      const double spc = A.cov[row * A.D + off + r];        // Sigma_pc entry (written by cov_lm_pose_kernel)
      for (int b = 0; b < 3; ++b) out[b] -= spc * cov_B(A, p2, j, b, r);
    }
  }
  for (int b = 0; b < 3; ++b) {
    double v = out[b];
    if (p2 == p) v += sym6(A.Vinv + 6 * (size_t)p, a, b);
    A.cov[row * A.D + (size_t)A.n_pad + 3 * (size_t)p2 + b] = v;
  }
}

}  // namespace bs

// ransac.cuh -- frame-to-frame RANSAC on the device: SURVEY 8 f2.
// Replaces, for all hypotheses of FrameToFrameRANSAC.perform_ransac at once (pyslam/pipelines/ransac.py:107-165):
//   compute_transform_fast   (ransac.py:12-56)   rigid transform of a minimal set by the SVD method
//                                                C_21 = U diag(1, 1, det U det V) V^T,  W = 1/n sum p2c p1c^T
//   compute_ransac_cost      (ransac.py:155-165) inlier mask  |pi(T_21 p_1) - obs_2|^2 < thresh  per hypothesis
// One thread per hypothesis for the transform (3 x 3 problem), one CTA per hypothesis for the count over the
// N points, then the first arg-max (np.argmax) and the mask of the winner.
#pragma once
#include "common.cuh"

namespace bs {

// Eigen-decomposition of a symmetric 3x3 matrix (cyclic Jacobi): A = V diag(l) V^T, columns of V.
BS_D void jacobi_eig3(double A[3][3], double V[3][3], double l[3]) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) V[i][j] = i == j ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 12; ++sweep) {
    const double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
    if (off < 1e-300) break;
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
      for (int q = p + 1; q < 3; ++q) {
        if (fabs(A[p][q]) < 1e-300) continue;
        const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
#pragma unroll
        for (int k = 0; k < 3; ++k) {          // A <- A J
          const double akp = A[k][p], akq = A[k][q];
          A[k][p] = c * akp - s * akq;
          A[k][q] = s * akp + c * akq;
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {          // A <- J^T A
          const double apk = A[p][k], aqk = A[q][k];
          A[p][k] = c * apk - s * aqk;
          A[q][k] = s * apk + c * aqk;
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq;
          V[k][q] = s * vkp + c * vkq;
        }
      }
  }
  l[0] = A[0][0]; l[1] = A[1][1]; l[2] = A[2][2];
}

// T_21 (4x4 row-major) of every minimal set: idx [n_hyp][n_min] rows of pts_1 / pts_2.
// C = u1 v1^T + u2 v2^T + (u1 x u2)(v1 x v2)^T with (sigma_i, u_i, v_i) the two leading singular triplets of W:
// this equals U diag(1, 1, det U det V) V^T for every sign choice of the third singular vectors.
__global__ void __launch_bounds__(128) ransac_transform_kernel(int n_hyp, int n_min, const int* __restrict__ idx,
                                                               const double* __restrict__ pts1, const double* __restrict__ pts2,
                                                               double* __restrict__ T_out) {
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= n_hyp) return;
  double c1[3] = {0, 0, 0}, c2[3] = {0, 0, 0};
  for (int k = 0; k < n_min; ++k) {
    const int i = idx[(size_t)h * n_min + k];
#pragma unroll
    for (int d = 0; d < 3; ++d) { c1[d] += pts1[3 * (size_t)i + d]; c2[d] += pts2[3 * (size_t)i + d]; }
  }
#pragma unroll
  for (int d = 0; d < 3; ++d) { c1[d] /= n_min; c2[d] /= n_min; }
  double W[3][3] = {};
  for (int k = 0; k < n_min; ++k) {
    const int i = idx[(size_t)h * n_min + k];
    double a[3], b[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) { a[d] = pts1[3 * (size_t)i + d] - c1[d]; b[d] = pts2[3 * (size_t)i + d] - c2[d]; }
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) W[r][c] += b[r] * a[c] / n_min;
  }
  double A[3][3], V[3][3], l[3];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) A[r][c] = W[0][r] * W[0][c] + W[1][r] * W[1][c] + W[2][r] * W[2][c];   // W^T W
  jacobi_eig3(A, V, l);
  // the two largest eigenvalues
  int i0 = 0, i1 = 1, i2 = 2;
  if (l[i0] < l[i1]) { int t = i0; i0 = i1; i1 = t; }
  if (l[i1] < l[i2]) { int t = i1; i1 = i2; i2 = t; }
  if (l[i0] < l[i1]) { int t = i0; i0 = i1; i1 = t; }
  double v1[3], v2[3], u1[3], u2[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) { v1[d] = V[d][i0]; v2[d] = V[d][i1]; }
  auto mulW = [&](const double* v, double* u) {
#pragma unroll
    for (int r = 0; r < 3; ++r) u[r] = W[r][0] * v[0] + W[r][1] * v[1] + W[r][2] * v[2];
  };
  auto normalize = [](double* u) {
    const double n = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
    if (n > 0.0) { u[0] /= n; u[1] /= n; u[2] /= n; }
  };
  mulW(v1, u1); normalize(u1);
  mulW(v2, u2);
  const double dp = u1[0] * u2[0] + u1[1] * u2[1] + u1[2] * u2[2];      // re-orthogonalise (sigma_2 may be tiny)
#pragma unroll
  for (int d = 0; d < 3; ++d) u2[d] -= dp * u1[d];
  normalize(u2);
  const double u3[3] = {u1[1] * u2[2] - u1[2] * u2[1], u1[2] * u2[0] - u1[0] * u2[2], u1[0] * u2[1] - u1[1] * u2[0]};
  const double v3[3] = {v1[1] * v2[2] - v1[2] * v2[1], v1[2] * v2[0] - v1[0] * v2[2], v1[0] * v2[1] - v1[1] * v2[0]};
  double* T = T_out + 16 * (size_t)h;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    double tr = c2[r];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double C = u1[r] * v1[c] + u2[r] * v2[c] + u3[r] * v3[c];
      T[4 * r + c] = C;
      tr -= C * c1[c];
    }
    T[4 * r + 3] = tr;
  }
  T[12] = 0.0; T[13] = 0.0; T[14] = 0.0; T[15] = 1.0;
}

// inlier count of hypothesis blockIdx.x (T 4x4 row-major); intr = (cu, cv, fu, fv, b), b <= 0: RGB-D camera
__global__ void __launch_bounds__(256) ransac_count_kernel(int n_pts, const double* __restrict__ T_all, const double* __restrict__ pts1,
                                                           const double* __restrict__ obs2, double cu, double cv, double fu, double fv,
                                                           double b, double thresh, int* __restrict__ counts) {
  __shared__ int sred[8];
  const double* T = T_all + 16 * (size_t)blockIdx.x;
  double P[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) P[k] = T[k];
  int cnt = 0;
  for (int i = threadIdx.x; i < n_pts; i += blockDim.x) {
    const double X = pts1[3 * (size_t)i], Y = pts1[3 * (size_t)i + 1], Z = pts1[3 * (size_t)i + 2];
    const double x = P[0] * X + P[1] * Y + P[2] * Z + P[3];
    const double y = P[4] * X + P[5] * Y + P[6] * Z + P[7];
    const double z = P[8] * X + P[9] * Y + P[10] * Z + P[11];
    const double iz = 1.0 / z;
    const double e0 = fu * x * iz + cu - obs2[3 * (size_t)i];
    const double e1 = fv * y * iz + cv - obs2[3 * (size_t)i + 1];
    const double e2 = (b > 0.0 ? fu * b * iz : z) - obs2[3 * (size_t)i + 2];
    cnt += (e0 * e0 + e1 * e1 + e2 * e2 < thresh) ? 1 : 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sred[w];
    counts[blockIdx.x] = t;
  }
}

// first arg-max of counts (np.argmax) -> best[0], count -> best[1]; then the winner's inlier mask
__global__ void __launch_bounds__(256) ransac_best_kernel(int n_hyp, int n_pts, const int* __restrict__ counts, const double* __restrict__ T_all,
                                                          const double* __restrict__ pts1, const double* __restrict__ obs2, double cu,
                                                          double cv, double fu, double fv, double b, double thresh,
                                                          int* __restrict__ best, unsigned char* __restrict__ mask) {
  __shared__ int s_best;
  if (threadIdx.x == 0) {
    int bi = 0;
    for (int h = 1; h < n_hyp; ++h)
      if (counts[h] > counts[bi]) bi = h;
    s_best = bi;
    best[0] = bi;
    best[1] = counts[bi];
  }
  __syncthreads();
  const double* T = T_all + 16 * (size_t)s_best;
  for (int i = threadIdx.x; i < n_pts; i += blockDim.x) {
    const double X = pts1[3 * (size_t)i], Y = pts1[3 * (size_t)i + 1], Z = pts1[3 * (size_t)i + 2];
    const double x = T[0] * X + T[1] * Y + T[2] * Z + T[3];
    const double y = T[4] * X + T[5] * Y + T[6] * Z + T[7];
    const double z = T[8] * X + T[9] * Y + T[10] * Z + T[11];
    const double iz = 1.0 / z;
    const double e0 = fu * x * iz + cu - obs2[3 * (size_t)i];
    const double e1 = fv * y * iz + cv - obs2[3 * (size_t)i + 1];
    const double e2 = (b > 0.0 ? fu * b * iz : z) - obs2[3 * (size_t)i + 2];
    mask[i] = (e0 * e0 + e1 * e1 + e2 * e2 < thresh) ? 1 : 0;
  }
}

}  // namespace bs

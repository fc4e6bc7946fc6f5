// schur.cuh -- landmark elimination.  The reference solves the full sparse
// system with SuperLU (`splinalg.spsolve`, pyslam/problem.py:186); here the
// 3x3 landmark blocks are inverted in a batch, the reduced camera system
//     (U - sum_p W_p V_p^-1 W_p^T) dx_c = b_c - sum_p W_p V_p^-1 b_p
// is formed in the dense lower triangle of S, and after the dense solve the
// landmark updates follow from  dx_p = V_p^-1 (b_p - W_p^T dx_c).
#pragma once
#include "common.cuh"

namespace bs {

// Vg[q] = (xx,xy,xz,yy,yz,zz | b0,b1,b2)  ->  Vinv[q] = (xx,xy,xz,yy,yz,zz) of (V + lambda diag V)^-1
__global__ void __launch_bounds__(256) landmark_invert_kernel(int n_lm, const double* __restrict__ Vg,
                                                              double lambda, double* __restrict__ Vinv) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_lm) return;
  const double* v = Vg + 9 * (size_t)q;
  const double s = 1.0 + lambda;
  const double a = v[0] * s, b = v[1], c = v[2], d = v[3] * s, e = v[4], f = v[5] * s;
  // cofactors of the symmetric matrix [[a,b,c],[b,d,e],[c,e,f]]
  const double c00 = d * f - e * e;
  const double c01 = c * e - b * f;
  const double c02 = b * e - c * d;
  const double c11 = a * f - c * c;
  const double c12 = b * c - a * e;
  const double c22 = a * d - b * b;
  const double det = a * c00 + b * c01 + c * c02;
  const double id = 1.0 / det;
  double* o = Vinv + 6 * (size_t)q;
  o[0] = c00 * id; o[1] = c01 * id; o[2] = c02 * id;
  o[3] = c11 * id; o[4] = c12 * id; o[5] = c22 * id;
}

struct SchurArgs {
  int n_obs;
  int n_lm;
  const int* __restrict__ obs_pose;
  const int* __restrict__ obs_pt;
  const int* __restrict__ lm_start;   // [n_lm+1] observation range of each landmark
  const int* __restrict__ pose_off;
  const double* __restrict__ W;       // [N][18]
  const double* __restrict__ Vg;      // [n_lm][9]
  const double* __restrict__ Vinv;    // [n_lm][6]
  double* __restrict__ S;
  int ldS;
  double* __restrict__ rhs;
};

// One thread per observation i: Y_i = W_i V^-1, then for every observation j of
// the same landmark whose pose block does not lie above i's:  S(i,j) -= Y_i W_j^T.
__global__ void __launch_bounds__(128) schur_kernel(const SchurArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n_obs) return;
  const int q = a.obs_pt[i];
  if (q >= a.n_lm) return;
  const int oi = a.pose_off[a.obs_pose[i]];
  if (oi < 0) return;
  const double* vi = a.Vinv + 6 * (size_t)q;
  const double m00 = vi[0], m01 = vi[1], m02 = vi[2], m11 = vi[3], m12 = vi[4], m22 = vi[5];
  const double* Wi = a.W + 18 * (size_t)i;
  double Y[18];
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    const double w0 = Wi[3 * r], w1 = Wi[3 * r + 1], w2 = Wi[3 * r + 2];
    Y[3 * r + 0] = w0 * m00 + w1 * m01 + w2 * m02;
    Y[3 * r + 1] = w0 * m01 + w1 * m11 + w2 * m12;
    Y[3 * r + 2] = w0 * m02 + w1 * m12 + w2 * m22;
  }
  const double* g = a.Vg + 9 * (size_t)q + 6;
  const double g0 = g[0], g1 = g[1], g2 = g[2];
#pragma unroll
  for (int r = 0; r < 6; ++r) red_add(a.rhs + oi + r, -(Y[3 * r] * g0 + Y[3 * r + 1] * g1 + Y[3 * r + 2] * g2));

  const int j0 = a.lm_start[q], j1 = a.lm_start[q + 1];
  for (int j = j0; j < j1; ++j) {
    const int oj = a.pose_off[a.obs_pose[j]];
    if (oj < 0 || oj > oi) continue;
    const double* Wj = a.W + 18 * (size_t)j;
    double wj[18];
#pragma unroll
    for (int k = 0; k < 18; ++k) wj[k] = Wj[k];
    double* Sd = a.S + (size_t)oi * a.ldS + oj;
    const bool diag = (oj == oi);
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        if (diag && c > r) continue;
        const double v = Y[3 * r] * wj[3 * c] + Y[3 * r + 1] * wj[3 * c + 1] + Y[3 * r + 2] * wj[3 * c + 2];
        red_add(Sd + (size_t)r * a.ldS + c, -v);
      }
  }
}

struct BacksubArgs {
  int n_lm;
  int lm_off;                         // landmark slice of dx starts here
  const int* __restrict__ obs_pose;
  const int* __restrict__ lm_start;
  const int* __restrict__ pose_off;
  const double* __restrict__ W;
  const double* __restrict__ Vg;
  const double* __restrict__ Vinv;
  double* __restrict__ dx;            // [reduced (padded) | 3 n_lm]
};

// One thread per landmark: dx_p = V^-1 (b_p - sum_j W_j^T dx_c(j)).
__global__ void __launch_bounds__(128) backsub_kernel(const BacksubArgs a) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= a.n_lm) return;
  const double* g = a.Vg + 9 * (size_t)q + 6;
  double s0 = g[0], s1 = g[1], s2 = g[2];
  for (int j = a.lm_start[q]; j < a.lm_start[q + 1]; ++j) {
    const int oj = a.pose_off[a.obs_pose[j]];
    if (oj < 0) continue;
    const double* Wj = a.W + 18 * (size_t)j;
    const double* d = a.dx + oj;
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      const double dr = d[r];
      s0 -= Wj[3 * r] * dr;
      s1 -= Wj[3 * r + 1] * dr;
      s2 -= Wj[3 * r + 2] * dr;
    }
  }
  const double* vi = a.Vinv + 6 * (size_t)q;
  double* o = a.dx + a.lm_off + 3 * (size_t)q;
  o[0] = vi[0] * s0 + vi[1] * s1 + vi[2] * s2;
  o[1] = vi[1] * s0 + vi[3] * s1 + vi[4] * s2;
  o[2] = vi[2] * s0 + vi[4] * s1 + vi[5] * s2;
}

}  // namespace bs

// motion_only.cuh -- frame-to-frame (pose-only) reprojection: SURVEY 8 f2.
// Replaces ReprojectionMotionOnlyBatchResidual.evaluate (and the single-point
// ReprojectionMotionOnlyResidual) of pyslam/residuals/reprojection_motion_only_residual.py:36-113
//     pts_2 = T_2_1 pts_1,  r = S (pi(pts_2) - obs_2),  J = S Jpi [I | -pts_2^]          (3N x 6)
// together with the IRLS scaling and the 6 x 6 / 6 x 1 reduction of pyslam/problem.py:329-360: it is the
// camera side of the bundle-adjustment linearisation (reproj_blocks) with the points held fixed, reduced over
// the whole batch: one thread per point (grid-stride), 28 register accumulators, block reduction, one fp64
// atomic per value and CTA.  HBM: 48 B per point (pts_1, obs_2).
#pragma once
#include "common.cuh"
#include "loss.cuh"
#include "reproj.cuh"

namespace bs {

struct MotionArgs {
  int n;
  const double* __restrict__ pts1;   // [n][3] triangulated points in frame 1
  const double* __restrict__ obs2;   // [n][3] observations in frame 2
  ReprojGroup g;                     // camera, stiffness, loss
  const double* __restrict__ pose;   // 12 doubles [R|t] of T_2_1
  int pose_off;                      // reduced offset or -1
  double* __restrict__ S;
  int ldS;
  double* __restrict__ rhs;
  double* __restrict__ scalars;
};

constexpr int kMotionThreads = 256;

template <bool kCostOnly>
__global__ void __launch_bounds__(kMotionThreads) motion_only_kernel(const MotionArgs a, int slot) {
  __shared__ double sred[28][kMotionThreads / 32];
  double P[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) P[k] = a.pose[k];
  double U[28];            // 21 lower-triangle entries of U_c (row-major), 6 of b_c, cost
#pragma unroll
  for (int k = 0; k < 28; ++k) U[k] = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += gridDim.x * blockDim.x) {
    const double X[3] = {ld_stream(a.pts1 + 3 * (size_t)i), ld_stream(a.pts1 + 3 * (size_t)i + 1), ld_stream(a.pts1 + 3 * (size_t)i + 2)};
    const double u = ld_stream(a.obs2 + 3 * (size_t)i), v = ld_stream(a.obs2 + 3 * (size_t)i + 1), d = ld_stream(a.obs2 + 3 * (size_t)i + 2);
    if (kCostOnly) {
      double r[3];
      reproj_residual_only(a.g, P, X, u, v, d, r);
#pragma unroll
      for (int k = 0; k < 3; ++k) U[27] += loss_rho(a.g.loss, r[k]);
      continue;
    }
    ReprojBlocks o;
    reproj_blocks<-1>(a.g, P, X, u, v, d, o);
    U[27] += o.cost;
    // U_c lower triangle (row-major): rows 0-2 M; rows 3-5 [(M B)^T | B^T M B]; then b_c = -[t; B^T t]
    U[0] += o.M[0]; U[1] += o.M[1]; U[2] += o.M[3]; U[3] += o.M[2]; U[4] += o.M[4]; U[5] += o.M[5];
    U[6] += o.MB[0]; U[7] += o.MB[3]; U[8] += o.MB[6]; U[9] += o.BMB[0];
    U[10] += o.MB[1]; U[11] += o.MB[4]; U[12] += o.MB[7]; U[13] += o.BMB[1]; U[14] += o.BMB[3];
    U[15] += o.MB[2]; U[16] += o.MB[5]; U[17] += o.MB[8]; U[18] += o.BMB[2]; U[19] += o.BMB[4]; U[20] += o.BMB[5];
    U[21] -= o.t[0]; U[22] -= o.t[1]; U[23] -= o.t[2];
    U[24] -= o.y * o.t[2] - o.z * o.t[1];
    U[25] -= o.z * o.t[0] - o.x * o.t[2];
    U[26] -= o.x * o.t[1] - o.y * o.t[0];
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 28; ++k) {
    if (kCostOnly && k != 27) continue;
    const double v = warp_sum(U[k]);
    if (lane == 0) sred[k][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < 28) {
    const int k = threadIdx.x;
    if (kCostOnly && k != 27) return;
    double v = 0.0;
#pragma unroll
    for (int w8 = 0; w8 < kMotionThreads / 32; ++w8) v += sred[k][w8];
    if (k == 27) { if (v != 0.0) red_add(a.scalars + slot, v); return; }
    if (a.pose_off < 0 || v == 0.0) return;
    if (k < 21) {
      int rr = 0, base = 0;
      while (base + rr + 1 <= k) { base += rr + 1; ++rr; }
      red_add(a.S + (size_t)(a.pose_off + rr) * a.ldS + a.pose_off + (k - base), v);
    } else {
      red_add(a.rhs + a.pose_off + (k - 21), v);
    }
  }
}

}  // namespace bs

// reproj.cuh -- stereo reprojection blocks: residual, manifold Jacobians,
// robust re-weighting and normal-equation assembly in one pass.
//
// Replaces, per observation, the reference's
//   ReprojectionResidual.evaluate        pyslam/residuals/reprojection_residual.py:13-37
//   StereoCamera.project (+ Jacobian)    pyslam/sensors/stereo_camera.py:100-134
//   SE3.dot / SE3.odot / rot.as_matrix   (liegroups)
//   sqrt(loss.weight) row scaling, cost  pyslam/problem.py:349-360
//   HT.HT^T, -HT.e for these blocks      pyslam/problem.py:329-333
//
// Data layout (all fp64, device resident, built once by Solver::finalize):
//   observations sorted by landmark, SoA:  obs_u/obs_v/obs_d[N], obs_pose[N], obs_pt[N]
//   poses   [K][12] = R row-major | t      (K small: L1/L2 resident)
//   points  [P][3], landmarks to be eliminated first (index < n_lm)
// Outputs per launch:
//   W   [N][18]   J_T^T w J_p   (6x3 row-major)   -- pose/landmark coupling blocks
//   Vg  [n_lm][9] V_p (xx,xy,xz,yy,yz,zz) | b_p   -- landmark blocks
//   S   lower triangle of the dense reduced matrix: U_c added at the pose's offset
//   rhs b_c = -J_T^T w r
//   scalars[COST_LIN] += sum rho(r)
#pragma once
#include "common.cuh"
#include "loss.cuh"

namespace bs {

struct ReprojGroup {     // constants shared by a batch of blocks
  double cu, cv, fu, fv, b;
  double S[9];           // stiffness, row-major
  Loss loss;
};

struct ReprojArgs {
  int n_obs;
  int n_lm;                       // points with index < n_lm are eliminated landmarks
  const double* __restrict__ obs_u;
  const double* __restrict__ obs_v;
  const double* __restrict__ obs_d;
  const int* __restrict__ obs_pose;
  const int* __restrict__ obs_pt;
  const int* __restrict__ obs_grp;          // nullptr when there is a single group
  const ReprojGroup* __restrict__ groups;
  const double* __restrict__ poses;         // [K][12]
  const int* __restrict__ pose_off;         // reduced offset of the pose or -1 (constant)
  const double* __restrict__ pts;           // [P][3]
  double* __restrict__ W;                   // [N][18]
  double* __restrict__ Vg;                  // [n_lm][9]
  double* __restrict__ S;                   // [n_pad][ldS]
  int ldS;
  double* __restrict__ rhs;
  double* __restrict__ scalars;
};

constexpr int kReprojThreads = 256;
constexpr int kWStride = 19;      // 18 doubles + 1 pad: conflict-free 64-bit smem rows

// Residual and the two Jacobians of one observation.
struct ReprojLin {
  double r[3];
  double JT[18];   // 3x6 row-major
  double Jp[9];    // 3x3 row-major
};

BS_D void reproj_residual_only(const ReprojGroup& g, const double* __restrict__ P, const double* __restrict__ X,
                               double u, double v, double d, double* r) {
  const double x = P[0] * X[0] + P[1] * X[1] + P[2] * X[2] + P[9];
  const double y = P[3] * X[0] + P[4] * X[1] + P[5] * X[2] + P[10];
  const double z = P[6] * X[0] + P[7] * X[1] + P[8] * X[2] + P[11];
  const double iz = 1.0 / z;
  const double e0 = g.fu * x * iz + g.cu - u;
  const double e1 = g.fv * y * iz + g.cv - v;
  const double e2 = g.fu * g.b * iz - d;
#pragma unroll
  for (int i = 0; i < 3; ++i) r[i] = g.S[3 * i] * e0 + g.S[3 * i + 1] * e1 + g.S[3 * i + 2] * e2;
}

BS_D void reproj_linearize_one(const ReprojGroup& g, const double* __restrict__ P, const double* __restrict__ X,
                               double u, double v, double d, ReprojLin& L) {
  const double x = P[0] * X[0] + P[1] * X[1] + P[2] * X[2] + P[9];
  const double y = P[3] * X[0] + P[4] * X[1] + P[5] * X[2] + P[10];
  const double z = P[6] * X[0] + P[7] * X[1] + P[8] * X[2] + P[11];
  const double iz = 1.0 / z;
  const double iz2 = iz * iz;
  const double e0 = g.fu * x * iz + g.cu - u;
  const double e1 = g.fv * y * iz + g.cv - v;
  const double e2 = g.fu * g.b * iz - d;
  // camera Jacobian non-zeros (stereo_camera.py:112-134)
  const double j00 = g.fu * iz, j11 = g.fv * iz;
  const double j02 = -g.fu * x * iz2, j12 = -g.fv * y * iz2, j22 = -g.fu * g.b * iz2;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double s0 = g.S[3 * i], s1 = g.S[3 * i + 1], s2 = g.S[3 * i + 2];
    L.r[i] = s0 * e0 + s1 * e1 + s2 * e2;
    // A = S * Jcam (row i)
    const double a0 = s0 * j00, a1 = s1 * j11, a2 = s0 * j02 + s1 * j12 + s2 * j22;
    // J_T = A [I | -p^]
    L.JT[6 * i + 0] = a0;
    L.JT[6 * i + 1] = a1;
    L.JT[6 * i + 2] = a2;
    L.JT[6 * i + 3] = a2 * y - a1 * z;
    L.JT[6 * i + 4] = a0 * z - a2 * x;
    L.JT[6 * i + 5] = a1 * x - a0 * y;
    // J_p = A R
    L.Jp[3 * i + 0] = a0 * P[0] + a1 * P[3] + a2 * P[6];
    L.Jp[3 * i + 1] = a0 * P[1] + a1 * P[4] + a2 * P[7];
    L.Jp[3 * i + 2] = a0 * P[2] + a1 * P[5] + a2 * P[8];
  }
}

// One thread per observation.  W is staged through shared memory so the
// 144-byte rows leave the SM as fully coalesced 8-byte-per-lane stores.
__global__ void __launch_bounds__(kReprojThreads)
reproj_linearize_kernel(const ReprojArgs a) {
  __shared__ double sW[kReprojThreads * kWStride];
  __shared__ double sred[kReprojThreads / 32];
  const int tid = threadIdx.x;
  const int base = blockIdx.x * kReprojThreads;
  const int i = base + tid;
  double cost = 0.0;
  double w18[18];
#pragma unroll
  for (int k = 0; k < 18; ++k) w18[k] = 0.0;

  if (i < a.n_obs) {
    const int pi = a.obs_pose[i];
    const int qi = a.obs_pt[i];
    const int poff = a.pose_off[pi];
    const bool pt_var = qi < a.n_lm;
    if (poff >= 0 || pt_var) {
      const ReprojGroup& g = a.groups[a.obs_grp ? a.obs_grp[i] : 0];
      const double* P = a.poses + 12 * (size_t)pi;
      const double* X = a.pts + 3 * (size_t)qi;
      ReprojLin L;
      reproj_linearize_one(g, P, X, a.obs_u[i], a.obs_v[i], a.obs_d[i], L);
      double w[3], wr[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        w[k] = loss_weight(g.loss, L.r[k]);
        wr[k] = w[k] * L.r[k];
        cost += loss_rho(g.loss, L.r[k]);
      }
      if (poff >= 0) {
        // U_c (lower triangle incl. diagonal) and b_c
        double* Sd = a.S + (size_t)poff * a.ldS + poff;
#pragma unroll
        for (int r = 0; r < 6; ++r) {
#pragma unroll
          for (int c = 0; c <= r; ++c) {
            const double v = w[0] * L.JT[r] * L.JT[c] + w[1] * L.JT[6 + r] * L.JT[6 + c] +
                             w[2] * L.JT[12 + r] * L.JT[12 + c];
            red_add(Sd + (size_t)r * a.ldS + c, v);
          }
          red_add(a.rhs + poff + r, -(L.JT[r] * wr[0] + L.JT[6 + r] * wr[1] + L.JT[12 + r] * wr[2]));
        }
      }
      if (pt_var) {
        double* vg = a.Vg + 9 * (size_t)qi;
        int k = 0;
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int c = r; c < 3; ++c)
            red_add(vg + (k++), w[0] * L.Jp[r] * L.Jp[c] + w[1] * L.Jp[3 + r] * L.Jp[3 + c] +
                                    w[2] * L.Jp[6 + r] * L.Jp[6 + c]);
#pragma unroll
        for (int r = 0; r < 3; ++r)
          red_add(vg + 6 + r, -(L.Jp[r] * wr[0] + L.Jp[3 + r] * wr[1] + L.Jp[6 + r] * wr[2]));
      }
      if (poff >= 0 && pt_var) {
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
          for (int c = 0; c < 3; ++c)
            w18[3 * r + c] = w[0] * L.JT[r] * L.Jp[c] + w[1] * L.JT[6 + r] * L.Jp[3 + c] +
                             w[2] * L.JT[12 + r] * L.Jp[6 + c];
      }
    }
  }
  // stage W rows, then stream them out coalesced
#pragma unroll
  for (int k = 0; k < 18; ++k) sW[tid * kWStride + k] = w18[k];
  __syncthreads();
  const int n_here = min(kReprojThreads, a.n_obs - base);
  double* Wg = a.W + 18 * (size_t)base;
  for (int e = tid; e < 18 * n_here; e += kReprojThreads) {
    const int o = e / 18, k = e - 18 * o;
    Wg[e] = sW[o * kWStride + k];
  }
  block_sum_to(cost, a.scalars + 0 /*COST_LIN*/, sred);
}

// Cost only: sum rho(r) over ALL reprojection blocks (Problem.eval_cost,
// pyslam/problem.py:110-128) -> scalars[slot].
__global__ void __launch_bounds__(kReprojThreads)
reproj_cost_kernel(const ReprojArgs a, int slot) {
  __shared__ double sred[kReprojThreads / 32];
  double cost = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n_obs; i += gridDim.x * blockDim.x) {
    const ReprojGroup& g = a.groups[a.obs_grp ? a.obs_grp[i] : 0];
    double r[3];
    reproj_residual_only(g, a.poses + 12 * (size_t)a.obs_pose[i], a.pts + 3 * (size_t)a.obs_pt[i],
                         a.obs_u[i], a.obs_v[i], a.obs_d[i], r);
#pragma unroll
    for (int k = 0; k < 3; ++k) cost += loss_rho(g.loss, r[k]);
  }
  block_sum_to(cost, a.scalars + slot, sred);
}

}  // namespace bs

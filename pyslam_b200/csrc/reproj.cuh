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
//   points  [P][3], landmarks to be eliminated first (index < n_lm), ordered by the
//           first pose that sees them so that neighbouring landmarks share cameras
//   "landmark blocks": runs of whole landmarks with <= 128 observations; per block the
//           distinct variable poses ("slots"), a per-observation slot id and the
//           block's observations grouped by slot (cam_perm / seg_start)
// Outputs per launch:
//   W   [18][N]   J_T^T w J_p (6x3 row-major index k = 3r+c), SoA planes of N doubles
//   Vg  [n_lm][9] V_p (xx,xy,xz,yy,yz,zz) | b_p   -- landmark blocks
//   S   lower triangle of the dense reduced matrix: U_c added at the pose's offset
//   rhs b_c = -J_T^T w r
//   scalars[COST_LIN] += sum rho(r)
//
// reproj_block_kernel (the fast path): one CTA per landmark block, one thread per
// observation.  Each thread leaves its 27 camera values (U_c lower triangle, b_c)
// and 9 landmark values (V_p, b_p) in a shared-memory row; the CTA then reduces
// them per slot / per landmark, so HBM sees one fp64 atomic per (slot, value) and
// a plain store per landmark value instead of 36 atomics per observation.
// reproj_generic_kernel: same arithmetic with global atomics for the tail
// (landmarks with more than 128 observations, observations of constant points).
#pragma once
#include "common.cuh"
#include "loss.cuh"

namespace bs {

struct ReprojGroup {     // constants shared by a batch of blocks
  double cu, cv, fu, fv, b;
  double S[9];           // stiffness, row-major
  Loss loss;
};

struct LmBlock {         // a run of whole landmarks
  int obs_begin, n_obs;  // n_obs <= kBlkObs
  int lm_begin, n_lms;
  int slot_begin, n_slots;   // into slot_pose[] / (slot_begin + block index) into seg_start[]
  int seg_begin;             // into seg_start[]: n_slots + 1 entries (local positions in cam_perm order)
  int pad;
};

struct ReprojArgs {
  int n_obs;                      // all observations (SoA stride of W)
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
  const int* __restrict__ lm_start;         // [n_lm+1]
  // landmark blocks
  int n_blocks;
  const LmBlock* __restrict__ blocks;
  const int* __restrict__ slot_pose;
  const unsigned char* __restrict__ cam_perm;   // [N] local obs index, grouped by slot inside each block
  const unsigned char* __restrict__ seg_start;
  // tail processed by the generic kernel
  int tail_begin;
  double* __restrict__ W;                   // [18][N]
  double* __restrict__ Vg;                  // [n_lm][9]
  double* __restrict__ S;                   // [n_pad][ldS]
  int ldS;
  double* __restrict__ rhs;
  double* __restrict__ scalars;
};

constexpr int kBlkObs = 128;      // observations per landmark block = threads per CTA
constexpr int kMaxTrack = 64;     // longer tracks go through the generic (atomic) kernels
constexpr int kSchurCap = 768;    // n_slots * ldk cap of a multi-landmark block (96 KB of Schur operands)
constexpr int kRow = 37;          // 27 camera + 9 landmark values + 1 pad (odd stride: conflict-free rows)

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

// value index 0..20 -> (row, col) of the lower triangle of a 6x6 block, row-major
__device__ __constant__ unsigned char kTriRow[21] = {0, 1, 1, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 4, 5, 5, 5, 5, 5, 5};
__device__ __constant__ unsigned char kTriCol[21] = {0, 0, 1, 0, 1, 2, 0, 1, 2, 3, 0, 1, 2, 3, 4, 0, 1, 2, 3, 4, 5};

__global__ void __launch_bounds__(kBlkObs, 4)
reproj_block_kernel(const ReprojArgs a) {
  __shared__ double sT[kBlkObs * kRow];
  __shared__ double sred[kBlkObs / 32];
  const int tid = threadIdx.x;
  const LmBlock blk = a.blocks[blockIdx.x];
  const int N = a.n_obs;
  double cost = 0.0;

  if (tid < blk.n_obs) {
    const int i = blk.obs_begin + tid;
    const int pi = a.obs_pose[i];
    const int qi = a.obs_pt[i];
    const bool pose_var = a.pose_off[pi] >= 0;
    const ReprojGroup& g = a.groups[a.obs_grp ? a.obs_grp[i] : 0];
    ReprojLin L;
    reproj_linearize_one(g, a.poses + 12 * (size_t)pi, a.pts + 3 * (size_t)qi, ld_stream(a.obs_u + i),
                         ld_stream(a.obs_v + i), ld_stream(a.obs_d + i), L);
    double w[3], wr[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      w[k] = loss_weight(g.loss, L.r[k]);
      wr[k] = w[k] * L.r[k];
      cost += loss_rho(g.loss, L.r[k]);
    }
    double* row = sT + tid * kRow;
    // weighted Jacobian rows, reused by all three products
    double wJT[18];
#pragma unroll
    for (int k = 0; k < 18; ++k) wJT[k] = w[k / 6] * L.JT[k];
    if (pose_var) {
      int v = 0;
#pragma unroll
      for (int r = 0; r < 6; ++r) {
#pragma unroll
        for (int c = 0; c <= r; ++c)
          row[v++] = wJT[r] * L.JT[c] + wJT[6 + r] * L.JT[6 + c] + wJT[12 + r] * L.JT[12 + c];
      }
#pragma unroll
      for (int r = 0; r < 6; ++r) row[21 + r] = -(L.JT[r] * wr[0] + L.JT[6 + r] * wr[1] + L.JT[12 + r] * wr[2]);
      // W = J_T^T w J_p, SoA planes
      double* Wp = a.W + i;
#pragma unroll
      for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c)
          Wp[(size_t)(3 * r + c) * N] = wJT[r] * L.Jp[c] + wJT[6 + r] * L.Jp[3 + c] + wJT[12 + r] * L.Jp[6 + c];
    }
    {
      int v = 27;
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = r; c < 3; ++c)
          row[v++] = w[0] * L.Jp[r] * L.Jp[c] + w[1] * L.Jp[3 + r] * L.Jp[3 + c] + w[2] * L.Jp[6 + r] * L.Jp[6 + c];
#pragma unroll
      for (int r = 0; r < 3; ++r) row[33 + r] = -(L.Jp[r] * wr[0] + L.Jp[3 + r] * wr[1] + L.Jp[6 + r] * wr[2]);
    }
  }
  __syncthreads();

  // camera side: one task per (slot, value); observations of a slot are contiguous in cam_perm order
  {
    const unsigned char* perm = a.cam_perm + blk.obs_begin;
    const unsigned char* seg = a.seg_start + blk.seg_begin;
    const int n_tasks = blk.n_slots * 27;
    for (int t = tid; t < n_tasks; t += kBlkObs) {
      const int s = t / 27, v = t - 27 * s;
      double acc = 0.0;
      for (int k = seg[s]; k < seg[s + 1]; ++k) acc += sT[perm[k] * kRow + v];
      const int off = a.pose_off[a.slot_pose[blk.slot_begin + s]];
      if (v < 21) red_add(a.S + (size_t)(off + kTriRow[v]) * a.ldS + off + kTriCol[v], acc);
      else red_add(a.rhs + off + (v - 21), acc);
    }
  }
  // landmark side: one task per (landmark, value); a landmark's observations are contiguous
  {
    const int n_tasks = blk.n_lms * 9;
    for (int t = tid; t < n_tasks; t += kBlkObs) {
      const int l = t / 9, v = t - 9 * l;
      const int q = blk.lm_begin + l;
      const int k0 = a.lm_start[q] - blk.obs_begin, k1 = a.lm_start[q + 1] - blk.obs_begin;
      double acc = 0.0;
      for (int k = k0; k < k1; ++k) acc += sT[k * kRow + 27 + v];
      a.Vg[9 * (size_t)q + v] = acc;
    }
  }
  block_sum_to(cost, a.scalars + 0 /*COST_LIN*/, sred);
}

// Tail observations [tail_begin, n_obs): one thread per observation, global atomics.
__global__ void __launch_bounds__(128)
reproj_generic_kernel(const ReprojArgs a) {
  __shared__ double sred[4];
  const int i = a.tail_begin + blockIdx.x * blockDim.x + threadIdx.x;
  const int N = a.n_obs;
  double cost = 0.0;
  if (i < N) {
    const int pi = a.obs_pose[i];
    const int qi = a.obs_pt[i];
    const int poff = a.pose_off[pi];
    const bool pt_var = qi < a.n_lm;
    if (poff >= 0 || pt_var) {
      const ReprojGroup& g = a.groups[a.obs_grp ? a.obs_grp[i] : 0];
      ReprojLin L;
      reproj_linearize_one(g, a.poses + 12 * (size_t)pi, a.pts + 3 * (size_t)qi, a.obs_u[i], a.obs_v[i], a.obs_d[i], L);
      double w[3], wr[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        w[k] = loss_weight(g.loss, L.r[k]);
        wr[k] = w[k] * L.r[k];
        cost += loss_rho(g.loss, L.r[k]);
      }
      if (poff >= 0) {
        double* Sd = a.S + (size_t)poff * a.ldS + poff;
#pragma unroll
        for (int r = 0; r < 6; ++r) {
#pragma unroll
          for (int c = 0; c <= r; ++c)
            red_add(Sd + (size_t)r * a.ldS + c, w[0] * L.JT[r] * L.JT[c] + w[1] * L.JT[6 + r] * L.JT[6 + c] +
                                                    w[2] * L.JT[12 + r] * L.JT[12 + c]);
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
        double* Wp = a.W + i;
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
          for (int c = 0; c < 3; ++c)
            Wp[(size_t)(3 * r + c) * N] = w[0] * L.JT[r] * L.Jp[c] + w[1] * L.JT[6 + r] * L.Jp[3 + c] +
                                          w[2] * L.JT[12 + r] * L.Jp[6 + c];
      }
    }
  }
  block_sum_to(cost, a.scalars + 0 /*COST_LIN*/, sred);
}

// Cost only: sum rho(r) over ALL reprojection blocks (Problem.eval_cost,
// pyslam/problem.py:110-128) -> scalars[slot].
__global__ void __launch_bounds__(256)
reproj_cost_kernel(const ReprojArgs a, int slot) {
  __shared__ double sred[8];
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

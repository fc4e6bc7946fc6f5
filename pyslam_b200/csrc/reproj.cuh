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
// Data layout (all fp64, device resident, built once by bslam_finalize):
//   poses   [K][12] = R row-major | t      (K small: L1/L2 resident)
//   points  [P][3], landmarks to be eliminated first (index < n_lm), ordered by the
//           first pose that sees them so that neighbouring landmarks share cameras
//   "landmark blocks": runs of whole landmarks with <= 128 observations; per block the
//           distinct variable poses ("slots"), seg_start (first observation of every slot)
//   observations, SoA: obs_u/obs_v/obs_d[N], obs_code[N] (slot | block-local landmark | group),
//           obs_pose[N], obs_pt[N]; SLOT-MAJOR inside each block (grouped by pose);
//           lm_start (CSR over landmarks) + lm_obs / lm_obs_local give the landmark order
// Outputs per launch:
//   W   tiled (common.cuh: w_index)  J_T^T w J_p (6x3 row-major index k = 3r+c) per observation
//   Vg  [n_lm][9] V_p (xx,xy,xz,yy,yz,zz) | b_p   -- landmark blocks
//   S   lower triangle of the dense reduced matrix: U_c added at the pose's offset
//   rhs b_c = -J_T^T w r
//   scalars[COST_LIN] += sum rho(r)
//
// reproj_block_kernel (the fast path): persistent CTAs walking landmark blocks, one thread
// per observation.  Each thread leaves its 27 camera values (U_c lower triangle, b_c)
// and 9 landmark values (V_p, b_p) in a shared-memory row; the CTA then reduces
// them per slot / per landmark, so HBM sees one fp64 atomic per (slot, value) and
// a plain store per landmark value instead of 36 atomics per observation.
// reproj_generic_kernel: same arithmetic with global atomics for the tail (landmarks with
// more than 64 observations or two observations by one pose, observations of points that
// are not eliminated).
#pragma once
#include "common.cuh"
#include "loss.cuh"

namespace bs {

struct ReprojGroup {     // constants shared by a batch of blocks
  double cu, cv, fu, fv, b;   // b > 0: stereo camera (third measurement = disparity fu b / z); b <= 0: RGB-D camera (= depth z)
  double S[9];           // stiffness, row-major
  Loss loss;
};

struct LmBlock {         // a run of whole landmarks; its observations are stored SLOT-MAJOR (grouped by pose)
  int obs_begin, n_obs;  // n_obs <= kBlkObs
  int lm_begin, n_lms;
  int slot_begin, n_slots;   // into slot_pose[] / slot_off[] / slot_poses[]
  int seg_begin;             // into seg_start[]: n_slots + 1 entries (first observation of every slot, block-local)
  int pad;
};

struct ReprojArgs {
  int n_obs;                      // all observations
  int n_lm;                       // points with index < n_lm are eliminated landmarks
  const double* __restrict__ obs_u;
  const double* __restrict__ obs_v;
  const double* __restrict__ obs_d;
  const int* __restrict__ obs_pose;
  const int* __restrict__ obs_pt;
  const int* __restrict__ obs_grp;          // nullptr when there is a single group
  const ReprojGroup* __restrict__ groups;
  ReprojGroup g0;                           // groups[0] by value (constant bank) for the single-group fast path
  const unsigned* __restrict__ obs_code;    // [N] slot (bits 0-7, 255: constant pose) | block-local landmark (8-15) | group (16-31)
  const double* __restrict__ poses;         // [K][12]
  const int* __restrict__ pose_off;         // reduced offset of the pose or -1 (constant)
  const double* __restrict__ pts;           // [P][3]
  const int* __restrict__ lm_start;         // [n_lm+1] CSR over landmarks
  const int* __restrict__ lm_obs;           // [N] CSR position -> observation index (identity in the tail)
  // landmark blocks
  int n_blocks;
  const LmBlock* __restrict__ blocks;
  const unsigned char* __restrict__ lm_obs_local;   // [N] CSR position -> block-local observation index
  const unsigned char* __restrict__ seg_start;
  const int* __restrict__ slot_off;             // pose_off[slot_pose[e]] per slot entry (static)
  const double* __restrict__ slot_poses;        // [n_slot_entries][12], gathered per linearisation
  int stage_len;                                // doubles per staging buffer (max over blocks of 12 n_slots + 3 n_lms)
  // tail processed by the generic kernel
  int tail_begin;
  double* __restrict__ W;                   // tiled, see w_index()
  double* __restrict__ Vg;                  // [n_lm][9]
  double* __restrict__ S;                   // [n_pad][ldS]
  int ldS;
  double* __restrict__ rhs;
  double* __restrict__ scalars;
};

constexpr int kBlkObs = 128;      // observations per landmark block = threads per CTA
constexpr int kMaxTrack = 64;     // longer tracks go through the generic (atomic) kernels

constexpr int kRow = 38;          // 27 camera + 9 landmark values + 2 pad: 16-byte aligned rows whose 128-bit
                                  // stores are bank-conflict free (row stride = 12 banks mod 32)

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
  const double iz = fast_rcp(z);
  const double e0 = g.fu * x * iz + g.cu - u;
  const double e1 = g.fv * y * iz + g.cv - v;
  const double e2 = (g.b > 0.0 ? g.fu * g.b * iz : z) - d;       // b <= 0: RGB-D pinhole model, third measurement = depth
#pragma unroll
  for (int i = 0; i < 3; ++i) r[i] = g.S[3 * i] * e0 + g.S[3 * i + 1] * e1 + g.S[3 * i + 2] * e2;
}

BS_D void reproj_linearize_one(const ReprojGroup& g, const double* __restrict__ P, const double* __restrict__ X,
                               double u, double v, double d, ReprojLin& L) {
  const double x = P[0] * X[0] + P[1] * X[1] + P[2] * X[2] + P[9];
  const double y = P[3] * X[0] + P[4] * X[1] + P[5] * X[2] + P[10];
  const double z = P[6] * X[0] + P[7] * X[1] + P[8] * X[2] + P[11];
  const double iz = fast_rcp(z);
  const double iz2 = iz * iz;
  const double e0 = g.fu * x * iz + g.cu - u;
  const double e1 = g.fv * y * iz + g.cv - v;
  const bool stereo = g.b > 0.0;
  const double e2 = (stereo ? g.fu * g.b * iz : z) - d;
  // camera Jacobian non-zeros (stereo_camera.py:112-134; rgbd_camera.py:127-146: last row [0 0 1])
  const double j00 = g.fu * iz, j11 = g.fv * iz;
  const double j02 = -g.fu * x * iz2, j12 = -g.fv * y * iz2, j22 = stereo ? -g.fu * g.b * iz2 : 1.0;
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

// Structured linearisation of one observation.  With e = pi(p_c) - z, r = S e,
// w = loss weights, Q = S^T diag(w) S, Jc the (sparse) camera Jacobian and
// B = -p_c^, the reference's J_T = S Jc [I | B] and J_p = S Jc R give
//     M = Jc^T Q Jc  (3x3 symmetric),   t = Jc^T Q e,
//     U_c = [I|B]^T M [I|B],  b_c = -[I|B]^T t,
//     V_p = R^T M R,          b_p = -R^T t,        W = [I|B]^T M R,
// which needs ~40% of the flops of forming the 3x6 / 3x3 Jacobians and their
// weighted outer products entry by entry (same values up to rounding).
struct ReprojBlocks {
  double M[6];     // xx xy xz yy yz zz
  double t[3];
  double MB[9];    // M B   (row-major 3x3)
  double BMB[6];   // B^T M B (xx xy xz yy yz zz)
  double MR[9];    // M R   (row-major 3x3)
  double x, y, z;  // p_c
  double cost;
};

template <int kLoss>
BS_D void reproj_blocks(const ReprojGroup& g, const double* __restrict__ P, const double* __restrict__ X,
                        double u, double v, double d, ReprojBlocks& o) {
  const double x = P[0] * X[0] + P[1] * X[1] + P[2] * X[2] + P[9];
  const double y = P[3] * X[0] + P[4] * X[1] + P[5] * X[2] + P[10];
  const double z = P[6] * X[0] + P[7] * X[1] + P[8] * X[2] + P[11];
  o.x = x; o.y = y; o.z = z;
  const double iz = fast_rcp(z);
  const double iz2 = iz * iz;
  const double e0 = g.fu * x * iz + g.cu - u;
  const double e1 = g.fv * y * iz + g.cv - v;
  const bool stereo = g.b > 0.0;                                  // b <= 0: RGB-D pinhole model (rgbd_camera.py:113-146)
  const double e2 = (stereo ? g.fu * g.b * iz : z) - d;
  // residual, weights, cost;  Q = S^T diag(w) S
  double q00 = 0, q01 = 0, q02 = 0, q11 = 0, q12 = 0, q22 = 0;
  double cost = 0.0;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double s0 = g.S[3 * k], s1 = g.S[3 * k + 1], s2 = g.S[3 * k + 2];
    const double r = s0 * e0 + s1 * e1 + s2 * e2;
    const double w = loss_weight_t<kLoss>(g.loss, r);
    cost += loss_rho_t<kLoss>(g.loss, r);
    const double ws0 = w * s0, ws1 = w * s1, ws2 = w * s2;
    q00 = fma(ws0, s0, q00); q01 = fma(ws0, s1, q01); q02 = fma(ws0, s2, q02);
    q11 = fma(ws1, s1, q11); q12 = fma(ws1, s2, q12); q22 = fma(ws2, s2, q22);
  }
  o.cost = cost;
  const double qe0 = q00 * e0 + q01 * e1 + q02 * e2;
  const double qe1 = q01 * e0 + q11 * e1 + q12 * e2;
  const double qe2 = q02 * e0 + q12 * e1 + q22 * e2;
  // camera Jacobian non-zeros (stereo_camera.py:112-134): [[a,0,c0],[0,b,c1],[0,0,c2]]
  const double a = g.fu * iz, b = g.fv * iz;
  const double c0 = -g.fu * x * iz2, c1 = -g.fv * y * iz2, c2 = stereo ? -g.fu * g.b * iz2 : 1.0;
  // QJ = Q Jc (only the entries M needs), M = Jc^T QJ
  const double k02 = c0 * q00 + c1 * q01 + c2 * q02;
  const double k12 = c0 * q01 + c1 * q11 + c2 * q12;
  const double k22 = c0 * q02 + c1 * q12 + c2 * q22;
  const double m00 = a * a * q00, m01 = a * b * q01, m02 = a * k02;
  const double m11 = b * b * q11, m12 = b * k12;
  const double m22 = c0 * k02 + c1 * k12 + c2 * k22;
  o.M[0] = m00; o.M[1] = m01; o.M[2] = m02; o.M[3] = m11; o.M[4] = m12; o.M[5] = m22;
  o.t[0] = a * qe0; o.t[1] = b * qe1; o.t[2] = c0 * qe0 + c1 * qe1 + c2 * qe2;
  // M B with B = [[0, z, -y], [-z, 0, x], [y, -x, 0]]
  const double Mr[9] = {m00, m01, m02, m01, m11, m12, m02, m12, m22};
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    o.MB[3 * i + 0] = y * Mr[3 * i + 2] - z * Mr[3 * i + 1];
    o.MB[3 * i + 1] = z * Mr[3 * i + 0] - x * Mr[3 * i + 2];
    o.MB[3 * i + 2] = x * Mr[3 * i + 1] - y * Mr[3 * i + 0];
  }
  // B^T (M B): rows of B^T are [0,-z,y], [z,0,-x], [-y,x,0]
  o.BMB[0] = y * o.MB[6] - z * o.MB[3];
  o.BMB[1] = y * o.MB[7] - z * o.MB[4];
  o.BMB[2] = y * o.MB[8] - z * o.MB[5];
  o.BMB[3] = z * o.MB[1] - x * o.MB[7];
  o.BMB[4] = z * o.MB[2] - x * o.MB[8];
  o.BMB[5] = x * o.MB[5] - y * o.MB[2];
  // M R
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      o.MR[3 * i + j] = Mr[3 * i] * P[j] + Mr[3 * i + 1] * P[3 + j] + Mr[3 * i + 2] * P[6 + j];
}

// Only what the back-substitution needs of an observation: M = Jc^T Q Jc (xx xy xz yy yz zz) and p_c, so that
//   W^T dx_c = (M R)^T (d rho + B d phi) = R^T (M e),  e = d rho + B d phi
// costs one 3x3 symmetric product instead of forming M R.  Same arithmetic for M as reproj_blocks.
template <int kLoss>
BS_D void reproj_M(const ReprojGroup& g, const double* __restrict__ P, const double* __restrict__ X, double u, double v, double d,
                   double* __restrict__ M, double& x, double& y, double& z) {
  x = P[0] * X[0] + P[1] * X[1] + P[2] * X[2] + P[9];
  y = P[3] * X[0] + P[4] * X[1] + P[5] * X[2] + P[10];
  z = P[6] * X[0] + P[7] * X[1] + P[8] * X[2] + P[11];
  const double iz = fast_rcp(z);
  const double iz2 = iz * iz;
  const bool stereo = g.b > 0.0;
  const double e0 = g.fu * x * iz + g.cu - u;
  const double e1 = g.fv * y * iz + g.cv - v;
  const double e2 = (stereo ? g.fu * g.b * iz : z) - d;
  double q00 = 0, q01 = 0, q02 = 0, q11 = 0, q12 = 0, q22 = 0;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double s0 = g.S[3 * k], s1 = g.S[3 * k + 1], s2 = g.S[3 * k + 2];
    const double r = s0 * e0 + s1 * e1 + s2 * e2;
    const double w = loss_weight_t<kLoss>(g.loss, r);
    const double ws0 = w * s0, ws1 = w * s1, ws2 = w * s2;
    q00 = fma(ws0, s0, q00); q01 = fma(ws0, s1, q01); q02 = fma(ws0, s2, q02);
    q11 = fma(ws1, s1, q11); q12 = fma(ws1, s2, q12); q22 = fma(ws2, s2, q22);
  }
  const double a = g.fu * iz, b = g.fv * iz;
  const double c0 = -g.fu * x * iz2, c1 = -g.fv * y * iz2, c2 = stereo ? -g.fu * g.b * iz2 : 1.0;
  const double k02 = c0 * q00 + c1 * q01 + c2 * q02;
  const double k12 = c0 * q01 + c1 * q11 + c2 * q12;
  const double k22 = c0 * q02 + c1 * q12 + c2 * q22;
  M[0] = a * a * q00; M[1] = a * b * q01; M[2] = a * k02;
  M[3] = b * b * q11; M[4] = b * k12;
  M[5] = c0 * k02 + c1 * k12 + c2 * k22;
}

// ---- cp.async helpers (LDGSTS): global -> shared without staging registers -------------------
BS_D void cp_async8(void* smem_dst, const void* gsrc) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gsrc) : "memory");
}
BS_D void cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc) : "memory");
}
BS_D void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
BS_D void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Software-pipelined persistent kernel.  A CTA walks landmark blocks b, b + grid, ...  While block i is
// being processed, everything block i+1 needs is already in flight: its descriptor (cp.async, two blocks
// ahead), its slot poses and landmark coordinates (cp.async into the other half of a double-buffered
// staging area) and its per-observation inputs (registers, untouched until the next round so that no
// load is waited for early).
//   phase 1  one thread per observation (slot-major order, so the pose reads of a warp are shared-memory
//            broadcasts): structured linearisation, nine 16-byte W stores into the tiled layout, 36 values
//            -> a shared row (eighteen 16-byte stores)
//   phase 2  warp-parallel reductions: a warp owns a slot (lanes = the 27 camera values, the slot's rows
//            are contiguous) or three landmarks (lanes = 3 x 9 landmark values, rows gathered through the
//            block's landmark -> observation list); one fp64 atomic per (slot, value) leaves the SM,
//            V_p / b_p are plain stores.
// kLoss >= 0: single group with that loss kind (compile time); kLoss < 0: per-observation groups.
#ifndef BSLAM_REPROJ_CTAS
#define BSLAM_REPROJ_CTAS 5
#endif
constexpr int kReprojCtas = BSLAM_REPROJ_CTAS;     // resident CTAs per SM (register budget = 65536 / (128 * this))
template <int kLoss>
__global__ void __launch_bounds__(kBlkObs, kReprojCtas)
reproj_block_kernel(const ReprojArgs a) {
  extern __shared__ __align__(16) double sStage[];    // 2 x stage_len doubles: [slot poses | landmark points]
  __shared__ __align__(16) double sT[kBlkObs * kRow];
  __shared__ double sred[kBlkObs / 32];
  __shared__ __align__(16) LmBlock sDesc[3];          // descriptor ring
  __shared__ unsigned char sLmObs[kBlkObs], sSeg[kBlkObs + 4], sLmStart[kBlkObs + 4];
  __shared__ int sOff[kBlkObs];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int stride = gridDim.x;
  double cost = 0.0;
  int b = blockIdx.x;
  if (b >= a.n_blocks) return;
  // lane -> position of its camera value inside the pose's diagonal block of S (lower triangle, row-major)
  int tri_off;
  {
    int r = 0;
    while ((r + 1) * (r + 2) / 2 <= lane) ++r;
    tri_off = lane < 21 ? r * a.ldS + (lane - r * (r + 1) / 2) : 0;
  }
  const int lm_sub = lane / 9, lm_v = lane - 9 * lm_sub;    // landmark-side role of the lane (lanes 27..31 idle)

  auto stage = [&](const LmBlock& d, double* dst) {    // async copy of a block's poses and points
    const int np2 = 6 * d.n_slots, nq = 3 * d.n_lms;
    const double* gp = a.slot_poses + 12 * (size_t)d.slot_begin;
    const double* gq = a.pts + 3 * (size_t)d.lm_begin;
    for (int e = tid; e < np2; e += kBlkObs) cp_async16(dst + 2 * e, gp + 2 * e);
    for (int e = tid; e < nq; e += kBlkObs) cp_async8(dst + 2 * np2 + e, gq + e);
  };
  auto fetch_desc = [&](int blk_id, int slot) {        // 32-byte descriptor, two 16-byte async copies
    if (tid < 2 && blk_id < a.n_blocks)
      cp_async16(reinterpret_cast<char*>(&sDesc[slot]) + 16 * tid, reinterpret_cast<const char*>(a.blocks + blk_id) + 16 * tid);
  };
  // per-observation inputs and the block's side tables, as loaded (nothing derived: deriving would wait)
  struct Inputs { double u, v, d; unsigned code; int lmobs, lmstart, seg, soff; };
  auto load_inputs = [&](const LmBlock& d, Inputs& in) {
    in.u = in.v = in.d = 0.0; in.code = 255u; in.lmobs = 0; in.lmstart = 0; in.seg = 0; in.soff = 0;
    if (tid < d.n_obs) {
      const int i = d.obs_begin + tid;
      in.u = ld_stream(a.obs_u + i); in.v = ld_stream(a.obs_v + i); in.d = ld_stream(a.obs_d + i);
      in.code = ld_stream(a.obs_code + i);
      in.lmobs = a.lm_obs_local[i];
    }
    if (tid <= d.n_slots) in.seg = a.seg_start[d.seg_begin + tid];
    if (tid < d.n_slots) in.soff = a.slot_off[d.slot_begin + tid];
    if (tid <= d.n_lms) in.lmstart = a.lm_start[d.lm_begin + tid];
  };

  // ---- prologue: block b synchronously, descriptors of b + stride and b + 2 stride in flight
  LmBlock blk = a.blocks[b];
  fetch_desc(b + stride, 1);
  fetch_desc(b + 2 * stride, 2);
  stage(blk, sStage);
  cp_async_commit();
  Inputs in;
  load_inputs(blk, in);

  int it = 0;
  for (;;) {
    double* cur = sStage + (it & 1) * a.stage_len;
    double* nxt = sStage + ((it + 1) & 1) * a.stage_len;
    const int bn = b + stride;
    const bool has_next = bn < a.n_blocks;
    sLmObs[tid] = (unsigned char)in.lmobs; sSeg[tid] = (unsigned char)in.seg; sOff[tid] = in.soff;
    sLmStart[tid] = (unsigned char)(in.lmstart - blk.obs_begin);
    if (tid == 0) { sSeg[kBlkObs] = (unsigned char)blk.n_obs; sLmStart[kBlkObs] = (unsigned char)blk.n_obs; }   // n_slots / n_lms == 128
    cp_async_wait_all();
    __syncthreads();                                   // staging of this block and the next descriptor have landed
    // ---- everything the NEXT block needs goes in flight now
    Inputs nin = in;
    LmBlock nblk = blk;
    if (has_next) {
      nblk = sDesc[(it + 1) % 3];
      stage(nblk, nxt);
      load_inputs(nblk, nin);
    }
    fetch_desc(b + 3 * stride, it % 3);                // slot of the current block's descriptor is free again
    cp_async_commit();

    // ---- phase 1
    if (tid < blk.n_obs) {
      const int i = blk.obs_begin + tid;
      const int sl = in.code & 255, ql = (in.code >> 8) & 255;
      double P[12], X[3];
      if (sl != 255) {
        const double2* Ps = reinterpret_cast<const double2*>(cur + 12 * sl);
#pragma unroll
        for (int k = 0; k < 6; ++k) { const double2 t = Ps[k]; P[2 * k] = t.x; P[2 * k + 1] = t.y; }
      } else {                                         // constant pose: not a slot, read it directly
        const double* Pg = a.poses + 12 * (size_t)a.obs_pose[i];
#pragma unroll
        for (int k = 0; k < 12; ++k) P[k] = Pg[k];
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) X[k] = cur[12 * blk.n_slots + 3 * ql + k];
      const ReprojGroup& grp = kLoss >= 0 ? a.g0 : a.groups[in.code >> 16];
      ReprojBlocks o;
      reproj_blocks<kLoss>(grp, P, X, in.u, in.v, in.d, o);
      cost += o.cost;
      double2* row = reinterpret_cast<double2*>(sT + tid * kRow);
      if (sl != 255) {
        // U_c lower triangle (row-major), rows 0-2: M; rows 3-5: [(M B)^T | B^T M B];  then b_c = -[t; B^T t]
        const double bc3 = -(o.y * o.t[2] - o.z * o.t[1]);
        const double bc4 = -(o.z * o.t[0] - o.x * o.t[2]);
        const double bc5 = -(o.x * o.t[1] - o.y * o.t[0]);
        row[0] = make_double2(o.M[0], o.M[1]);
        row[1] = make_double2(o.M[3], o.M[2]);
        row[2] = make_double2(o.M[4], o.M[5]);
        row[3] = make_double2(o.MB[0], o.MB[3]);
        row[4] = make_double2(o.MB[6], o.BMB[0]);
        row[5] = make_double2(o.MB[1], o.MB[4]);
        row[6] = make_double2(o.MB[7], o.BMB[1]);
        row[7] = make_double2(o.BMB[3], o.MB[2]);
        row[8] = make_double2(o.MB[5], o.MB[8]);
        row[9] = make_double2(o.BMB[2], o.BMB[4]);
        row[10] = make_double2(o.BMB[5], -o.t[0]);
        row[11] = make_double2(-o.t[1], -o.t[2]);
        row[12] = make_double2(bc3, bc4);
        // W = [M R; B^T M R] as nine (k, k+1) pairs, 512 bytes apart
        double2* Wp = reinterpret_cast<double2*>(a.W + w_pair_base(i));
        double w[18];
#pragma unroll
        for (int k = 0; k < 9; ++k) w[k] = o.MR[k];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          w[9 + j] = o.y * o.MR[6 + j] - o.z * o.MR[3 + j];
          w[12 + j] = o.z * o.MR[j] - o.x * o.MR[6 + j];
          w[15 + j] = o.x * o.MR[3 + j] - o.y * o.MR[j];
        }
#pragma unroll
        for (int k = 0; k < 9; ++k) Wp[kWTile * k] = make_double2(w[2 * k], w[2 * k + 1]);
        // V_p = R^T (M R) (xx xy xz yy yz zz), b_p = -R^T t
        row[13] = make_double2(bc5, P[0] * o.MR[0] + P[3] * o.MR[3] + P[6] * o.MR[6]);
      } else {
        row[13] = make_double2(0.0, P[0] * o.MR[0] + P[3] * o.MR[3] + P[6] * o.MR[6]);
      }
      row[14] = make_double2(P[0] * o.MR[1] + P[3] * o.MR[4] + P[6] * o.MR[7], P[0] * o.MR[2] + P[3] * o.MR[5] + P[6] * o.MR[8]);
      row[15] = make_double2(P[1] * o.MR[1] + P[4] * o.MR[4] + P[7] * o.MR[7], P[1] * o.MR[2] + P[4] * o.MR[5] + P[7] * o.MR[8]);
      row[16] = make_double2(P[2] * o.MR[2] + P[5] * o.MR[5] + P[8] * o.MR[8], -(P[0] * o.t[0] + P[3] * o.t[1] + P[6] * o.t[2]));
      row[17] = make_double2(-(P[1] * o.t[0] + P[4] * o.t[1] + P[7] * o.t[2]), -(P[2] * o.t[0] + P[5] * o.t[1] + P[8] * o.t[2]));
    }
    __syncthreads();

    // ---- phase 2
    if (lane < 27) {
      // camera side: a warp owns a slot, lane = value index; the slot's rows are contiguous
      for (int s_ = warp; s_ < blk.n_slots; s_ += kBlkObs / 32) {
        int k = sSeg[s_];
        const int k1 = sSeg[s_ + 1];
        const double* p = sT + k * kRow + lane;
        double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
        for (; k + 4 <= k1; k += 4, p += 4 * kRow) {
          acc0 += p[0]; acc1 += p[kRow]; acc2 += p[2 * kRow]; acc3 += p[3 * kRow];
        }
        for (; k < k1; ++k, p += kRow) acc0 += p[0];
        const double acc = (acc0 + acc1) + (acc2 + acc3);
        const int off = sOff[s_];
        if (lane < 21) red_add(a.S + (size_t)off * (a.ldS + 1) + tri_off, acc);
        else red_add(a.rhs + off + (lane - 21), acc);
      }
      // landmark side: a warp owns three landmarks at a time, lane = 9 * (landmark in group) + value
      const double* q = sT + 27 + lm_v;
      for (int l0 = 3 * warp; l0 < blk.n_lms; l0 += 3 * (kBlkObs / 32)) {
        const int l = l0 + lm_sub;
        if (l < blk.n_lms) {
          int k = sLmStart[l];
          const int k1 = sLmStart[l + 1];
          double acc0 = 0.0, acc1 = 0.0;
          for (; k + 2 <= k1; k += 2) {
            acc0 += q[sLmObs[k] * kRow];
            acc1 += q[sLmObs[k + 1] * kRow];
          }
          if (k < k1) acc0 += q[sLmObs[k] * kRow];
          a.Vg[9 * (size_t)(blk.lm_begin + l) + lm_v] = acc0 + acc1;
        }
      }
    }
    if (!has_next) break;
    __syncthreads();                                   // everybody is done with sT and the side arrays
    b = bn; blk = nblk; ++it;
    in = nin;
  }
  cp_async_wait_all();
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
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
          for (int c = 0; c < 3; ++c)
            a.W[w_index(i, 3 * r + c)] = w[0] * L.JT[r] * L.Jp[c] + w[1] * L.JT[6 + r] * L.Jp[3 + c] +
                                          w[2] * L.JT[12 + r] * L.Jp[6 + c];
      }
    }
  }
  block_sum_to(cost, a.scalars + 0 /*COST_LIN*/, sred);
}

// Cost only: sum rho(r) over ALL reprojection blocks (Problem.eval_cost,
// pyslam/problem.py:110-128) -> scalars[slot].
__global__ void __launch_bounds__(256)
reproj_cost_kernel(const ReprojArgs a, int slot, int obs_begin) {
  __shared__ double sred[8];
  double cost = 0.0;
  for (int i = obs_begin + blockIdx.x * blockDim.x + threadIdx.x; i < a.n_obs; i += gridDim.x * blockDim.x) {
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

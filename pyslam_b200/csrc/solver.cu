// solver.cu -- host side of libbslam.so: the device-resident problem, the
// lowering (ordering / Schur partition / layout), the per-iteration kernel
// schedule and the C ABI of include/bslam.h.  No CPU fallback exists anywhere
// in this file: every arithmetic step of the iteration is a kernel launch.
#include <chrono>
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <string>
#include <map>
#include <vector>

#include "../../include/bslam.h"
#include "cholesky.cuh"
#include "common.cuh"
#include "covariance.cuh"
#include "dense_blocks.cuh"
#include "photometric.cuh"
#include "posegraph.cuh"
#include "reproj.cuh"
#include "retract.cuh"
#include "schur.cuh"
#include "panel.cuh"
#include "peer.cuh"
#include "motion_only.cuh"
#include "ransac.cuh"
#include "image.cuh"

namespace {

std::string g_create_error;

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  cudaError_t alloc(size_t count) {
    if (count == n && p) return cudaSuccess;
    release();
    if (count == 0) return cudaSuccess;
    cudaError_t e = cudaMalloc(&p, count * sizeof(T));
    if (e == cudaSuccess) n = count;
    return e;
  }
};

struct EdgeBatch {
  int group = 3;
  bool binary = false;
  bool orient = false;     // PoseToPoseOrientationResidual: Tobs holds 9 doubles (C_2_1_obs) per factor, stiffness 3x3
  int n = 0;
  int per_block = 0;
  bs::Loss loss{0, 0.0};
  std::vector<int> i1, i2;
  std::vector<double> Tobs, stiff;
  DevBuf<int> d_i1, d_i2;
  DevBuf<double> d_Tobs, d_stiff;
};

struct MotionBlock {       // ReprojectionMotionOnly(Batch)Residual: one pose, n fixed points
  int pose_idx = 0, n = 0;
  bs::ReprojGroup g{};
  std::vector<double> pts1, obs2;
  DevBuf<double> d_pts1, d_obs2;
};

struct PhotoBlock {
  int pose_idx = 0, n_px = 0, w = 0, h = 0;
  int rot_idx = -1, vec_idx = -1;       // (SO3, t) parameter form: indices into the SO3 and vector tables (pose_idx unused)
  double intr[5] = {}, intensity_covar = 0, depth_covar = 0;
  bs::Loss loss{0, 0.0};
  std::vector<double> uvd, im_ref, im_jac, im_track;
  DevBuf<double> d_uvd, d_im_ref, d_im_jac, d_im_track;
};

}  // namespace

struct bslam_solver {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  bool finalized = false;
  int64_t launches = 0;
  bool timing = false;
  cudaEvent_t ev[14] = {};
  double timings[BSLAM_N_TIMINGS] = {};
  int shard_rank = 0;

  // ---- parameter tables (user order) ----
  int n_se3 = 0, n_se2 = 0, n_pt = 0, n_vec = 0, n_vec_entries = 0, n_so3 = 0;
  std::vector<uint8_t> se3_const, se2_const, pt_const, vec_const, so3_const;
  std::vector<double> h_so3;
  std::vector<int> so3_off;
  DevBuf<double> d_so3, b_so3;
  DevBuf<int> d_so3_off;
  std::vector<int> vec_dims, vec_start;
  std::vector<double> h_se3, h_se2, h_pts, h_vec;   // staging until finalize

  // ---- blocks (host staging) ----
  std::vector<int> ob_pose, ob_pt, ob_grp;
  std::vector<double> ob_uvd;
  std::vector<bs::ReprojGroup> groups;
  std::vector<EdgeBatch*> edges;
  std::vector<PhotoBlock*> photos;
  std::vector<MotionBlock*> motions;
  // dense (host-evaluated) blocks
  int dn_blocks = 0;
  std::vector<int> dn_rows, dn_pptr, dn_pkind, dn_pindex;
  std::vector<int> dn_row_ptr, dn_col_ptr, dn_col_index;
  std::vector<long long> dn_j_ptr;
  bool dn_uploaded = false;
  double dn_cost = 0.0;

  // ---- layout ----
  int n_lm = 0, n_red = 0, n_pad = 0, nblk = 0, dim = 0, n_obs = 0;
  DevBuf<unsigned char> d_used;                      // [n_pad] 1: real unknown, 0: padding entry of the reduced system
  std::vector<int> pt_perm, pt_iperm;               // user -> internal, internal -> user
  std::vector<int> se3_off, se2_off, vec_off, pt_off_user, vec_entry_off, pt_red_entry_off;

  // ---- device ----
  DevBuf<double> d_se3, d_se2, d_pts, d_vec, d_stage;
  DevBuf<double> b_se3, b_se2, b_pts, b_vec;        // snapshot
  DevBuf<int> d_se3_off, d_se2_off, d_vec_entry_off, d_ptred_entry_off, d_pt_perm;
  DevBuf<double> d_ou, d_ov, d_od;
  DevBuf<int> d_opose, d_opt, d_ogrp, d_lm_start;
  // landmark blocks of the fast reprojection / Schur kernels
  int n_lmblocks = 0, tail_begin = 0, n_regular = 0;
  size_t schur_smem = 0;
  int schur_grid = 0;
  int schur_max_lms = 1, schur_max_pairs = 1, schur_max_runs = 1;
  DevBuf<bs::SchurDesc> d_sch_descs;
  DevBuf<bs::SchurPair> d_sch_pairs;
  DevBuf<unsigned> d_sch_combos;
  DevBuf<unsigned> d_obs_code;                       // slot | block-local landmark << 8 | group << 16
  DevBuf<int> d_slot_off, d_lm_obs;                  // d_lm_obs: CSR position (landmark order) -> observation index
  int loss_kind = -1;                                // loss kind of the single reprojection group, -1: several groups
  DevBuf<double> d_slot_poses, d_slot_dx;
  int max_slots = 1;
  int n_slot_entries = 0, stage_len = 0;
  DevBuf<bs::LmBlock> d_blocks;
  DevBuf<int> d_slot_pose, d_tile_pose_ptr, d_tile_pose;     // tile of the reduced system -> SE3 poses inside it (fused retraction)
  DevBuf<unsigned char> d_lm_obs_local, d_seg_start;
  DevBuf<bs::ReprojGroup> d_groups;
  DevBuf<double> d_W, d_Vg, d_Vinv, d_red, d_dx, d_Linv;
  // dense landmark panels (panel.cuh): fused linearise + eliminate, W never materialised.  Landmarks
  // [0, n_fused) belong to panels; the landmark blocks [0, nb_fused) cover the same landmarks for the
  // step-wise / inspection API (bslam_linearize, bslam_reduce, bslam_get_normal_equations, bslam_covariance),
  // which keeps the materialised-W kernels.  bslam_iterate uses the panels and blocks [nb_fused, n_lmblocks).
  int fused_mode = 1;                                // 0: off, 1: well-filled panels only, 2: every panel that fits
  int n_panels = 0, n_fused = 0, nb_fused = 0, panel_max_var = 1, panel_grid = 1, finish_grid = 1;
  DevBuf<bs::PanelDesc> d_pdescs;
  DevBuf<double> d_pobs, d_se3_prev;
  DevBuf<unsigned short> d_pgrp;
  DevBuf<int> d_dn_row_ptr, d_dn_col_ptr, d_dn_col_index;
  DevBuf<long long> d_dn_j_ptr;
  DevBuf<double> d_dn_J, d_dn_e;
  double* h_scalars = nullptr;   // pinned

  // ---- CUDA graph of one whole iteration (single GPU, built-in blocks only) ----
  cudaGraphExec_t graph_exec = nullptr;
  // the two halves of a SHARDED iteration (multi-GPU: the host all-reduces the packed payload in between)
  cudaGraphExec_t graph_pre = nullptr, graph_post = nullptr;
  double graph_pre_lambda = -1.0;
  int graph_post_eval = -1;
  int64_t graph_pre_launches = 0, graph_post_launches = 0;
  double graph_lambda = -1.0;
  int graph_eval = -1;
  int64_t graph_launches = 0;
  bool use_graph = true;

  // ---- tile structure of the reduced system and the Cholesky task plan ----
  std::vector<uint8_t> tile_mask;                   // [(nblk+1) * nblk], lower triangle + rhs row
  bool plan_valid = false;
  int chol_epoch = 0, chol_grid = 0, n_tile_tasks = 0;
  DevBuf<bs::CholTask> d_tasks;
  DevBuf<double> d_cll, d_xll, d_yll;     // (value, epoch) records of the factorisation's critical hops (CholPlan::xll / yll)
  DevBuf<int> d_klist, d_bwd_ptr, d_bwd_rows, d_ready, d_xready, d_ticket;
  DevBuf<unsigned char> d_fill_mask;               // tile mask after symbolic fill-in
  DevBuf<long long> d_trace;                       // debug: per-task timestamps (bslam_debug_chol_trace)
  std::vector<bs::CholTask> h_tasks;
  std::vector<int> h_opose_lm, h_lm_start;         // observations in landmark order (host copy): local tile contributions of a shard
  DevBuf<int> d_dirty_tiles;                       // tiles (i*nt+j) that carry data: zeroed before every linearisation
  int n_dirty_tiles = 0;
  DevBuf<int> d_nz_tiles;                          // structurally non-zero tiles BEFORE fill-in (what assembly/Schur write)
  int n_nz_tiles = 0;
  DevBuf<double> d_pack;                           // exchange region: [n_nz_tiles * kNB^2 | rhs n_pad | scalars] = the multi-GPU
                                                   // payload, then the mailbox and flags of peer.cuh (kXchgTail doubles)
  size_t pack_len = 0;                             // doubles before the mailbox
  // declared couplings without residuals (bslam_add_coupling): rank-independent ordering of sharded problems
  std::vector<int> cp_group, cp_i1, cp_i2;
  // peer exchange (peer.cuh): world > 1 after bslam_peer_connect
  int world = 1;
  double* peer_region[bs::kMaxPeers] = {};
  bool peer_opened[bs::kMaxPeers] = {};
  double* mc_region = nullptr;                     // NVLS multicast mapping of the regions (bslam_peer_connect_symmetric)
  DevBuf<long long> d_peer_ctl;

  double* S() { return d_red.p; }
  double* rhs() { return d_red.p + (size_t)n_pad * n_pad; }
  double* scalars() { return d_red.p + (size_t)n_pad * n_pad + n_pad; }
  size_t red_len() const { return (size_t)n_pad * n_pad + n_pad + BSLAM_N_SCALARS; }

  ~bslam_solver() {
    for (auto* e : edges) delete e;
    for (auto* e : photos) delete e;
    for (auto* e : motions) delete e;
    for (auto& e : ev)
      if (e) cudaEventDestroy(e);
    if (graph_exec) cudaGraphExecDestroy(graph_exec);
    if (graph_pre) cudaGraphExecDestroy(graph_pre);
    if (graph_post) cudaGraphExecDestroy(graph_post);
    if (h_scalars) cudaFreeHost(h_scalars);
    for (int r = 0; r < bs::kMaxPeers; ++r)
      if (peer_opened[r]) cudaIpcCloseMemHandle(peer_region[r]);
    if (stream) cudaStreamDestroy(stream);
  }
};

namespace {

int fail(bslam_solver* s, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (s) s->err = buf;
  else g_create_error = buf;
  return code;
}

#define CU(call)                                                                              \
  do {                                                                                        \
    cudaError_t e__ = (call);                                                                 \
    if (e__ != cudaSuccess)                                                                   \
      return fail(s, BSLAM_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__),   \
                  __FILE__, __LINE__);                                                        \
  } while (0)

#define NEED(cond, ...)                                        \
  do {                                                         \
    if (!(cond)) return fail(s, BSLAM_E_INVALID, __VA_ARGS__); \
  } while (0)

#define LAUNCH(s, kernel, grid, block, smem, ...)                 \
  do {                                                            \
    kernel<<<grid, block, smem, (s)->stream>>>(__VA_ARGS__);      \
    (s)->launches++;                                              \
  } while (0)

inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

int build_chol_plan(bslam_solver* s);

template <typename T>
cudaError_t upload(DevBuf<T>& d, const std::vector<T>& h, cudaStream_t st) {
  cudaError_t e = d.alloc(h.size());
  if (e != cudaSuccess || h.empty()) return e;
  return cudaMemcpyAsync(d.p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, st);
}

bool valid_loss(int kind, double k) {
  if (kind < 0 || kind > BSLAM_LOSS_TDIST) return false;
  if (kind >= BSLAM_LOSS_CAUCHY && !(k > 0.0)) return false;
  return true;
}

// scatter/gather rows between user order and internal (permuted) point order
__global__ void permute_rows_kernel(int n, int width, const double* __restrict__ src, double* __restrict__ dst,
                                    const int* __restrict__ perm, int scatter) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * width) return;
  const int r = e / width, c = e - r * width;
  if (scatter) dst[(size_t)perm[r] * width + c] = src[e];
  else dst[e] = src[(size_t)perm[r] * width + c];
}

void drop_graph(bslam_solver* s) {
  if (s->graph_exec) cudaGraphExecDestroy(s->graph_exec);
  if (s->graph_pre) cudaGraphExecDestroy(s->graph_pre);
  if (s->graph_post) cudaGraphExecDestroy(s->graph_post);
  s->graph_pre = s->graph_post = nullptr;
  s->graph_exec = nullptr;
}

void record(bslam_solver* s, int i) {
  if (s->timing) cudaEventRecord(s->ev[i], s->stream);
}

bs::ReprojArgs reproj_args(bslam_solver* s) {
  bs::ReprojArgs a;
  a.n_obs = s->n_obs;
  a.n_lm = s->n_lm;
  a.obs_u = s->d_ou.p; a.obs_v = s->d_ov.p; a.obs_d = s->d_od.p;
  a.obs_pose = s->d_opose.p; a.obs_pt = s->d_opt.p;
  a.obs_grp = s->groups.size() > 1 ? s->d_ogrp.p : nullptr;
  a.groups = s->d_groups.p;
  if (!s->groups.empty()) a.g0 = s->groups[0];
  a.obs_code = s->d_obs_code.p;
  a.poses = s->d_se3.p;
  a.pose_off = s->d_se3_off.p;
  a.pts = s->d_pts.p;
  a.lm_start = s->d_lm_start.p; a.lm_obs = s->d_lm_obs.p;
  a.n_blocks = s->n_lmblocks;
  a.blocks = s->d_blocks.p;
  a.lm_obs_local = s->d_lm_obs_local.p; a.seg_start = s->d_seg_start.p;
  a.slot_off = s->d_slot_off.p; a.slot_poses = s->d_slot_poses.p; a.stage_len = s->stage_len;
  a.tail_begin = s->tail_begin;
  a.W = s->d_W.p; a.Vg = s->d_Vg.p;
  a.S = s->S(); a.ldS = s->n_pad; a.rhs = s->rhs(); a.scalars = s->scalars();
  return a;
}

bs::EdgeArgs edge_args(bslam_solver* s, EdgeBatch* b) {
  bs::EdgeArgs a;
  a.n = b->n;
  a.i1 = b->d_i1.p;
  a.i2 = b->binary ? b->d_i2.p : nullptr;
  a.Tobs = b->d_Tobs.p;
  a.stiff = b->d_stiff.p;
  a.stiff_per_block = b->per_block;
  a.loss = b->loss;
  a.poses = b->group == 3 ? s->d_se3.p : s->d_se2.p;
  a.pose_off = b->group == 3 ? s->d_se3_off.p : s->d_se2_off.p;
  a.S = s->S(); a.ldS = s->n_pad; a.rhs = s->rhs(); a.scalars = s->scalars();
  return a;
}

template <bool kCostOnly>
void launch_edges(bslam_solver* s, EdgeBatch* b, int slot) {
  if (b->n == 0) return;
  const bs::EdgeArgs a = edge_args(s, b);
  const int grid = cdiv(b->n, 128);
  if (b->orient) {
    LAUNCH(s, (bs::orientation_edge_kernel<kCostOnly>), grid, 128, 0, a, slot);
    return;
  }
  if (b->group == 3) {
    if (b->binary) LAUNCH(s, (bs::edge_kernel<3, true, kCostOnly>), grid, 128, 0, a, slot);
    else LAUNCH(s, (bs::edge_kernel<3, false, kCostOnly>), grid, 128, 0, a, slot);
  } else {
    if (b->binary) LAUNCH(s, (bs::edge_kernel<2, true, kCostOnly>), grid, 128, 0, a, slot);
    else LAUNCH(s, (bs::edge_kernel<2, false, kCostOnly>), grid, 128, 0, a, slot);
  }
}

bs::PhotoArgs photo_args(bslam_solver* s, PhotoBlock* b) {
  bs::PhotoArgs a;
  a.n_px = b->n_px;
  a.uvd = b->d_uvd.p; a.im_ref = b->d_im_ref.p; a.im_jac = b->d_im_jac.p; a.im_track = b->d_im_track.p;
  a.w = b->w; a.h = b->h;
  a.cu = b->intr[0]; a.cv = b->intr[1]; a.fu = b->intr[2]; a.fv = b->intr[3]; a.b = b->intr[4];
  a.intensity_covar = b->intensity_covar; a.depth_covar = b->depth_covar;
  a.loss = b->loss;
  if (b->rot_idx >= 0) {          // (SO3, t) form: photometric_residual.py:83-84,147-157
    a.R = s->d_so3.p + 9 * (size_t)b->rot_idx;
    a.t = s->d_vec.p + s->vec_start[b->vec_idx];
    a.off_rot = s->so3_off[b->rot_idx];
    a.off_trans = s->vec_off[b->vec_idx];
  } else {
    a.R = s->d_se3.p + 12 * (size_t)b->pose_idx;
    a.t = a.R + 9;
    a.off_trans = s->se3_off[b->pose_idx];
    a.off_rot = a.off_trans < 0 ? -1 : a.off_trans + 3;
  }
  a.S = s->S(); a.ldS = s->n_pad; a.rhs = s->rhs(); a.scalars = s->scalars();
  return a;
}

template <bool kCostOnly>
void launch_photos(bslam_solver* s, int slot) {
  for (auto* b : s->photos) {
    if (b->n_px == 0) continue;
    if (!kCostOnly) {                                            // all parameters constant: dropped (problem.py:343-348)
      const bool all_const = b->rot_idx >= 0 ? (s->so3_off[b->rot_idx] < 0 && s->vec_off[b->vec_idx] < 0) : s->se3_off[b->pose_idx] < 0;
      if (all_const) continue;
    }
    const int grid = std::min(cdiv(b->n_px, bs::kPhotoThreads), 148 * 2);
    LAUNCH(s, bs::photometric_kernel<kCostOnly>, grid, bs::kPhotoThreads, 0, photo_args(s, b), slot);
  }
}

template <bool kCostOnly>
void launch_motions(bslam_solver* s, int slot) {
  for (auto* b : s->motions) {
    if (b->n == 0) continue;
    if (!kCostOnly && s->se3_off[b->pose_idx] < 0) continue;    // the only parameter is constant: dropped (problem.py:343-348)
    bs::MotionArgs a;
    a.n = b->n; a.pts1 = b->d_pts1.p; a.obs2 = b->d_obs2.p; a.g = b->g;
    a.pose = s->d_se3.p + 12 * (size_t)b->pose_idx;
    a.pose_off = s->se3_off[b->pose_idx];
    a.S = s->S(); a.ldS = s->n_pad; a.rhs = s->rhs(); a.scalars = s->scalars();
    const int grid = std::max(1, std::min(cdiv(b->n, bs::kMotionThreads), 148));
    LAUNCH(s, bs::motion_only_kernel<kCostOnly>, grid, bs::kMotionThreads, 0, a, slot);
  }
}

// sum rho over all built-in blocks at the current parameters -> scalars[slot]
void launch_cost(bslam_solver* s, int slot) {
  if (s->n_obs > 0) {
    const int grid = std::min(cdiv(s->n_obs, 256), 148 * 8);
    LAUNCH(s, bs::reproj_cost_kernel, grid, 256, 0, reproj_args(s), slot, 0);
  }
  if (s->shard_rank == 0) {      // not sharded: counted once across GPUs
    for (auto* b : s->edges) launch_edges<true>(s, b, slot);
    launch_photos<true>(s, slot);
    launch_motions<true>(s, slot);
  }
}

// W (144 bytes per observation) exists only for the materialised-W kernels: observations outside the panels
// and the step-wise / inspection API.  Allocated on first use.
int ensure_W(bslam_solver* s) {
  if (s->d_W.p || s->n_obs == 0) return BSLAM_OK;
  CU(s->d_W.alloc(bs::w_alloc_len(s->n_obs)));
  CU(cudaMemsetAsync(s->d_W.p, 0, s->d_W.n * sizeof(double), s->stream));
  return BSLAM_OK;
}

bs::PanelArgs panel_args(bslam_solver* s, double lambda) {
  bs::PanelArgs a{};
  a.n_panels = s->n_panels;
  a.descs = s->d_pdescs.p; a.pobs = s->d_pobs.p; a.pgrp = s->d_pgrp.p;
  a.groups = s->d_groups.p;
  if (!s->groups.empty()) a.g0 = s->groups[0];
  a.poses = s->d_se3.p; a.poses_new = s->d_se3.p; a.pts_in = s->d_pts.p; a.pts = s->d_pts.p;
  a.lambda = lambda;
  a.Vg = s->d_Vg.p; a.Vinv = s->d_Vinv.p;
  a.S = s->S(); a.ldS = s->n_pad; a.rhs = s->rhs(); a.scalars = s->scalars();
  a.dx_red = s->d_dx.p; a.dx_lm = s->d_dx.p + s->n_pad;
  a.eval_cost = 0;
  a.max_var = s->panel_max_var;
  return a;
}

// `panels`: the landmarks of the dense panels are linearised AND eliminated by fused_panel_kernel in
// do_reduce; only the remaining landmark blocks are linearised here.
int do_linearize(bslam_solver* s, bool panels) {
  NEED(s->finalized, "bslam_finalize has not been called");
  NEED(s->dn_blocks == 0 || s->dn_uploaded, "dense blocks declared but bslam_upload_dense_values not called");
  record(s, 0);
  if (!s->plan_valid) { int rc = build_chol_plan(s); if (rc) return rc; }
  const int b0 = panels ? s->nb_fused : 0;                 // first landmark block handled by the block kernels
  const int nb = s->n_lmblocks - b0;
  if (nb > 0 || s->n_obs > s->tail_begin) { int rc = ensure_W(s); if (rc) return rc; }
  // zero only what the iteration dirties: the structurally non-zero tiles (after fill-in), rhs, scalars;
  // the rest of S was zeroed at finalize and is never written
  {
    bs::PrepareArgs a;
    a.S = s->S(); a.ld = s->n_pad; a.nt = s->nblk; a.n_tiles = s->n_dirty_tiles; a.tiles = s->d_dirty_tiles.p;
    a.used = s->d_used.p;
    a.rhs = s->rhs(); a.n_rhs = s->n_pad + BSLAM_N_SCALARS;
    a.vg_tail = s->d_Vg.p + 9 * (size_t)s->n_regular; a.n_vg_tail = 9 * (s->n_lm - s->n_regular);
    a.n_slot_entries = nb > 0 ? s->n_slot_entries : 0;
    a.slot_pose = s->d_slot_pose.p; a.poses = s->d_se3.p; a.slot_poses = s->d_slot_poses.p;
    a.poses_prev = s->d_se3_prev.p; a.n_prev = panels && s->n_panels > 0 ? 12 * s->n_se3 : 0;
    const int work = std::max(std::max(std::max(a.n_rhs, a.n_vg_tail), 12 * a.n_slot_entries), a.n_prev);
    LAUNCH(s, bs::prepare_kernel, s->n_dirty_tiles + std::max(1, cdiv(work, 256)), 256, 0, a);   // one element per thread: all gathers in flight at once
  }
  record(s, 1);
  if (nb > 0) {
    const int grid = std::min(nb, 148 * bs::kReprojCtas);      // persistent CTAs, all resident
    const size_t smem = 2 * (size_t)s->stage_len * sizeof(double);
    bs::ReprojArgs ra = reproj_args(s);
    ra.blocks += b0; ra.n_blocks = nb;
    switch (s->loss_kind) {
      case 0: LAUNCH(s, bs::reproj_block_kernel<0>, grid, bs::kBlkObs, smem, ra); break;
      case 1: LAUNCH(s, bs::reproj_block_kernel<1>, grid, bs::kBlkObs, smem, ra); break;
      case 2: LAUNCH(s, bs::reproj_block_kernel<2>, grid, bs::kBlkObs, smem, ra); break;
      case 3: LAUNCH(s, bs::reproj_block_kernel<3>, grid, bs::kBlkObs, smem, ra); break;
      case 4: LAUNCH(s, bs::reproj_block_kernel<4>, grid, bs::kBlkObs, smem, ra); break;
      case 5: LAUNCH(s, bs::reproj_block_kernel<5>, grid, bs::kBlkObs, smem, ra); break;
      default: LAUNCH(s, bs::reproj_block_kernel<-1>, grid, bs::kBlkObs, smem, ra); break;
    }
  }
  record(s, 2);
  if (s->n_obs > s->tail_begin)
    LAUNCH(s, bs::reproj_generic_kernel, cdiv(s->n_obs - s->tail_begin, 128), 128, 0, reproj_args(s));
  // blocks that are not sharded over GPUs (pose priors, relative-pose factors, photometric and
  // host-evaluated blocks) are assembled on shard 0 only: the all-reduce must count them once
  if (s->shard_rank == 0) {
    for (auto* b : s->edges) launch_edges<false>(s, b, BSLAM_S_COST_LIN);
    launch_photos<false>(s, BSLAM_S_COST_LIN);
    launch_motions<false>(s, BSLAM_S_COST_LIN);
    if (s->dn_blocks > 0) {
      bs::DenseArgs a;
      a.n_blocks = s->dn_blocks;
      a.row_ptr = s->d_dn_row_ptr.p; a.col_ptr = s->d_dn_col_ptr.p; a.j_ptr = s->d_dn_j_ptr.p;
      a.col_index = s->d_dn_col_index.p; a.J = s->d_dn_J.p; a.e = s->d_dn_e.p;
      a.S = s->S(); a.ldS = s->n_pad; a.rhs = s->rhs();
      int max_rows = 1;
      for (int r : s->dn_rows) max_rows = std::max(max_rows, r);
      dim3 grid(s->dn_blocks, cdiv(max_rows, bs::kDenseRowChunk));
      LAUNCH(s, bs::dense_blocks_kernel, grid, 256, 0, a);
      // the host computed sum rho(r) of its own blocks
      LAUNCH(s, bs::add_scalar_kernel, 1, 1, 0, s->scalars() + BSLAM_S_COST_LIN, s->dn_cost);
    }
  }
  s->dn_uploaded = false;
  record(s, 3);
  CU(cudaGetLastError());
  return BSLAM_OK;
}

int do_reduce(bslam_solver* s, double lambda, bool panels) {
  if (lambda > 0.0 && s->n_red > 0)
    LAUNCH(s, bs::damp_diag_kernel, cdiv(s->n_red, 128), 128, 0, s->S(), s->n_pad, s->n_red, lambda);
  record(s, 10);
  if (panels && s->n_panels > 0) {
    // after the damping of everything assembled so far: the kernel damps its own U_c diagonals
    const bs::PanelArgs pa = panel_args(s, lambda);
    const size_t smem = bs::panel_smem_bytes(s->panel_max_var);
    const int grid = std::min(s->n_panels, s->panel_grid);
    switch (s->loss_kind) {
      case 0: LAUNCH(s, bs::fused_panel_kernel<0>, grid, bs::kPanelThreads, smem, pa); break;
      case 1: LAUNCH(s, bs::fused_panel_kernel<1>, grid, bs::kPanelThreads, smem, pa); break;
      case 2: LAUNCH(s, bs::fused_panel_kernel<2>, grid, bs::kPanelThreads, smem, pa); break;
      case 3: LAUNCH(s, bs::fused_panel_kernel<3>, grid, bs::kPanelThreads, smem, pa); break;
      case 4: LAUNCH(s, bs::fused_panel_kernel<4>, grid, bs::kPanelThreads, smem, pa); break;
      case 5: LAUNCH(s, bs::fused_panel_kernel<5>, grid, bs::kPanelThreads, smem, pa); break;
      default: LAUNCH(s, bs::fused_panel_kernel<-1>, grid, bs::kPanelThreads, smem, pa); break;
    }
  }
  record(s, 11);
  const int b0 = panels ? s->nb_fused : 0;
  const int nb = s->n_lmblocks - b0;
  if (s->n_lm > 0 && (nb > 0 || s->n_lm > s->n_regular)) {
    bs::SchurArgs a;
    a.n_obs = s->n_obs; a.n_lm = s->n_lm; a.obs_begin = s->tail_begin; a.lambda = lambda;
    a.n_blocks = nb; a.blocks = s->d_blocks.p + b0; a.slot_pose = s->d_slot_pose.p; a.obs_code = s->d_obs_code.p;
    a.Vinv_out = s->d_Vinv.p;
    a.obs_pose = s->d_opose.p; a.obs_pt = s->d_opt.p; a.lm_start = s->d_lm_start.p; a.lm_obs = s->d_lm_obs.p;
    a.pose_off = s->d_se3_off.p;
    a.W = s->d_W.p; a.Vg = s->d_Vg.p; a.Vinv = s->d_Vinv.p;
    a.S = s->S(); a.ldS = s->n_pad; a.rhs = s->rhs();
    a.slot_off = s->d_slot_off.p;
    a.pairs = s->d_sch_pairs.p; a.combos = s->d_sch_combos.p;
    a.max_lms = s->schur_max_lms; a.max_pairs = s->schur_max_pairs; a.max_runs = s->schur_max_runs;
    a.descs = s->d_sch_descs.p + b0;
    if (nb > 0) {
      LAUNCH(s, bs::schur_block_kernel, std::min(nb, s->schur_grid), bs::kSchurThreads, s->schur_smem, a);
    }
    if (s->n_lm > s->n_regular) {
      LAUNCH(s, bs::landmark_invert_kernel, cdiv(s->n_lm - s->n_regular, 256), 256, 0, s->n_regular, s->n_lm, s->d_Vg.p,
             lambda, s->d_Vinv.p);
      const int lm_obs_end = s->n_obs;   // generic kernel skips observations of non-eliminated points itself
      if (lm_obs_end > s->tail_begin) LAUNCH(s, bs::schur_generic_kernel, cdiv(lm_obs_end - s->tail_begin, 128), 128, 0, a);
    }
  }
  record(s, 4);
  CU(cudaGetLastError());
  return BSLAM_OK;
}

// Tile-level structure of the reduced matrix from the co-visibility graph:
// tile (a,b) is non-zero if some residual block couples a parameter in tile a
// with one in tile b.
void mark_tiles(bslam_solver* s, const std::vector<int>& tiles) {
  const int nt = s->nblk;
  for (int a : tiles)
    for (int b : tiles)
      if (a >= b) s->tile_mask[(size_t)a * nt + b] = 1;
}

void tiles_of(int off, int dof, std::vector<int>& out) {
  if (off < 0) return;
  for (int t = off / bs::kNB; t <= (off + dof - 1) / bs::kNB; ++t)
    if (std::find(out.begin(), out.end(), t) == out.end()) out.push_back(t);
}

void build_tile_mask(bslam_solver* s, const std::vector<int>& opose, const std::vector<int>& lm_start) {
  const int nt = s->nblk;
  s->tile_mask.assign((size_t)(nt + 1) * nt, 0);
  for (int t = 0; t < nt; ++t) s->tile_mask[(size_t)t * nt + t] = 1;
  for (int t = 0; t < nt; ++t) s->tile_mask[(size_t)nt * nt + t] = 1;      // right-hand side row
  std::vector<int> tiles;
  for (int q = 0; q < s->n_lm; ++q) {                                      // Schur fill: poses sharing a landmark
    tiles.clear();
    for (int k = lm_start[q]; k < lm_start[q + 1]; ++k) tiles_of(s->se3_off[opose[k]], 6, tiles);
    mark_tiles(s, tiles);
  }
  for (auto* b : s->edges) {
    if (!b->binary) continue;
    const std::vector<int>& off = b->group == 3 ? s->se3_off : s->se2_off;
    const int dof = b->group == 3 ? 6 : 3;
    for (int e = 0; e < b->n; ++e) {
      tiles.clear();
      tiles_of(off[b->i1[e]], dof, tiles);
      tiles_of(off[b->i2[e]], dof, tiles);
      mark_tiles(s, tiles);
    }
  }
  for (int b = 0; b < s->dn_blocks; ++b) {
    tiles.clear();
    for (int c = s->dn_col_ptr[b]; c < s->dn_col_ptr[b + 1]; ++c) tiles_of(s->dn_col_index[c], 1, tiles);
    mark_tiles(s, tiles);
  }
  for (auto* b : s->photos) {                                                // (SO3, t) photometric blocks couple their two parameters
    if (b->rot_idx < 0) continue;
    tiles.clear();
    tiles_of(s->so3_off[b->rot_idx], 3, tiles);
    tiles_of(s->vec_off[b->vec_idx], 3, tiles);
    mark_tiles(s, tiles);
  }
  for (size_t e = 0; e < s->cp_i1.size(); ++e) {                             // declared couplings (bslam_add_coupling)
    const std::vector<int>& off = s->cp_group[e] == 3 ? s->se3_off : s->se2_off;
    const int dof = s->cp_group[e] == 3 ? 6 : 3;
    tiles.clear();
    tiles_of(off[s->cp_i1[e]], dof, tiles);
    tiles_of(off[s->cp_i2[e]], dof, tiles);
    mark_tiles(s, tiles);
  }
  s->plan_valid = false;
}

// Symbolic factorisation on the tile grid + task list of chol_solve_kernel.
int build_chol_plan(bslam_solver* s) {
  const int nt = s->nblk;
  std::vector<uint8_t> m = s->tile_mask;
  auto at = [&](int i, int j) -> uint8_t& { return m[(size_t)i * nt + j]; };
  std::vector<int> rows;
  for (int k = 0; k < nt; ++k) {
    rows.clear();
    for (int i = k + 1; i <= nt; ++i)
      if (at(i, k)) rows.push_back(i);
    for (int a : rows)
      for (int b : rows)
        if (a >= b && b < nt) at(a, b) = 1;
  }
  std::vector<bs::CholTask> tasks;
  std::vector<int> klist, bwd_ptr(nt + 1, 0), bwd_rows;
  // position of every structurally non-zero tile (before fill-in) in the packed multi-GPU payload
  std::vector<int> nz, nz_slot((size_t)nt * nt, -1);
  for (int i = 0; i < nt; ++i)
    for (int j = 0; j <= i; ++j)
      if (s->tile_mask[(size_t)i * nt + j]) { nz_slot[(size_t)i * nt + j] = (int)nz.size(); nz.push_back(i * nt + j); }
  for (int j = 0; j < nt; ++j)
    for (int i = j; i <= nt; ++i) {
      if (!at(i, j)) continue;
      bs::CholTask t{};
      t.slot = i == nt ? -2 : nz_slot[(size_t)i * nt + j];
      t.ranks = 0xff;
      t.i = i; t.j = j; t.kbeg = (int)klist.size();
      for (int k = 0; k < j; ++k)
        if (at(i, k) && at(j, k)) klist.push_back(k);
      t.kend = (int)klist.size();
      tasks.push_back(t);
    }
  // Ticket order = dependency LEVEL (longest chain of producer tiles below the task), not column order:
  // with nested dissection the leaves of all subtrees sit at scattered column indices, and a CTA that
  // holds an early ticket for a task high in the tree would only spin while ready leaves wait for a CTA.
  // Any topological order keeps the no-deadlock argument (a waiting CTA waits on earlier tickets only).
  {
    std::vector<int> tix((size_t)(nt + 1) * nt, -1), level(tasks.size(), 0), order(tasks.size());
    for (size_t t = 0; t < tasks.size(); ++t) tix[(size_t)tasks[t].i * nt + tasks[t].j] = (int)t;
    for (size_t t = 0; t < tasks.size(); ++t) {          // column-major: producers come first
      const bs::CholTask& T = tasks[t];
      int lv = 0;
      for (int kk = T.kbeg; kk < T.kend; ++kk) {
        const int k = klist[kk];
        lv = std::max(lv, level[tix[(size_t)T.i * nt + k]] + 1);
        lv = std::max(lv, level[tix[(size_t)T.j * nt + k]] + 1);
      }
      if (T.i != T.j) lv = std::max(lv, level[tix[(size_t)T.j * nt + T.j]] + 1);
      level[t] = lv;
    }
    // inside a task, consume the producer tiles in the order they become available (lowest level first), so
    // that only one accumulation step is left when the last producer arrives
    for (size_t t = 0; t < tasks.size(); ++t) {
      const bs::CholTask& T = tasks[t];
      auto key = [&](int k) { return std::max(level[tix[(size_t)T.i * nt + k]], level[tix[(size_t)T.j * nt + k]]); };
      std::stable_sort(klist.begin() + T.kbeg, klist.begin() + T.kend, [&](int x, int y) { return key(x) < key(y); });
    }
    // early C tiles: the diagonal task of row i takes its last (critical) producer column k straight from
    // C_ik and X_kk; the off-diagonal task (i, k) publishes C_ik for it (see CholPlan::cscr)
    for (size_t t = 0; t < tasks.size(); ++t) {
      bs::CholTask& T = tasks[t];
      if (T.i != T.j) continue;
      T.early = std::min(bs::kEarly, T.kend - T.kbeg);
      for (int e = 0; e < T.early; ++e)
        tasks[tix[(size_t)T.i * nt + klist[T.kend - T.early + e]]].early = e + 1;
    }
    // Ticket levels.  A diagonal task takes its early producers (i, k) as C tiles, which those tasks publish BEFORE they
    // wait for diag(k): for the ticket order the diagonal task sits at the level of these tasks, right behind them
    // (same level, same row, diagonal last), instead of two levels up behind every other task of the levels between --
    // near the leaves, where a level holds more tasks than there are CTAs, it would otherwise get its CTA microseconds late.
    std::vector<int> tlevel(tasks.size(), 0);
    for (size_t t = 0; t < tasks.size(); ++t) {          // column-major: producers come first
      const bs::CholTask& T = tasks[t];
      const int n_early = T.i == T.j ? T.early : 0;
      int lv = 0;
      for (int kk = T.kbeg; kk < T.kend; ++kk) {
        const int k = klist[kk];
        const int add = kk >= T.kend - n_early ? 0 : 1;
        lv = std::max(lv, tlevel[tix[(size_t)T.i * nt + k]] + add);
        if (T.i != T.j) lv = std::max(lv, tlevel[tix[(size_t)T.j * nt + k]] + 1);
      }
      if (T.i != T.j) lv = std::max(lv, tlevel[tix[(size_t)T.j * nt + T.j]] + 1);
      tlevel[t] = lv;
    }
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) {
      if (tlevel[x] != tlevel[y]) return tlevel[x] < tlevel[y];
      if (tasks[x].i != tasks[y].i) return tasks[x].i < tasks[y].i;
      return (tasks[x].i == tasks[x].j) < (tasks[y].i == tasks[y].j);
    });
    std::vector<bs::CholTask> sorted(tasks.size());
    for (size_t t = 0; t < tasks.size(); ++t) sorted[t] = tasks[order[t]];
    tasks.swap(sorted);
  }
  for (int k = 0; k < nt; ++k) {
    for (int i = nt - 1; i > k; --i)
      if (at(i, k)) bwd_rows.push_back(i);
    bwd_ptr[k + 1] = (int)bwd_rows.size();
  }
  if (klist.empty()) klist.push_back(0);
  if (bwd_rows.empty()) bwd_rows.push_back(0);
  s->n_tile_tasks = (int)tasks.size();
  s->h_tasks = tasks;
  cudaStream_t st = s->stream;
  {
    std::vector<int> dirty;
    for (int i = 0; i < nt; ++i)
      for (int j = 0; j <= i; ++j)
        if (at(i, j)) dirty.push_back(i * nt + j);
    s->n_dirty_tiles = (int)dirty.size();
    if (dirty.empty()) dirty.push_back(0);
    CU(upload(s->d_dirty_tiles, dirty, st));
    s->n_nz_tiles = (int)nz.size();
    if (nz.empty()) nz.push_back(0);
    CU(upload(s->d_nz_tiles, nz, st));
    if (s->world > 1) return fail(s, BSLAM_E_INVALID, "the tile structure changed after bslam_peer_connect");
    s->pack_len = (size_t)s->n_nz_tiles * bs::kNB * bs::kNB + s->n_pad + BSLAM_N_SCALARS;
    CU(s->d_pack.alloc(s->pack_len + bs::kXchgTail));
    CU(cudaMemsetAsync(s->d_pack.p, 0, s->d_pack.n * sizeof(double), st));
  }
  CU(upload(s->d_fill_mask, m, st));
  CU(upload(s->d_tasks, tasks, st));
  CU(upload(s->d_klist, klist, st));
  CU(upload(s->d_bwd_ptr, bwd_ptr, st));
  CU(upload(s->d_bwd_rows, bwd_rows, st));
  CU(s->d_ready.alloc((size_t)(nt + 1) * nt));
  CU(s->d_xready.alloc(nt));
  CU(s->d_cll.alloc((size_t)2 * bs::kEarly * nt * bs::kNB * bs::kNB));
  CU(cudaMemsetAsync(s->d_cll.p, 0, s->d_cll.n * sizeof(double), st));
  CU(s->d_xll.alloc((size_t)2 * nt * bs::kNB * bs::kNB));
  CU(s->d_yll.alloc((size_t)2 * nt * bs::kNB));
  CU(cudaMemsetAsync(s->d_xll.p, 0, s->d_xll.n * sizeof(double), st));
  CU(cudaMemsetAsync(s->d_yll.p, 0, s->d_yll.n * sizeof(double), st));
  CU(s->d_ticket.alloc(4));                     // [ticket, epoch of the last completed launch, CTAs done]
  CU(cudaMemsetAsync(s->d_ticket.p, 0, 4 * sizeof(int), st));
  CU(cudaMemsetAsync(s->d_ready.p, 0, s->d_ready.n * sizeof(int), st));
  CU(cudaMemsetAsync(s->d_xready.p, 0, s->d_xready.n * sizeof(int), st));
  s->chol_epoch = 0;
  int per_sm = 0, sms = 0;
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bs::chol_solve_kernel, bs::kCholThreads, bs::kCholSmem));
  CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device));
  if (per_sm < 1) return fail(s, BSLAM_E_CUDA, "chol_solve_kernel does not fit on an SM");
  s->chol_grid = std::min(per_sm * sms, s->n_tile_tasks + nt);
  s->plan_valid = true;
  return BSLAM_OK;
}

// iteration path of pure panel problems: the Cholesky kernel's backward tasks retract the SE3 poses themselves
bool fuse_retract(bslam_solver* s, bool panels) {
  return panels && s->n_panels > 0 && s->n_lmblocks == s->nb_fused && s->n_se3 > 0 && s->d_trace.p == nullptr;
}

int do_solve_reduced(bslam_solver* s, bool retract_poses = false) {
  if (!s->plan_valid) {
    int rc = build_chol_plan(s);
    if (rc) return rc;
  }
  bs::CholPlan p;
  p.nt = s->nblk;
  p.n_tile_tasks = s->n_tile_tasks;
  p.tasks = s->d_tasks.p; p.klist = s->d_klist.p; p.bwd_ptr = s->d_bwd_ptr.p; p.bwd_rows = s->d_bwd_rows.p;
  p.ready = s->d_ready.p; p.xready = s->d_xready.p; p.ticket = s->d_ticket.p;
  p.cll = s->d_cll.p; p.xll = s->d_xll.p; p.yll = s->d_yll.p;
  p.trace = s->d_trace.p;
  p.rt_poses = retract_poses ? s->d_se3.p : nullptr;
  p.rt_pose_off = s->d_se3_off.p; p.rt_tile_ptr = s->d_tile_pose_ptr.p; p.rt_tile_pose = s->d_tile_pose.p;
  p.rt_norm = s->shard_rank == 0 ? 1 : 0;
  p.world = s->world;
  p.rhs_off = s->n_nz_tiles * bs::kNB * bs::kNB;
  for (int r = 0; r < bs::kCholMaxPeers; ++r) p.peer_pack[r] = r < s->world ? s->peer_region[r] : nullptr;
  p.mc_pack = s->world > 1 ? s->mc_region : nullptr;
  // no per-launch memsets: the flags carry the launch epoch, which the kernel advances itself
  LAUNCH(s, bs::chol_solve_kernel, s->chol_grid, bs::kCholThreads, bs::kCholSmem, s->S(), s->n_pad, s->d_Linv.p,
         s->d_dx.p, s->scalars(), p);
  record(s, 5);
  record(s, 6);
  if (s->n_lm > s->n_regular) {      // regular landmarks are back-substituted by lm_finish_kernel (do_retract)
    bs::BacksubArgs a;
    a.n_lm = s->n_lm; a.q_begin = s->n_regular; a.n_obs = s->n_obs; a.lm_off = s->n_pad;
    a.obs_pose = s->d_opose.p; a.lm_start = s->d_lm_start.p; a.lm_obs = s->d_lm_obs.p; a.pose_off = s->d_se3_off.p;
    a.W = s->d_W.p; a.Vg = s->d_Vg.p; a.Vinv = s->d_Vinv.p; a.dx = s->d_dx.p;
    LAUNCH(s, bs::backsub_kernel, cdiv(s->n_lm - s->n_regular, 128), 128, 0, a);
  }
  record(s, 7);
  CU(cudaGetLastError());
  return BSLAM_OK;
}

// poses_done: the SE3 table was already retracted (and ||dx_c||^2 accumulated) by the Cholesky kernel (fuse_retract)
int do_retract(bslam_solver* s, int eval_new_cost, bool panels, bool poses_done = false) {
  const double* dx = s->d_dx.p;
  // ||dx||^2: the reduced part is replicated across shards, count it on shard 0 only
  const int n_red_here = s->shard_rank == 0 ? s->n_red : 0;
  const int b0 = panels ? s->nb_fused : 0;
  const int nb = s->n_lmblocks - b0;
  const bool use_panels = panels && s->n_panels > 0;
  if (s->n_se3 > 0 && !poses_done) {
    const int n_slots = nb > 0 ? s->n_slot_entries : 0;
    LAUNCH(s, bs::retract_se3_slots_kernel, cdiv(s->n_se3 + n_slots, 128), 128, 0, s->n_se3, s->d_se3.p, s->d_se3_off.p, dx, n_slots,
           s->d_slot_off.p, s->d_slot_poses.p, s->d_slot_dx.p, n_red_here, s->scalars() + BSLAM_S_DX_NORM2, nullptr);
  }
  if (s->n_se2 > 0) LAUNCH(s, bs::retract_poses_kernel<2>, cdiv(s->n_se2, 128), 128, 0, s->n_se2, s->d_se2.p, s->d_se2_off.p, dx);
  if (s->n_so3 > 0) LAUNCH(s, bs::retract_so3_kernel, cdiv(s->n_so3, 128), 128, 0, s->n_so3, s->d_so3.p, s->d_so3_off.p, dx);
  if (s->n_vec_entries > 0)
    LAUNCH(s, bs::retract_flat_kernel, cdiv(s->n_vec_entries, 256), 256, 0, s->n_vec_entries, s->d_vec.p, s->d_vec_entry_off.p, dx);
  if (s->n_pt > s->n_lm) {
    const int n3 = 3 * (s->n_pt - s->n_lm);
    LAUNCH(s, bs::retract_flat_kernel, cdiv(n3, 256), 256, 0, n3, s->d_pts.p + 3 * (size_t)s->n_lm, s->d_ptred_entry_off.p, dx);
  }
  if (s->n_se3 == 0 && n_red_here > 0)
    LAUNCH(s, bs::sumsq_kernel, std::min(cdiv(s->n_red, 256), 148), 256, 0, s->n_red, dx, s->scalars() + BSLAM_S_DX_NORM2);
  record(s, 8);
  // panels: back-substitution (W^T dx_c recomputed) + retraction + ||dx_p||^2 + cost at the new point, fused
  if (use_panels) {
    bs::PanelArgs pa = panel_args(s, 0.0);
    pa.poses = s->d_se3_prev.p; pa.poses_new = s->d_se3.p; pa.eval_cost = eval_new_cost;
    const size_t smem = 0;
    const int grid = std::min(cdiv((long long)s->n_panels * bs::kUnitsPerPanel, bs::kFinishThreads / 32), s->finish_grid);
    switch (s->loss_kind) {
      case 0: LAUNCH(s, bs::panel_finish_kernel<0>, grid, bs::kFinishThreads, smem, pa); break;
      case 1: LAUNCH(s, bs::panel_finish_kernel<1>, grid, bs::kFinishThreads, smem, pa); break;
      case 2: LAUNCH(s, bs::panel_finish_kernel<2>, grid, bs::kFinishThreads, smem, pa); break;
      case 3: LAUNCH(s, bs::panel_finish_kernel<3>, grid, bs::kFinishThreads, smem, pa); break;
      case 4: LAUNCH(s, bs::panel_finish_kernel<4>, grid, bs::kFinishThreads, smem, pa); break;
      case 5: LAUNCH(s, bs::panel_finish_kernel<5>, grid, bs::kFinishThreads, smem, pa); break;
      default: LAUNCH(s, bs::panel_finish_kernel<-1>, grid, bs::kFinishThreads, smem, pa); break;
    }
  }
  // landmark blocks: the same from the materialised W
  if (nb > 0) {
    bs::FinishArgs a;
    a.n_obs = s->n_obs; a.lm_off = s->n_pad; a.eval_cost = eval_new_cost; a.n_blocks = nb;
    a.max_slots = s->max_slots;
    a.blocks = s->d_blocks.p + b0; a.obs_code = s->d_obs_code.p;
    a.slot_poses = s->d_slot_poses.p; a.slot_dx = s->d_slot_dx.p;
    a.lm_obs_local = s->d_lm_obs_local.p;
    a.obs_pose = s->d_opose.p; a.groups = s->d_groups.p;
    if (!s->groups.empty()) a.g0 = s->groups[0];
    a.lm_start = s->d_lm_start.p;
    a.obs_u = s->d_ou.p; a.obs_v = s->d_ov.p; a.obs_d = s->d_od.p;
    a.poses = s->d_se3.p; a.pts = s->d_pts.p; a.W = s->d_W.p; a.Vg = s->d_Vg.p; a.Vinv = s->d_Vinv.p;
    a.dx = s->d_dx.p; a.scalars = s->scalars();
    const int fgrid = std::min(nb, 148 * bs::kFinishCtas);      // persistent CTAs, all resident
    const size_t fsmem = bs::finish_smem_bytes(s->max_slots);
    switch (s->loss_kind) {
      case 0: LAUNCH(s, bs::lm_finish_kernel<0>, fgrid, bs::kBlkObs, fsmem, a); break;
      case 1: LAUNCH(s, bs::lm_finish_kernel<1>, fgrid, bs::kBlkObs, fsmem, a); break;
      case 2: LAUNCH(s, bs::lm_finish_kernel<2>, fgrid, bs::kBlkObs, fsmem, a); break;
      case 3: LAUNCH(s, bs::lm_finish_kernel<3>, fgrid, bs::kBlkObs, fsmem, a); break;
      case 4: LAUNCH(s, bs::lm_finish_kernel<4>, fgrid, bs::kBlkObs, fsmem, a); break;
      case 5: LAUNCH(s, bs::lm_finish_kernel<5>, fgrid, bs::kBlkObs, fsmem, a); break;
      default: LAUNCH(s, bs::lm_finish_kernel<-1>, fgrid, bs::kBlkObs, fsmem, a); break;
    }
  }
  // tail: landmarks with long tracks (already back-substituted), then everything that is not a landmark block
  if (s->n_lm > s->n_regular) {
    const int n3 = 3 * (s->n_lm - s->n_regular);
    LAUNCH(s, bs::retract_landmarks_kernel, cdiv(n3, 256), 256, 0, n3, s->d_pts.p + 3 * (size_t)s->n_regular,
           dx + s->n_pad + 3 * (size_t)s->n_regular);
    LAUNCH(s, bs::sumsq_kernel, std::min(cdiv(n3, 256), 148 * 4), 256, 0, n3, dx + s->n_pad + 3 * (size_t)s->n_regular,
           s->scalars() + BSLAM_S_DX_NORM2);
  }
  if (eval_new_cost) {
    if (s->n_obs > s->tail_begin) {
      const int n = s->n_obs - s->tail_begin;
      LAUNCH(s, bs::reproj_cost_kernel, std::min(cdiv(n, 256), 148 * 8), 256, 0, reproj_args(s), BSLAM_S_COST_NEW, s->tail_begin);
    }
    if (s->shard_rank == 0) {
      for (auto* b : s->edges) launch_edges<true>(s, b, BSLAM_S_COST_NEW);
      launch_photos<true>(s, BSLAM_S_COST_NEW);
      launch_motions<true>(s, BSLAM_S_COST_NEW);
    }
  }
  record(s, 9);
  CU(cudaGetLastError());
  return BSLAM_OK;
}

// CUDA loads kernels lazily, and a first-use load may synchronise the context.  Handles of ONE process that
// rendezvous on the device (one spinning while another is still enqueueing) must therefore have every kernel of the
// iteration resident before the first sharded iteration (CUDA programming guide, "lazy loading", concurrent execution).
template <int kLoss>
void preload_loss_kernels() {
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, (const void*)bs::reproj_block_kernel<kLoss>);
  cudaFuncGetAttributes(&fa, (const void*)bs::fused_panel_kernel<kLoss>);
  cudaFuncGetAttributes(&fa, (const void*)bs::panel_finish_kernel<kLoss>);
  cudaFuncGetAttributes(&fa, (const void*)bs::lm_finish_kernel<kLoss>);
}
void preload_iteration_kernels(bslam_solver* s) {
  cudaFuncAttributes fa;
  const void* fns[] = {
      (const void*)bs::prepare_kernel, (const void*)bs::reproj_generic_kernel, (const void*)bs::reproj_cost_kernel,
      (const void*)bs::edge_kernel<3, true, false>, (const void*)bs::edge_kernel<3, false, false>,
      (const void*)bs::edge_kernel<2, true, false>, (const void*)bs::edge_kernel<2, false, false>,
      (const void*)bs::edge_kernel<3, true, true>, (const void*)bs::edge_kernel<3, false, true>,
      (const void*)bs::edge_kernel<2, true, true>, (const void*)bs::edge_kernel<2, false, true>,
      (const void*)bs::photometric_kernel<false>, (const void*)bs::photometric_kernel<true>,
      (const void*)bs::dense_blocks_kernel, (const void*)bs::add_scalar_kernel, (const void*)bs::damp_diag_kernel,
      (const void*)bs::schur_block_kernel, (const void*)bs::landmark_invert_kernel, (const void*)bs::schur_generic_kernel,
      (const void*)bs::chol_solve_kernel, (const void*)bs::backsub_kernel, (const void*)bs::retract_se3_slots_kernel,
      (const void*)bs::retract_poses_kernel<2>, (const void*)bs::retract_flat_kernel, (const void*)bs::sumsq_kernel,
      (const void*)bs::retract_landmarks_kernel, (const void*)bs::pack_tiles_kernel,
      (const void*)bs::peer_pack_signal_kernel, (const void*)bs::peer_scalar_exchange_kernel, (const void*)permute_rows_kernel};
  for (const void* f : fns) cudaFuncGetAttributes(&fa, f);
  switch (s->loss_kind) {
    case 0: preload_loss_kernels<0>(); break;
    case 1: preload_loss_kernels<1>(); break;
    case 2: preload_loss_kernels<2>(); break;
    case 3: preload_loss_kernels<3>(); break;
    case 4: preload_loss_kernels<4>(); break;
    case 5: preload_loss_kernels<5>(); break;
    default: preload_loss_kernels<-1>(); break;
  }
  cudaGetLastError();
}

bs::PeerCtx peer_ctx(bslam_solver* s) {
  bs::PeerCtx pc{};
  pc.world = s->world; pc.rank = s->shard_rank;
  for (int r = 0; r < bs::kMaxPeers; ++r) pc.region[r] = s->peer_region[r];
  pc.pack_len = s->pack_len;
  pc.ctl = s->d_peer_ctl.p;
  return pc;
}

// sharded iteration, after the rank's linearise + eliminate kernels: publish the partial reduced system to the peers
int do_peer_publish(bslam_solver* s) {
  LAUNCH(s, bs::peer_pack_signal_kernel, s->n_nz_tiles + 1, 256, 0, s->S(), s->n_pad, s->nblk, s->d_nz_tiles.p, s->n_nz_tiles,
         s->rhs(), s->n_pad + BSLAM_N_SCALARS, s->scalars(), peer_ctx(s));
  record(s, 12);
  CU(cudaGetLastError());
  return BSLAM_OK;
}

// sharded iteration, end: sum the partial scalars over the ranks (and order the next iteration after the peers' reads)
int do_peer_scalars(bslam_solver* s) {
  LAUNCH(s, bs::peer_scalar_exchange_kernel, 1, 32, 0, s->scalars(), peer_ctx(s), 1);
  record(s, 13);
  CU(cudaGetLastError());
  return BSLAM_OK;
}

int sync_and_timings(bslam_solver* s);

// everything bslam_iterate* allocates lazily, before a graph capture starts
int prepare_iterate(bslam_solver* s) {
  int rc;
  if (!s->plan_valid && (rc = build_chol_plan(s))) return rc;
  if (s->n_lmblocks > s->nb_fused || s->n_obs > s->tail_begin) return ensure_W(s);
  return BSLAM_OK;
}

int fetch_scalars(bslam_solver* s) {
  CU(cudaMemcpyAsync(s->h_scalars, s->scalars(), BSLAM_N_SCALARS * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  return sync_and_timings(s);
}

// wait for the handle's stream; with timing enabled, turn the recorded events into per-phase milliseconds
int sync_and_timings(bslam_solver* s) {
  CU(cudaStreamSynchronize(s->stream));
  if (s->timing) {
    auto el = [&](int a, int b) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, s->ev[a], s->ev[b]);
      return (double)ms;
    };
    s->timings[BSLAM_T_LINEARIZE] = el(0, 3);
    s->timings[BSLAM_T_REPROJ] = el(1, 2);
    s->timings[BSLAM_T_SCHUR] = el(3, 4);
    s->timings[BSLAM_T_CHOLESKY] = el(4, 5);
    s->timings[BSLAM_T_TRSV] = el(5, 6);
    s->timings[BSLAM_T_BACKSUB] = el(6, 7);
    s->timings[BSLAM_T_RETRACT] = el(7, 8);
    s->timings[BSLAM_T_COST] = el(8, 9);
    s->timings[BSLAM_T_TOTAL] = el(0, 9);
    s->timings[BSLAM_T_FUSED] = el(10, 11);
    if (s->world > 1) {          // sharded iteration: publish + rendezvous, and the scalar exchange, on their own
      s->timings[BSLAM_T_PEER_PUBLISH] = el(4, 12);
      s->timings[BSLAM_T_CHOLESKY] = el(12, 5);
      s->timings[BSLAM_T_PEER_SCALARS] = el(9, 13);
    }
  }
  return BSLAM_OK;
}

}  // namespace

// Record `body` (kernel launches and async copies on the handle's stream) as a CUDA graph once, replay it after.
template <typename F>
int run_graphed(bslam_solver* s, cudaGraphExec_t* exec, int64_t* n_launches, F body) {
  if (!*exec) {
    cudaGraph_t graph = nullptr;
    const int64_t l0 = s->launches;
    CU(cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal));
    const int rc = body();
    const cudaError_t ee = cudaStreamEndCapture(s->stream, &graph);
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (ee != cudaSuccess || !graph) {
      if (graph) cudaGraphDestroy(graph);
      return fail(s, BSLAM_E_CUDA, "graph capture failed: %s", cudaGetErrorString(ee));
    }
    const cudaError_t ce = cudaGraphInstantiate(exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess) { *exec = nullptr; return fail(s, BSLAM_E_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(ce)); }
    *n_launches = s->launches - l0;
    s->launches = l0;
  }
  CU(cudaGraphLaunch(*exec, s->stream));
  s->launches += *n_launches;
  return BSLAM_OK;
}

// =============================================================== C ABI ====

extern "C" {

int bslam_version(void) { return 101; }

int bslam_tile_edge(void) { return bs::kNB; }

const char* bslam_last_error(const bslam_solver* s) { return s ? s->err.c_str() : g_create_error.c_str(); }

int bslam_create(bslam_solver** out, int device) {
  bslam_solver* s = nullptr;
  if (!out) return fail(s, BSLAM_E_INVALID, "out is NULL");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(s, BSLAM_E_CUDA, "no CUDA device available (%s); libbslam has no CPU fallback",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
  if (device < 0 || device >= count) return fail(s, BSLAM_E_INVALID, "device %d out of range [0,%d)", device, count);
  if ((e = cudaSetDevice(device)) != cudaSuccess) return fail(s, BSLAM_E_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
  bslam_solver* h = new bslam_solver();
  h->device = device;
  if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess ||
      (e = cudaMallocHost(&h->h_scalars, BSLAM_N_SCALARS * sizeof(double))) != cudaSuccess) {
    fail(s, BSLAM_E_CUDA, "stream/pinned allocation: %s", cudaGetErrorString(e));
    delete h;
    return BSLAM_E_CUDA;
  }
  for (auto& ev : h->ev) cudaEventCreate(&ev);
  { const char* e = getenv("BSLAM_NO_GRAPH"); if (e && atoi(e)) h->use_graph = false; }
  { const char* e = getenv("BSLAM_FUSED"); if (e) h->fused_mode = std::max(0, std::min(2, atoi(e))); }
  cudaFuncSetAttribute(bs::chol_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bs::kCholSmem);
  *out = h;
  return BSLAM_OK;
}

void bslam_destroy(bslam_solver* s) {
  if (!s) return;
  cudaSetDevice(s->device);
  if (s->stream) cudaStreamSynchronize(s->stream);
  delete s;
}

// ---------------------------------------------------------------- parameters

static int set_table(bslam_solver* s, const char* what, int n, int width, const double* vals, const uint8_t* is_const,
                     int& n_field, std::vector<uint8_t>& cflags, std::vector<double>& host, DevBuf<double>& dev) {
  NEED(s, "NULL solver");
  NEED(n >= 0 && (n == 0 || vals), "%s: bad arguments", what);
  CU(cudaSetDevice(s->device));
  if (s->finalized) {
    NEED(n == n_field, "%s: table size changed after finalize (%d -> %d); call bslam_clear_blocks first", what, n_field, n);
    if (is_const)
      for (int i = 0; i < n; ++i)
        NEED((is_const[i] != 0) == (cflags[i] != 0), "%s: constant flags changed after finalize", what);
    if (n > 0) CU(cudaMemcpyAsync(dev.p, vals, (size_t)n * width * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    return BSLAM_OK;
  }
  n_field = n;
  cflags.assign(n, 0);
  if (is_const)
    for (int i = 0; i < n; ++i) cflags[i] = is_const[i] ? 1 : 0;
  host.assign(vals, vals + (size_t)n * width);
  return BSLAM_OK;
}

int bslam_set_poses_se3(bslam_solver* s, int n, const double* Rt, const uint8_t* is_const) {
  NEED(s, "NULL solver");
  return set_table(s, "bslam_set_poses_se3", n, 12, Rt, is_const, s->n_se3, s->se3_const, s->h_se3, s->d_se3);
}

int bslam_set_poses_se2(bslam_solver* s, int n, const double* Rt, const uint8_t* is_const) {
  NEED(s, "NULL solver");
  return set_table(s, "bslam_set_poses_se2", n, 6, Rt, is_const, s->n_se2, s->se2_const, s->h_se2, s->d_se2);
}

int bslam_set_rotations_so3(bslam_solver* s, int n, const double* R, const uint8_t* is_const) {
  NEED(s, "NULL solver");
  return set_table(s, "bslam_set_rotations_so3", n, 9, R, is_const, s->n_so3, s->so3_const, s->h_so3, s->d_so3);
}

int bslam_set_points(bslam_solver* s, int n, const double* xyz, const uint8_t* is_const) {
  NEED(s, "NULL solver");
  if (!s->finalized) return set_table(s, "bslam_set_points", n, 3, xyz, is_const, s->n_pt, s->pt_const, s->h_pts, s->d_pts);
  NEED(n == s->n_pt, "bslam_set_points: table size changed after finalize");
  if (is_const)
    for (int i = 0; i < n; ++i)
      NEED((is_const[i] != 0) == (s->pt_const[i] != 0), "bslam_set_points: constant flags changed after finalize");
  if (n == 0) return BSLAM_OK;
  CU(cudaSetDevice(s->device));
  // user order -> internal (landmarks-first) order, permuted on the device
  CU(cudaMemcpyAsync(s->d_stage.p, xyz, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  LAUNCH(s, permute_rows_kernel, cdiv(3LL * n, 256), 256, 0, n, 3, s->d_stage.p, s->d_pts.p, s->d_pt_perm.p, 1);
  CU(cudaGetLastError());
  return BSLAM_OK;
}

int bslam_set_vectors(bslam_solver* s, int n, const int32_t* dims, const double* values, const uint8_t* is_const) {
  NEED(s, "NULL solver");
  NEED(n >= 0 && (n == 0 || (dims && values)), "bslam_set_vectors: bad arguments");
  long long total = 0;
  for (int i = 0; i < n; ++i) {
    NEED(dims[i] >= 0, "bslam_set_vectors: negative dimension");
    total += dims[i];
  }
  if (s->finalized) {
    NEED(n == s->n_vec, "bslam_set_vectors: table size changed after finalize");
    for (int i = 0; i < n; ++i) NEED(dims[i] == s->vec_dims[i], "bslam_set_vectors: dimensions changed after finalize");
    CU(cudaSetDevice(s->device));
    if (total > 0) CU(cudaMemcpyAsync(s->d_vec.p, values, total * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    return BSLAM_OK;
  }
  s->n_vec = n;
  s->vec_dims.assign(dims, dims + n);
  s->vec_start.assign(n + 1, 0);
  for (int i = 0; i < n; ++i) s->vec_start[i + 1] = s->vec_start[i] + dims[i];
  s->n_vec_entries = (int)total;
  s->vec_const.assign(n, 0);
  if (is_const)
    for (int i = 0; i < n; ++i) s->vec_const[i] = is_const[i] ? 1 : 0;
  s->h_vec.assign(values, values + total);
  return BSLAM_OK;
}

static int get_table(bslam_solver* s, const char* what, int n, int width, double* out, DevBuf<double>& dev) {
  NEED(s && s->finalized, "%s: solver not finalized", what);
  if (n == 0) return BSLAM_OK;
  NEED(out, "%s: NULL output", what);
  CU(cudaSetDevice(s->device));
  CU(cudaMemcpyAsync(out, dev.p, (size_t)n * width * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  return BSLAM_OK;
}

int bslam_get_poses_se3(bslam_solver* s, double* Rt) {
  NEED(s, "NULL solver");
  return get_table(s, "bslam_get_poses_se3", s->n_se3, 12, Rt, s->d_se3);
}
int bslam_get_poses_se2(bslam_solver* s, double* Rt) {
  NEED(s, "NULL solver");
  return get_table(s, "bslam_get_poses_se2", s->n_se2, 6, Rt, s->d_se2);
}
int bslam_get_rotations_so3(bslam_solver* s, double* R) {
  NEED(s, "NULL solver");
  return get_table(s, "bslam_get_rotations_so3", s->n_so3, 9, R, s->d_so3);
}
int bslam_get_points(bslam_solver* s, double* xyz) {
  NEED(s && s->finalized, "bslam_get_points: solver not finalized");
  if (s->n_pt == 0) return BSLAM_OK;
  NEED(xyz, "bslam_get_points: NULL output");
  CU(cudaSetDevice(s->device));
  LAUNCH(s, permute_rows_kernel, cdiv(3LL * s->n_pt, 256), 256, 0, s->n_pt, 3, s->d_pts.p, s->d_stage.p, s->d_pt_perm.p, 0);
  CU(cudaMemcpyAsync(xyz, s->d_stage.p, (size_t)s->n_pt * 3 * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  return BSLAM_OK;
}
int bslam_get_vectors(bslam_solver* s, double* values) {
  NEED(s, "NULL solver");
  return get_table(s, "bslam_get_vectors", s->n_vec_entries, 1, values, s->d_vec);
}

// -------------------------------------------------------------------- blocks

int bslam_add_reprojection_blocks(bslam_solver* s, int n, const int32_t* pose_idx, const int32_t* pt_idx,
                                  const double* obs, const double* stiffness, int per_block, const double intr[5],
                                  int loss_kind, double loss_k) {
  NEED(s, "NULL solver");
  NEED(!s->finalized, "bslam_add_reprojection_blocks after finalize; call bslam_clear_blocks first");
  NEED(n >= 0 && (n == 0 || (pose_idx && pt_idx && obs && stiffness && intr)), "bslam_add_reprojection_blocks: bad arguments");
  NEED(valid_loss(loss_kind, loss_k), "bslam_add_reprojection_blocks: invalid loss (kind %d, k %g)", loss_kind, loss_k);
  if (n == 0) return BSLAM_OK;
  for (int i = 0; i < n; ++i) {
    NEED(pose_idx[i] >= 0 && pose_idx[i] < s->n_se3, "reprojection block %d: pose index %d outside the SE3 table (%d)", i,
         pose_idx[i], s->n_se3);
    NEED(pt_idx[i] >= 0 && pt_idx[i] < s->n_pt, "reprojection block %d: point index %d outside the point table (%d)", i,
         pt_idx[i], s->n_pt);
  }
  auto make_group = [&](const double* S9) {
    bs::ReprojGroup g{};
    g.cu = intr[0]; g.cv = intr[1]; g.fu = intr[2]; g.fv = intr[3]; g.b = intr[4];
    for (int k = 0; k < 9; ++k) g.S[k] = S9[k];
    g.loss.kind = loss_kind;
    g.loss.k = loss_k;
    return g;
  };
  auto find_group = [&](const bs::ReprojGroup& g) {
    // newest groups first: consecutive blocks nearly always share their constants
    auto same = [](const bs::ReprojGroup& a, const bs::ReprojGroup& b) {
      if (a.cu != b.cu || a.cv != b.cv || a.fu != b.fu || a.fv != b.fv || a.b != b.b) return false;
      if (a.loss.kind != b.loss.kind || a.loss.k != b.loss.k) return false;
      for (int k = 0; k < 9; ++k) if (a.S[k] != b.S[k]) return false;
      return true;
    };
    for (int k = (int)s->groups.size() - 1; k >= 0 && k >= (int)s->groups.size() - 64; --k)
      if (same(s->groups[k], g)) return k;
    s->groups.push_back(g);
    return (int)s->groups.size() - 1;
  };
  int gshared = per_block ? -1 : find_group(make_group(stiffness));
  s->ob_pose.insert(s->ob_pose.end(), pose_idx, pose_idx + n);
  s->ob_pt.insert(s->ob_pt.end(), pt_idx, pt_idx + n);
  s->ob_uvd.insert(s->ob_uvd.end(), obs, obs + 3 * (size_t)n);
  for (int i = 0; i < n; ++i) s->ob_grp.push_back(per_block ? find_group(make_group(stiffness + 9 * (size_t)i)) : gshared);
  return BSLAM_OK;
}

static int add_edges(bslam_solver* s, const char* what, int group, bool binary, int n, const int32_t* i1, const int32_t* i2,
                     const double* Tobs, const double* stiffness, int per_block, int loss_kind, double loss_k) {
  NEED(s, "NULL solver");
  NEED(!s->finalized, "%s after finalize; call bslam_clear_blocks first", what);
  NEED(group == BSLAM_SE2 || group == BSLAM_SE3, "%s: group must be BSLAM_SE2 or BSLAM_SE3", what);
  NEED(n >= 0 && (n == 0 || (i1 && (!binary || i2) && Tobs && stiffness)), "%s: bad arguments", what);
  NEED(valid_loss(loss_kind, loss_k), "%s: invalid loss (kind %d, k %g)", what, loss_kind, loss_k);
  if (n == 0) return BSLAM_OK;
  const int table = group == 3 ? s->n_se3 : s->n_se2;
  for (int i = 0; i < n; ++i) {
    NEED(i1[i] >= 0 && i1[i] < table, "%s %d: pose index %d outside the table (%d)", what, i, i1[i], table);
    if (binary) NEED(i2[i] >= 0 && i2[i] < table, "%s %d: pose index %d outside the table (%d)", what, i, i2[i], table);
  }
  const int store = group == 3 ? 12 : 6, dof = group == 3 ? 6 : 3;
  EdgeBatch* b = new EdgeBatch();
  b->group = group; b->binary = binary; b->n = n; b->per_block = per_block ? 1 : 0;
  b->loss.kind = loss_kind; b->loss.k = loss_k;
  b->i1.assign(i1, i1 + n);
  if (binary) b->i2.assign(i2, i2 + n);
  b->Tobs.assign(Tobs, Tobs + (size_t)n * store);
  b->stiff.assign(stiffness, stiffness + (size_t)(per_block ? n : 1) * dof * dof);
  s->edges.push_back(b);
  return BSLAM_OK;
}

int bslam_add_pose_blocks(bslam_solver* s, int group, int n, const int32_t* pose_idx, const double* T_obs,
                          const double* stiffness, int per_block, int loss_kind, double loss_k) {
  return add_edges(s, "bslam_add_pose_blocks", group, false, n, pose_idx, nullptr, T_obs, stiffness, per_block, loss_kind, loss_k);
}

int bslam_add_pose_to_pose_blocks(bslam_solver* s, int group, int n, const int32_t* idx1, const int32_t* idx2,
                                  const double* T21_obs, const double* stiffness, int per_block, int loss_kind, double loss_k) {
  return add_edges(s, "bslam_add_pose_to_pose_blocks", group, true, n, idx1, idx2, T21_obs, stiffness, per_block, loss_kind, loss_k);
}

int bslam_add_photometric_block(bslam_solver* s, int pose_idx, int n_px, const double* uvd_ref, const double* im_ref,
                                const double* im_jac, const double* im_track, int width, int height, const double intr[5],
                                double intensity_stiffness, double depth_stiffness, int loss_kind, double loss_k) {
  NEED(s, "NULL solver");
  NEED(!s->finalized, "bslam_add_photometric_block after finalize; call bslam_clear_blocks first");
  NEED(n_px >= 0 && width > 0 && height > 0 && im_track && intr && (n_px == 0 || (uvd_ref && im_ref && im_jac)),
       "bslam_add_photometric_block: bad arguments");
  NEED(pose_idx >= 0 && pose_idx < s->n_se3, "photometric block: pose index %d outside the SE3 table (%d)", pose_idx, s->n_se3);
  NEED(intensity_stiffness > 0.0 && depth_stiffness > 0.0, "photometric block: stiffness must be positive");
  NEED(valid_loss(loss_kind, loss_k), "bslam_add_photometric_block: invalid loss (kind %d, k %g)", loss_kind, loss_k);
  PhotoBlock* b = new PhotoBlock();
  b->pose_idx = pose_idx; b->n_px = n_px; b->w = width; b->h = height;
  for (int k = 0; k < 5; ++k) b->intr[k] = intr[k];
  b->intensity_covar = 1.0 / (intensity_stiffness * intensity_stiffness);     // stiffness ** -2 (photometric_residual.py:59-60)
  b->depth_covar = 1.0 / (depth_stiffness * depth_stiffness);
  b->loss.kind = loss_kind; b->loss.k = loss_k;
  b->uvd.assign(uvd_ref, uvd_ref + 3 * (size_t)n_px);
  b->im_ref.assign(im_ref, im_ref + n_px);
  b->im_jac.assign(im_jac, im_jac + 2 * (size_t)n_px);
  b->im_track.assign(im_track, im_track + (size_t)width * height);
  s->photos.push_back(b);
  return BSLAM_OK;
}

int bslam_add_motion_only_blocks(bslam_solver* s, int pose_idx, int n, const double* pts_1, const double* obs_2,
                                 const double* stiffness, const double intr[5], int loss_kind, double loss_k) {
  NEED(s, "NULL solver");
  NEED(!s->finalized, "bslam_add_motion_only_blocks after finalize; call bslam_clear_blocks first");
  NEED(n >= 0 && stiffness && intr && (n == 0 || (pts_1 && obs_2)), "bslam_add_motion_only_blocks: bad arguments");
  NEED(pose_idx >= 0 && pose_idx < s->n_se3, "motion-only block: pose index %d outside the SE3 table (%d)", pose_idx, s->n_se3);
  NEED(valid_loss(loss_kind, loss_k), "bslam_add_motion_only_blocks: invalid loss (kind %d, k %g)", loss_kind, loss_k);
  MotionBlock* b = new MotionBlock();
  b->pose_idx = pose_idx; b->n = n;
  b->g.cu = intr[0]; b->g.cv = intr[1]; b->g.fu = intr[2]; b->g.fv = intr[3]; b->g.b = intr[4];
  for (int k = 0; k < 9; ++k) b->g.S[k] = stiffness[k];
  b->g.loss.kind = loss_kind; b->g.loss.k = loss_k;
  b->pts1.assign(pts_1, pts_1 + 3 * (size_t)n);
  b->obs2.assign(obs_2, obs_2 + 3 * (size_t)n);
  s->motions.push_back(b);
  return BSLAM_OK;
}

int bslam_add_orientation_blocks(bslam_solver* s, int n, const int32_t* idx1, const int32_t* idx2, const double* C21_obs,
                                 const double* stiffness, int per_block, int loss_kind, double loss_k) {
  NEED(s, "NULL solver");
  NEED(!s->finalized, "bslam_add_orientation_blocks after finalize; call bslam_clear_blocks first");
  NEED(n >= 0 && (n == 0 || (idx1 && idx2 && C21_obs && stiffness)), "bslam_add_orientation_blocks: bad arguments");
  NEED(valid_loss(loss_kind, loss_k), "bslam_add_orientation_blocks: invalid loss (kind %d, k %g)", loss_kind, loss_k);
  if (n == 0) return BSLAM_OK;
  for (int i = 0; i < n; ++i)
    NEED(idx1[i] >= 0 && idx1[i] < s->n_se3 && idx2[i] >= 0 && idx2[i] < s->n_se3, "orientation block %d: pose index outside the SE3 table (%d)", i,
         s->n_se3);
  EdgeBatch* b = new EdgeBatch();
  b->group = 3; b->binary = true; b->orient = true; b->n = n; b->per_block = per_block ? 1 : 0;
  b->loss.kind = loss_kind; b->loss.k = loss_k;
  b->i1.assign(idx1, idx1 + n);
  b->i2.assign(idx2, idx2 + n);
  b->Tobs.assign(C21_obs, C21_obs + 9 * (size_t)n);
  b->stiff.assign(stiffness, stiffness + (size_t)(per_block ? n : 1) * 9);
  s->edges.push_back(b);
  return BSLAM_OK;
}

int bslam_add_photometric_block_split(bslam_solver* s, int rot_idx, int vec_idx, int n_px, const double* uvd_ref, const double* im_ref,
                                      const double* im_jac, const double* im_track, int width, int height, const double intr[5],
                                      double intensity_stiffness, double depth_stiffness, int loss_kind, double loss_k) {
  NEED(s, "NULL solver");
  NEED(rot_idx >= 0 && rot_idx < s->n_so3, "photometric block: rotation index %d outside the SO3 table (%d)", rot_idx, s->n_so3);
  NEED(vec_idx >= 0 && vec_idx < s->n_vec && s->vec_dims[vec_idx] == 3, "photometric block: translation %d is not a 3-vector of the vector table",
       vec_idx);
  // same validation and storage as the SE3 form; the pose index is replaced by the (SO3, t) pair
  const int n_se3 = s->n_se3;
  s->n_se3 = std::max(1, n_se3);
  const int rc = bslam_add_photometric_block(s, 0, n_px, uvd_ref, im_ref, im_jac, im_track, width, height, intr, intensity_stiffness,
                                             depth_stiffness, loss_kind, loss_k);
  s->n_se3 = n_se3;
  if (rc) return rc;
  s->photos.back()->rot_idx = rot_idx;
  s->photos.back()->vec_idx = vec_idx;
  return BSLAM_OK;
}

int bslam_set_dense_blocks(bslam_solver* s, int n_blocks, const int32_t* rows, const int32_t* param_ptr,
                           const int32_t* param_kind, const int32_t* param_index) {
  NEED(s, "NULL solver");
  NEED(!s->finalized, "bslam_set_dense_blocks after finalize; call bslam_clear_blocks first");
  NEED(n_blocks >= 0 && (n_blocks == 0 || (rows && param_ptr && param_kind && param_index)), "bslam_set_dense_blocks: bad arguments");
  s->dn_blocks = n_blocks;
  s->dn_rows.assign(rows, rows + n_blocks);
  s->dn_pptr.assign(param_ptr, param_ptr + n_blocks + (n_blocks ? 1 : 0));
  const int np = n_blocks ? param_ptr[n_blocks] : 0;
  s->dn_pkind.assign(param_kind, param_kind + np);
  s->dn_pindex.assign(param_index, param_index + np);
  const int tables[4] = {s->n_se3, s->n_se2, s->n_pt, s->n_vec};
  for (int i = 0; i < np; ++i) {
    NEED(param_kind[i] >= 0 && param_kind[i] <= 3, "dense block parameter %d: kind %d invalid", i, param_kind[i]);
    NEED(param_index[i] >= 0 && param_index[i] < tables[param_kind[i]], "dense block parameter %d: index %d outside table", i, param_index[i]);
  }
  for (int b = 0; b < n_blocks; ++b) NEED(rows[b] > 0, "dense block %d has %d rows", b, rows[b]);
  return BSLAM_OK;
}

int bslam_upload_dense_values(bslam_solver* s, const double* e, size_t n_e, const double* J, size_t n_J, double cost) {
  NEED(s && s->finalized, "bslam_upload_dense_values: solver not finalized");
  NEED(s->dn_blocks > 0, "bslam_upload_dense_values: no dense blocks declared");
  NEED(n_e == (size_t)s->dn_row_ptr.back() && n_J == (size_t)s->dn_j_ptr.back(),
       "bslam_upload_dense_values: expected %d residual rows and %lld Jacobian entries, got %zu and %zu",
       s->dn_row_ptr.back(), s->dn_j_ptr.back(), n_e, n_J);
  CU(cudaSetDevice(s->device));
  CU(cudaMemcpyAsync(s->d_dn_e.p, e, n_e * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  CU(cudaMemcpyAsync(s->d_dn_J.p, J, n_J * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  CU(cudaStreamSynchronize(s->stream));   // the caller may reuse its buffers immediately
  s->dn_cost = cost;
  s->dn_uploaded = true;
  return BSLAM_OK;
}

int bslam_clear_blocks(bslam_solver* s) {
  NEED(s, "NULL solver");
  if (s->stream) cudaStreamSynchronize(s->stream);
  s->ob_pose.clear(); s->ob_pt.clear(); s->ob_grp.clear(); s->ob_uvd.clear(); s->groups.clear();
  for (auto* e : s->edges) delete e;
  s->edges.clear();
  for (auto* e : s->photos) delete e;
  s->photos.clear();
  for (auto* e : s->motions) delete e;
  s->motions.clear();
  s->dn_blocks = 0;
  s->dn_rows.clear(); s->dn_pptr.clear(); s->dn_pkind.clear(); s->dn_pindex.clear();
  s->cp_group.clear(); s->cp_i1.clear(); s->cp_i2.clear();
  for (int r = 0; r < bs::kMaxPeers; ++r) {
    if (s->peer_opened[r]) cudaIpcCloseMemHandle(s->peer_region[r]);
    s->peer_opened[r] = false;
    s->peer_region[r] = nullptr;
  }
  s->mc_region = nullptr;
  s->world = 1;
  drop_graph(s);
  s->finalized = false;
  return BSLAM_OK;
}

// ------------------------------------------------------------------ lowering

// BSLAM_FINALIZE_TIMING=1: wall time of the lowering stages on stderr (one-time cost per problem)
#define FIN_T(name)                                                                                       \
  do {                                                                                                    \
    if (fin_timing) {                                                                                     \
      const auto now = std::chrono::steady_clock::now();                                                  \
      fprintf(stderr, "finalize %-16s %8.1f ms\n", name, std::chrono::duration<double, std::milli>(now - fin_t0).count()); \
      fin_t0 = now;                                                                                       \
    }                                                                                                     \
  } while (0)

int bslam_finalize(bslam_solver* s) {
  const bool fin_timing = getenv("BSLAM_FINALIZE_TIMING") != nullptr;
  auto fin_t0 = std::chrono::steady_clock::now();
  NEED(s, "NULL solver");
  NEED(!s->finalized, "already finalized");
  CU(cudaSetDevice(s->device));
  const int N = (int)s->ob_pose.size();
  s->n_obs = N;

  FIN_T("start");
  // ---- which points are eliminated by the Schur complement ----
  std::vector<uint8_t> by_reproj(s->n_pt, 0), by_dense(s->n_pt, 0);
  for (int i = 0; i < N; ++i) by_reproj[s->ob_pt[i]] = 1;
  for (size_t i = 0; i < s->dn_pkind.size(); ++i)
    if (s->dn_pkind[i] == 2) by_dense[s->dn_pindex[i]] = 1;
  // Landmark order: regular landmarks (<= kMaxTrack observations) sorted by the first
  // pose that sees them (neighbouring landmarks then share cameras), then "big"
  // landmarks, then the points that are not eliminated.
  std::vector<int> n_obs_of(s->n_pt, 0), first_pose(s->n_pt, 1 << 30);
  for (int i = 0; i < N; ++i) {
    n_obs_of[s->ob_pt[i]]++;
    first_pose[s->ob_pt[i]] = std::min(first_pose[s->ob_pt[i]], s->ob_pose[i]);
  }
  // a landmark observed twice by the same pose goes through the generic (atomic) kernels: the block
  // kernels assume at most one observation per (pose, landmark)
  std::vector<uint8_t> dup_obs(s->n_pt, 0);
  std::vector<int> obs_ptr(s->n_pt + 1, 0), obs_ids(N);      // observations of every point (CSR, registration order)
  {
    for (int i = 0; i < N; ++i) obs_ptr[s->ob_pt[i] + 1]++;
    for (int p = 0; p < s->n_pt; ++p) obs_ptr[p + 1] += obs_ptr[p];
    std::vector<int> cur(obs_ptr.begin(), obs_ptr.end() - 1);
    for (int i = 0; i < N; ++i) obs_ids[cur[s->ob_pt[i]]++] = i;
    std::vector<int> seen_by(s->n_se3 + 1, -1);               // pose -> last point that stamped it
    for (int p = 0; p < s->n_pt; ++p)
      for (int k = obs_ptr[p]; k < obs_ptr[p + 1]; ++k) {
        int& stamp = seen_by[s->ob_pose[obs_ids[k]]];
        if (stamp == p) dup_obs[p] = 1;
        stamp = p;
      }
  }
  std::vector<int> regular, big, rest;
  for (int p = 0; p < s->n_pt; ++p) {
    bool elim = false;
    if (!s->pt_const[p]) {
      if (by_reproj[p] && by_dense[p])
        return fail(s, BSLAM_E_STRUCTURE,
                    "point %d is used by both reprojection blocks and host-evaluated blocks; this mix is not supported yet", p);
      if (!by_reproj[p] && !by_dense[p])
        return fail(s, BSLAM_E_STRUCTURE, "variable point %d is not referenced by any residual block (singular system)", p);
      elim = by_reproj[p] != 0;
    }
    if (!elim) rest.push_back(p);
    else if (n_obs_of[p] <= bs::kMaxTrack && !dup_obs[p]) regular.push_back(p);
    else big.push_back(p);
  }
  std::stable_sort(regular.begin(), regular.end(), [&](int a, int b) { return first_pose[a] < first_pose[b]; });

  FIN_T("classify");
  // ---- dense landmark panels (panel.cuh) ----------------------------------------------
  // Greedy runs of consecutive regular landmarks: up to kPanelLm landmarks whose observing poses number
  // at most kPanelRows.  A run becomes a panel when its (row, landmark) grid is well filled (mode 1) or
  // always (mode 2); the landmarks of rejected runs keep the landmark-block kernels.  Panel landmarks
  // come first in the internal order.
  struct HostPanel { int first, n_lms; std::vector<int> poses; int n_var; };
  std::vector<HostPanel> hpanels;
  if (s->fused_mode > 0 && s->groups.size() < 65536) {
    // Panels of at most kPanelLm = 64 landmarks.  (32-landmark panels were measured for small landmark shards --
    // tools/panel_cap_study.py, BSLAM_PANEL_CAP -- and never won once every panel keeps two CTAs per SM resident.)
    // Two passes: panels only where the grid is well filled; if only a few landmarks are left over, they are packed
    // into (sparser / shorter) panels as well, so that the iteration does not pay the whole landmark-block path
    // (three more kernels and a separate retraction) for a handful of landmarks.
    size_t cap = (size_t)bs::kPanelLm;
    { const char* e = getenv("BSLAM_PANEL_CAP"); if (e && (atoi(e) == 32 || atoi(e) == 64)) cap = (size_t)atoi(e); }   // study knob
    std::vector<int> fused_lms, irregular, poses, trial;
    for (int pass = 0; pass < 2; ++pass) {
    const bool accept_all = s->fused_mode >= 2 || pass == 1;
    fused_lms.clear(); irregular.clear(); hpanels.clear();
    size_t q = 0;
    while (q < regular.size()) {
      poses.clear();
      size_t q1 = q;
      long long n_cells = 0;
      while (q1 < regular.size() && q1 - q < cap) {
        trial = poses;
        const int p = regular[q1];
        for (int k = obs_ptr[p]; k < obs_ptr[p + 1]; ++k) {
          const int pose = s->ob_pose[obs_ids[k]];
          if (std::find(trial.begin(), trial.end(), pose) == trial.end()) trial.push_back(pose);
        }
        if ((int)trial.size() > bs::kPanelRows) break;
        // at most kPanelMaxVar VARIABLE rows: the kernel's shared memory grows with them, and the 8th row would cost the
        // second resident CTA per SM (half the throughput of the whole grid for a few wider panels)
        if (!poses.empty() && std::count_if(trial.begin(), trial.end(), [&](int q2) { return !s->se3_const[q2]; }) > bs::kPanelMaxVar) break;
        poses.swap(trial);
        n_cells += obs_ptr[p + 1] - obs_ptr[p];
        ++q1;
      }
      if (q1 == q) {                               // a single landmark seen by more poses than a panel holds
        irregular.push_back(regular[q]);
        ++q;
        continue;
      }
      const int n_lms = (int)(q1 - q);
      const bool dense = n_lms >= 16 && 2 * n_cells >= (long long)poses.size() * n_lms;
      if (accept_all || dense) {
        HostPanel hp;
        hp.first = (int)fused_lms.size(); hp.n_lms = n_lms;
        std::sort(poses.begin(), poses.end());
        for (int pose : poses) if (!s->se3_const[pose]) hp.poses.push_back(pose);
        hp.n_var = (int)hp.poses.size();
        for (int pose : poses) if (s->se3_const[pose]) hp.poses.push_back(pose);
        hpanels.push_back(std::move(hp));
        fused_lms.insert(fused_lms.end(), regular.begin() + q, regular.begin() + q1);
      } else {
        irregular.insert(irregular.end(), regular.begin() + q, regular.begin() + q1);
      }
      q = q1;
    }
    // second pass only when the first left a small remainder (<= 10 % of the landmarks) outside the panels
    if (accept_all || irregular.empty() || 10 * irregular.size() > regular.size()) break;
    }
    s->n_fused = (int)fused_lms.size();
    regular.swap(fused_lms);
    regular.insert(regular.end(), irregular.begin(), irregular.end());
  } else {
    s->n_fused = 0;
  }
  s->n_panels = (int)hpanels.size();
  s->pt_iperm.clear();
  s->pt_iperm.insert(s->pt_iperm.end(), regular.begin(), regular.end());
  s->pt_iperm.insert(s->pt_iperm.end(), big.begin(), big.end());
  s->n_lm = (int)s->pt_iperm.size();
  const int n_regular = (int)regular.size();
  s->pt_iperm.insert(s->pt_iperm.end(), rest.begin(), rest.end());
  s->pt_perm.assign(s->n_pt, -1);
  for (int q = 0; q < s->n_pt; ++q) s->pt_perm[s->pt_iperm[q]] = q;

  FIN_T("panels");
  // ---- reduced-system layout ----------------------------------------------------------
  // The non-eliminated parameter blocks (SE3 poses, SE2 poses, vectors, remaining points,
  // in table order) are packed into "supernodes" of <= kNB (32) tangent dimensions; every
  // supernode starts on a tile boundary of the reduced matrix (unused entries are padding:
  // identity diagonal, zero right-hand side).  The supernodes are then ordered by nested
  // dissection of their coupling graph -- or a greedy elimination order when that gives a
  // shallower elimination tree, see below -- so that the tile Cholesky's dependency DAG is a
  // bushy tree instead of a chain (trajectory-like problems are banded in table order).
  struct Item { int kind, idx, dof; };
  std::vector<Item> items;
  for (int i = 0; i < s->n_se3; ++i) if (!s->se3_const[i]) items.push_back({0, i, 6});
  for (int i = 0; i < s->n_se2; ++i) if (!s->se2_const[i]) items.push_back({1, i, 3});
  for (int i = 0; i < s->n_so3; ++i) if (!s->so3_const[i]) items.push_back({4, i, 3});
  for (int i = 0; i < s->n_vec; ++i) if (!s->vec_const[i] && s->vec_dims[i] > 0) items.push_back({3, i, s->vec_dims[i]});
  for (int q = s->n_lm; q < s->n_pt; ++q) if (!s->pt_const[s->pt_iperm[q]]) items.push_back({2, s->pt_iperm[q], 3});
  std::vector<int> sn_first, sn_tiles;          // supernode -> first item, number of tiles
  std::vector<int> item_sn(items.size());
  {
    int fill = bs::kNB + 1;
    for (size_t k = 0; k < items.size(); ++k) {
      if (fill + items[k].dof > bs::kNB) {       // open a new supernode
        sn_first.push_back((int)k);
        sn_tiles.push_back(std::max(1, cdiv(items[k].dof, bs::kNB)));
        fill = 0;
      }
      fill += items[k].dof;
      item_sn[k] = (int)sn_first.size() - 1;
    }
  }
  const int n_sn = (int)sn_first.size();
  std::vector<int> se3_sn(s->n_se3, -1), se2_sn(s->n_se2, -1), vec_sn(s->n_vec, -1), pt_sn(s->n_pt, -1), so3_sn(s->n_so3, -1);
  for (size_t k = 0; k < items.size(); ++k) {
    std::vector<int>& m = items[k].kind == 0 ? se3_sn : items[k].kind == 1 ? se2_sn : items[k].kind == 3 ? vec_sn : items[k].kind == 4 ? so3_sn : pt_sn;
    m[items[k].idx] = item_sn[k];
  }
  // coupling graph of the supernodes
  std::vector<std::vector<int>> adj(n_sn);
  {
    auto couple = [&](std::vector<int>& g) {
      std::sort(g.begin(), g.end());
      g.erase(std::unique(g.begin(), g.end()), g.end());
      for (int a : g)
        for (int b : g)
          if (a != b) adj[a].push_back(b);
    };
    std::vector<int> g;
    // poses that share an eliminated landmark (observations grouped by user point index)
    std::vector<int> ptr(s->n_pt + 1, 0), members(N);
    for (int i = 0; i < N; ++i) ptr[s->ob_pt[i] + 1]++;
    for (int p = 0; p < s->n_pt; ++p) ptr[p + 1] += ptr[p];
    {
      std::vector<int> cur(ptr.begin(), ptr.end() - 1);
      for (int i = 0; i < N; ++i) members[cur[s->ob_pt[i]]++] = s->ob_pose[i];
    }
    std::vector<int> last_sig;
    for (int p = 0; p < s->n_pt; ++p) {
      if (s->pt_perm[p] >= s->n_lm) continue;
      g.clear();
      for (int k = ptr[p]; k < ptr[p + 1]; ++k)
        if (se3_sn[members[k]] >= 0) g.push_back(se3_sn[members[k]]);
      std::sort(g.begin(), g.end());
      g.erase(std::unique(g.begin(), g.end()), g.end());
      if (g == last_sig) continue;               // same supernode set as the previous landmark
      last_sig = g;
      couple(g);
    }
    for (auto* b : s->edges) {
      if (!b->binary) continue;
      const std::vector<int>& m = b->group == 3 ? se3_sn : se2_sn;
      for (int e = 0; e < b->n; ++e) {
        g.clear();
        if (m[b->i1[e]] >= 0) g.push_back(m[b->i1[e]]);
        if (m[b->i2[e]] >= 0) g.push_back(m[b->i2[e]]);
        couple(g);
      }
    }
    for (auto* b : s->photos) {
      if (b->rot_idx < 0) continue;
      g.clear();
      if (so3_sn[b->rot_idx] >= 0) g.push_back(so3_sn[b->rot_idx]);
      if (vec_sn[b->vec_idx] >= 0) g.push_back(vec_sn[b->vec_idx]);
      couple(g);
    }
    for (size_t e = 0; e < s->cp_i1.size(); ++e) {
      const std::vector<int>& m = s->cp_group[e] == 3 ? se3_sn : se2_sn;
      g.clear();
      if (m[s->cp_i1[e]] >= 0) g.push_back(m[s->cp_i1[e]]);
      if (m[s->cp_i2[e]] >= 0) g.push_back(m[s->cp_i2[e]]);
      couple(g);
    }
    for (int b = 0; b < s->dn_blocks; ++b) {
      g.clear();
      for (int k = s->dn_pptr[b]; k < s->dn_pptr[b + 1]; ++k) {
        const int kind = s->dn_pkind[k], idx = s->dn_pindex[k];
        const int sn = kind == 0 ? se3_sn[idx] : kind == 1 ? se2_sn[idx] : kind == 2 ? pt_sn[idx] : vec_sn[idx];
        if (sn >= 0) g.push_back(sn);
      }
      couple(g);
    }
    for (auto& a : adj) {
      std::sort(a.begin(), a.end());
      a.erase(std::unique(a.begin(), a.end()), a.end());
    }
  }
  FIN_T("supernodes+adj");
  // nested dissection by recursive bisection of the table order
  std::vector<int> sn_order;
  {
    std::vector<int> side(n_sn, 0);
    std::vector<std::vector<int>> stack;
    std::vector<int> all(n_sn);
    std::iota(all.begin(), all.end(), 0);
    // explicit recursion: each work item is (nodes, separator to append after its children)
    struct Work { std::vector<int> nodes; bool emit; };
    std::vector<Work> todo;
    todo.push_back({all, false});
    while (!todo.empty()) {
      Work w = std::move(todo.back());
      todo.pop_back();
      if (w.emit || w.nodes.size() <= 2) {
        sn_order.insert(sn_order.end(), w.nodes.begin(), w.nodes.end());
        continue;
      }
      const size_t mid = w.nodes.size() / 2;
      std::vector<int> left(w.nodes.begin(), w.nodes.begin() + mid), right(w.nodes.begin() + mid, w.nodes.end());
      for (int u : left) side[u] = 1;
      for (int u : right) side[u] = 2;
      std::vector<int> sepL, sepR;
      for (int u : left)
        for (int v : adj[u])
          if (side[v] == 2) { sepL.push_back(u); break; }
      for (int u : right)
        for (int v : adj[u])
          if (side[v] == 1) { sepR.push_back(u); break; }
      for (int u : w.nodes) side[u] = 0;
      // the smaller separator; on ties the one that leaves the more balanced parts (for a chain this picks the
      // middle node of three and saves one level of the elimination tree)
      const size_t restL = std::max(left.size() - sepL.size(), right.size()), restR = std::max(left.size(), right.size() - sepR.size());
      const bool useL = sepL.size() != sepR.size() ? sepL.size() < sepR.size() : restL <= restR;
      std::vector<int>& sep = useL ? sepL : sepR;
      if (sep.size() * 2 >= w.nodes.size()) {      // no useful separator: keep the table order
        sn_order.insert(sn_order.end(), w.nodes.begin(), w.nodes.end());
        continue;
      }
      std::vector<int>& cut = useL ? left : right;
      std::vector<int> rest;
      for (int u : cut)
        if (!std::binary_search(sep.begin(), sep.end(), u)) rest.push_back(u);
      // order: left part, right part, separator  (stack is LIFO -> push in reverse)
      todo.push_back({sep, true});
      todo.push_back({useL ? right : rest, false});
      todo.push_back({useL ? rest : left, false});
    }
  }
  // Alternatives: greedy elimination orders.  Nested dissection of the table order is right for trajectory-like
  // (banded) graphs, where plain minimum degree would eat the chain from its ends (a sequential factorisation);
  // for graphs with long-range couplings (loop closures, random visibility) the bisection separators get wide
  // and dense and a greedy order gives a shallower elimination tree.  The tile Cholesky is bound by the LENGTH
  // OF THE CHAIN of dependent diagonal tiles, so the candidates are scored by the height of their elimination
  // tree (symbolic elimination on the supernode graph) and the shallowest wins; ties go to fewer filled tiles.
  if (n_sn > 2 && n_sn <= 768) {
    auto score = [&](const std::vector<int>& order, int& height, long long& fill) {
      std::vector<int> pos(n_sn);
      for (int i = 0; i < n_sn; ++i) pos[order[i]] = i;
      std::vector<uint8_t> m((size_t)n_sn * n_sn, 0);
      for (int u = 0; u < n_sn; ++u)
        for (int v : adj[u]) {
          const int hi = std::max(pos[u], pos[v]), lo = std::min(pos[u], pos[v]);
          m[(size_t)hi * n_sn + lo] = 1;
        }
      std::vector<int> rows, h(n_sn, 1);
      fill = 0; height = 1;
      for (int k = 0; k < n_sn; ++k) {
        rows.clear();
        for (int i = k + 1; i < n_sn; ++i)
          if (m[(size_t)i * n_sn + k]) rows.push_back(i);
        fill += (long long)rows.size();
        for (int a2 : rows) {
          h[a2] = std::max(h[a2], h[k] + 1);
          for (int b2 : rows)
            if (a2 > b2) m[(size_t)a2 * n_sn + b2] = 1;
        }
        height = std::max(height, h[k]);
      }
    };
    // greedy elimination orders: key (degree) = minimum degree; key (height of the vertex's subtree so far,
    // degree) = eliminate the shallowest vertices first, which keeps the elimination tree bushy
    auto greedy = [&](bool height_first) {
      std::vector<int> order;
      std::vector<std::vector<int>> nb(adj);
      std::vector<uint8_t> alive(n_sn, 1);
      std::vector<int> h(n_sn, 1);
      for (int step = 0; step < n_sn; ++step) {
        int best = -1;
        for (int i = 0; i < n_sn; ++i) {
          if (!alive[i]) continue;
          if (best < 0) { best = i; continue; }
          const bool better = height_first ? (h[i] != h[best] ? h[i] < h[best] : nb[i].size() < nb[best].size())
                                           : nb[i].size() < nb[best].size();
          if (better) best = i;
        }
        order.push_back(best);
        alive[best] = 0;
        const std::vector<int> ns = nb[best];
        for (int a2 : ns) {
          std::vector<int>& na = nb[a2];
          na.erase(std::remove(na.begin(), na.end(), best), na.end());
          h[a2] = std::max(h[a2], h[best] + 1);
          for (int b2 : ns)
            if (b2 != a2 && std::find(na.begin(), na.end(), b2) == na.end()) na.push_back(b2);
        }
      }
      return order;
    };
    int h_best = 0;
    long long f_best = 0;
    score(sn_order, h_best, f_best);
    for (int mode = 0; mode < 2; ++mode) {
      std::vector<int> cand = greedy(mode == 1);
      int h_c = 0;
      long long f_c = 0;
      score(cand, h_c, f_c);
      if (h_c < h_best || (h_c == h_best && f_c < f_best)) { sn_order.swap(cand); h_best = h_c; f_best = f_c; }
    }
  }
  FIN_T("ordering");
  // offsets
  s->se3_off.assign(s->n_se3, -1);
  s->se2_off.assign(s->n_se2, -1);
  s->vec_off.assign(s->n_vec, -1);
  s->so3_off.assign(s->n_so3, -1);
  s->vec_entry_off.assign(s->n_vec_entries, -1);
  s->pt_off_user.assign(s->n_pt, -1);
  s->pt_red_entry_off.assign(3 * (size_t)(s->n_pt - s->n_lm), -1);
  std::vector<uint8_t> used;
  int tile = 0;
  for (int sn : sn_order) {
    int off = tile * bs::kNB;
    const size_t k1 = sn + 1 < n_sn ? sn_first[sn + 1] : items.size();
    for (size_t k = sn_first[sn]; k < k1; ++k) {
      const Item& it = items[k];
      if (it.kind == 0) s->se3_off[it.idx] = off;
      else if (it.kind == 1) s->se2_off[it.idx] = off;
      else if (it.kind == 4) s->so3_off[it.idx] = off;
      else if (it.kind == 3) {
        s->vec_off[it.idx] = off;
        for (int c = 0; c < it.dof; ++c) s->vec_entry_off[s->vec_start[it.idx] + c] = off + c;
      } else {
        s->pt_off_user[it.idx] = off;
        for (int c = 0; c < 3; ++c) s->pt_red_entry_off[3 * (size_t)(s->pt_perm[it.idx] - s->n_lm) + c] = off + c;
      }
      off += it.dof;
    }
    tile += sn_tiles[sn];
  }
  s->nblk = std::max(1, tile);
  s->n_pad = s->nblk * bs::kNB;
  s->n_red = s->n_pad;                         // the reduced span includes the padding entries
  used.assign(s->n_pad, 0);
  auto mark_used = [&](int off, int dof) { for (int c = 0; c < dof; ++c) used[off + c] = 1; };
  for (const Item& it : items) {
    const int off = it.kind == 0 ? s->se3_off[it.idx] : it.kind == 1 ? s->se2_off[it.idx] : it.kind == 3 ? s->vec_off[it.idx]
                  : it.kind == 4 ? s->so3_off[it.idx] : s->pt_off_user[it.idx];
    mark_used(off, it.dof);
  }
  for (int q = 0; q < s->n_lm; ++q) s->pt_off_user[s->pt_iperm[q]] = s->n_red + 3 * q;
  s->dim = s->n_red + 3 * s->n_lm;

  {   // tile of the reduced system -> SE3 poses whose tangent block starts inside it (a block never straddles tiles)
    std::vector<int> ptr(s->nblk + 1, 0), idx;
    for (int i = 0; i < s->n_se3; ++i) if (s->se3_off[i] >= 0) ptr[s->se3_off[i] / bs::kNB + 1]++;
    for (int t = 0; t < s->nblk; ++t) ptr[t + 1] += ptr[t];
    idx.assign(std::max(1, ptr[s->nblk]), 0);
    std::vector<int> cur(ptr.begin(), ptr.end() - 1);
    for (int i = 0; i < s->n_se3; ++i) if (s->se3_off[i] >= 0) idx[cur[s->se3_off[i] / bs::kNB]++] = i;
    CU(upload(s->d_tile_pose_ptr, ptr, s->stream));
    CU(upload(s->d_tile_pose, idx, s->stream));
  }

  FIN_T("offsets");
  // ---- observations sorted by internal point index (stable) ----
  std::vector<int> order(N);
  {                                    // stable counting sort: the CSR above already groups by point in registration order
    std::vector<int> at(s->n_pt + 1, 0);
    for (int p = 0; p < s->n_pt; ++p) at[s->pt_perm[p] + 1] = obs_ptr[p + 1] - obs_ptr[p];
    for (int q = 0; q < s->n_pt; ++q) at[q + 1] += at[q];
    for (int p = 0; p < s->n_pt; ++p)
      std::copy(obs_ids.begin() + obs_ptr[p], obs_ids.begin() + obs_ptr[p + 1], order.begin() + at[s->pt_perm[p]]);
  }
  std::vector<double> ou(N), ov(N), od(N);
  std::vector<int> opose(N), opt(N), ogrp(N), lm_start(s->n_lm + 1, 0);
  for (int k = 0; k < N; ++k) {
    const int i = order[k];
    ou[k] = s->ob_uvd[3 * (size_t)i]; ov[k] = s->ob_uvd[3 * (size_t)i + 1]; od[k] = s->ob_uvd[3 * (size_t)i + 2];
    opose[k] = s->ob_pose[i];
    opt[k] = s->pt_perm[s->ob_pt[i]];
    ogrp[k] = s->ob_grp[i];
    if (opt[k] < s->n_lm) lm_start[opt[k] + 1]++;
  }
  for (int q = 0; q < s->n_lm; ++q) lm_start[q + 1] += lm_start[q];

  // ---- panel grids: cell (row, landmark) -> observation, padded to kPanelLm landmarks per row ----
  FIN_T("sort obs");
  std::vector<bs::PanelDesc> pdescs;
  std::vector<double> pobs;
  std::vector<unsigned short> pgrp;
  s->panel_max_var = 1;
  pdescs.reserve(hpanels.size());
  pobs.assign(std::max<size_t>(1, hpanels.size()) * bs::kPanelObs, 0.0);
  pgrp.assign(std::max<size_t>(1, hpanels.size()) * bs::kPanelRows * bs::kPanelLm, 0);
  for (size_t ip = 0; ip < hpanels.size(); ++ip) {
    const HostPanel& hp = hpanels[ip];
    bs::PanelDesc pd{};
    pd.hdr.lm_begin = hp.first; pd.hdr.n_lms = hp.n_lms;
    pd.hdr.n_rows = (int)hp.poses.size(); pd.hdr.n_var = hp.n_var;
    s->panel_max_var = std::max(s->panel_max_var, hp.n_var);
    for (int r = 0; r < bs::kPanelRows; ++r) { pd.rows[r].pose = 0; pd.rows[r].off = -1; }
    for (int r = 0; r < pd.hdr.n_rows; ++r) {
      pd.rows[r].pose = hp.poses[r];
      pd.rows[r].off = s->se3_off[hp.poses[r]];
    }
    double* ob = pobs.data() + ip * bs::kPanelObs;
    for (int j = 0; j < hp.n_lms; ++j) {
      const int q = hp.first + j;
      for (int k = lm_start[q]; k < lm_start[q + 1]; ++k) {
        const int r = (int)(std::find(hp.poses.begin(), hp.poses.end(), opose[k]) - hp.poses.begin());
        bs::PanelRow& row = pd.rows[r];
        if (j < 32) row.mask_lo |= 1u << j; else row.mask_hi |= 1u << (j - 32);
        const size_t cell = (size_t)r * bs::kPanelLm + j;
        ob[cell] = ou[k];
        ob[bs::kPanelRows * bs::kPanelLm + cell] = ov[k];
        ob[2 * bs::kPanelRows * bs::kPanelLm + cell] = od[k];
        pgrp[(ip * bs::kPanelRows + r) * bs::kPanelLm + j] = (unsigned short)ogrp[k];
      }
    }
    pdescs.push_back(pd);
  }
  if (pdescs.empty()) pdescs.push_back(bs::PanelDesc{});

  FIN_T("panel grids");
  // ---- landmark blocks: whole landmarks, <= kBlkObs observations, bounded Schur operands ----
  // Inside a block the observations are re-ordered SLOT-MAJOR (grouped by pose, constant poses last):
  // a warp of the block kernels then reads one or two poses (shared-memory broadcasts) and the camera-side
  // reduction runs over contiguous rows.  lm_obs / lm_obs_local map the landmark-ordered CSR positions
  // (lm_start) to the observations' new places.
  const std::vector<int> opose_lm = opose;      // landmark-sorted order (build_tile_mask below)
  std::vector<bs::LmBlock> blocks;
  std::vector<int> slot_pose, lm_obs(N);
  std::vector<unsigned char> lm_obs_local(N, 0), seg_start;
  std::vector<unsigned> obs_code(N, 255u);
  std::iota(lm_obs.begin(), lm_obs.end(), 0);
  std::vector<bs::SchurDesc> sch_descs;
  std::vector<bs::SchurPair> sch_pairs;
  std::vector<unsigned> sch_combos;             // runs of combos, see schur.cuh
  int max_pairs = 1, max_combos = 4, max_lms = 1;
  {
    int q = 0;
    std::vector<int> cur;                       // distinct variable poses of the block under construction
    std::vector<int> slot_of, new_pos, pair_cnt, pair_slot, gen_pair;
    std::vector<unsigned short> gen_combo;
    std::vector<double> tu, tv, td;
    std::vector<int> tpose, tpt, tgrp;
    s->nb_fused = 0;
    while (q < n_regular) {
      if (q == s->n_fused) s->nb_fused = (int)blocks.size();
      const int q_lim = q < s->n_fused ? s->n_fused : n_regular;     // no block straddles the panel boundary
      bs::LmBlock b{};
      b.obs_begin = lm_start[q]; b.lm_begin = q;
      cur.clear();
      int q1 = q;
      while (q1 < q_lim && lm_start[q1 + 1] - b.obs_begin <= bs::kBlkObs) {
        for (int k = lm_start[q1]; k < lm_start[q1 + 1]; ++k)
          if (s->se3_off[opose[k]] >= 0 && std::find(cur.begin(), cur.end(), opose[k]) == cur.end())
            cur.push_back(opose[k]);
        ++q1;
      }
      b.n_lms = q1 - q; b.n_obs = lm_start[q1] - b.obs_begin;
      b.slot_begin = (int)slot_pose.size(); b.seg_begin = (int)seg_start.size();
      slot_pose.insert(slot_pose.end(), cur.begin(), cur.end());
      b.n_slots = (int)cur.size();
      slot_of.assign(b.n_obs, 255);
      for (int k = 0; k < b.n_obs; ++k) {
        const int pose = opose[b.obs_begin + k];
        if (s->se3_off[pose] < 0) continue;
        slot_of[k] = (int)(std::find(cur.begin(), cur.end(), pose) - cur.begin());
      }
      // new position of every observation (k = landmark-ordered position inside the block)
      new_pos.assign(b.n_obs, 0);
      int pos = 0;
      for (int sl = 0; sl < b.n_slots; ++sl) {
        seg_start.push_back((unsigned char)pos);
        for (int k = 0; k < b.n_obs; ++k)
          if (slot_of[k] == sl) new_pos[k] = pos++;
      }
      seg_start.push_back((unsigned char)pos);
      for (int k = 0; k < b.n_obs; ++k)
        if (slot_of[k] == 255) new_pos[k] = pos++;
      tu.resize(b.n_obs); tv.resize(b.n_obs); td.resize(b.n_obs); tpose.resize(b.n_obs); tpt.resize(b.n_obs); tgrp.resize(b.n_obs);
      for (int k = 0; k < b.n_obs; ++k) {
        const int src = b.obs_begin + k, dst = new_pos[k];
        tu[dst] = ou[src]; tv[dst] = ov[src]; td[dst] = od[src];
        tpose[dst] = opose[src]; tpt[dst] = opt[src]; tgrp[dst] = ogrp[src];
        obs_code[b.obs_begin + dst] = (unsigned)slot_of[k] | ((unsigned)(opt[src] - b.lm_begin) << 8) | ((unsigned)ogrp[src] << 16);
        lm_obs[src] = b.obs_begin + dst;
        lm_obs_local[src] = (unsigned char)dst;
      }
      for (int k = 0; k < b.n_obs; ++k) {
        const int dst = b.obs_begin + k;
        ou[dst] = tu[k]; ov[dst] = tv[k]; od[dst] = td[k]; opose[dst] = tpose[k]; opt[dst] = tpt[k]; ogrp[dst] = tgrp[k];
      }
      // Schur combos of the block: for every landmark and every unordered pair of its (variable-pose)
      // observations, (row observation, col observation) with the row pose the one at the larger reduced
      // offset (lower triangle); grouped by slot pair, longest pairs first (balanced round-robin over warps)
      {
        // bucketed by slot pair with a stable counting pass (combos are generated in landmark order, which inside a
        // pair is the order of their row positions)
        const int ns = b.n_slots;
        pair_cnt.assign((size_t)ns * ns + 1, 0);
        gen_pair.clear(); gen_combo.clear();
        for (int l = 0; l < b.n_lms; ++l) {
          const int k0 = lm_start[b.lm_begin + l] - b.obs_begin, k1 = lm_start[b.lm_begin + l + 1] - b.obs_begin;
          for (int x = k0; x < k1; ++x) {
            if (slot_of[x] == 255) continue;
            for (int y = k0; y <= x; ++y) {
              if (slot_of[y] == 255) continue;
              int ro = x, co = y;
              if (s->se3_off[cur[slot_of[ro]]] < s->se3_off[cur[slot_of[co]]]) std::swap(ro, co);
              const int pk = slot_of[ro] * ns + slot_of[co];
              gen_pair.push_back(pk);
              gen_combo.push_back((unsigned short)(new_pos[ro] | (new_pos[co] << 8)));
              pair_cnt[pk + 1]++;
            }
          }
        }
        std::vector<std::pair<std::pair<int, int>, std::vector<unsigned short>>> pv;
        pair_slot.assign((size_t)ns * ns, -1);
        for (int pk = 0; pk < ns * ns; ++pk)
          if (pair_cnt[pk + 1] > 0) {
            pair_slot[pk] = (int)pv.size();
            pv.emplace_back(std::make_pair(pk / ns, pk % ns), std::vector<unsigned short>());
            pv.back().second.reserve(pair_cnt[pk + 1]);
          }
        for (size_t k = 0; k < gen_pair.size(); ++k) pv[pair_slot[gen_pair[k]]].second.push_back(gen_combo[k]);
        std::stable_sort(pv.begin(), pv.end(), [](const auto& x, const auto& y) { return x.second.size() > y.second.size(); });
        const int cb0 = (int)sch_combos.size(), pb0 = (int)sch_pairs.size();
        for (auto& e : pv) {
          // combos -> runs of (row + i, col + i)
          std::vector<unsigned short>& cl = e.second;
          const auto by_row = [](unsigned short x, unsigned short y) { return (x & 255) != (y & 255) ? (x & 255) < (y & 255) : x < y; };
          if (!std::is_sorted(cl.begin(), cl.end(), by_row)) std::sort(cl.begin(), cl.end(), by_row);
          const int rb0 = (int)sch_combos.size();
          for (size_t k = 0; k < cl.size();) {
            size_t k2 = k + 1;
            while (k2 < cl.size() && (cl[k2] & 255) == (cl[k] & 255) + (k2 - k) && (cl[k2] >> 8) == (cl[k] >> 8) + (k2 - k)) ++k2;
            sch_combos.push_back((unsigned)cl[k] | ((unsigned)(k2 - k) << 16));
            k = k2;
          }
          bs::SchurPair P;
          P.slots_n = (unsigned)e.first.first | ((unsigned)e.first.second << 8) | ((unsigned)(sch_combos.size() - rb0) << 16);
          P.rbeg = rb0 - cb0;
          sch_pairs.push_back(P);
        }
        {
          bs::SchurDesc sd{};
          sd.obs_begin = b.obs_begin; sd.n_obs = b.n_obs; sd.lm_begin = b.lm_begin; sd.n_lms = b.n_lms;
          sd.slot_begin = b.slot_begin; sd.n_slots = b.n_slots;
          sd.pair_begin = pb0; sd.n_pairs = (int)sch_pairs.size() - pb0;
          sd.run_begin = cb0; sd.n_runs = (int)sch_combos.size() - cb0;
          sch_descs.push_back(sd);
        }
        max_pairs = std::max(max_pairs, (int)pv.size());
        max_combos = std::max(max_combos, (int)sch_combos.size() - cb0);
        max_lms = std::max(max_lms, b.n_lms);
      }
      blocks.push_back(b);
      q = q1;
    }
    if (s->n_fused >= n_regular) s->nb_fused = (int)blocks.size();
  }
  s->loss_kind = s->groups.size() == 1 ? s->groups[0].loss.kind : -1;
  NEED(s->groups.size() < 65536, "too many reprojection groups (%zu)", s->groups.size());
  s->schur_smem = blocks.empty() ? 0 : bs::schur_smem_bytes(max_lms, max_pairs, max_combos);
  s->schur_max_lms = max_lms; s->schur_max_pairs = max_pairs; s->schur_max_runs = max_combos;
  if (sch_descs.empty()) sch_descs.push_back(bs::SchurDesc{});
  NEED(s->schur_smem <= 200 * 1024, "landmark block structure needs %zu bytes of shared memory", s->schur_smem);
  if (sch_pairs.empty()) sch_pairs.push_back(bs::SchurPair{0u, 0});
  if (sch_combos.empty()) sch_combos.push_back(0);
  s->n_regular = n_regular;
  std::vector<int> slot_off(slot_pose.size());
  for (size_t k = 0; k < slot_pose.size(); ++k) slot_off[k] = s->se3_off[slot_pose[k]];
  s->n_slot_entries = (int)slot_pose.size();
  s->stage_len = 2;
  s->max_slots = 1;
  for (const auto& b : blocks) s->stage_len = std::max(s->stage_len, 12 * b.n_slots + 3 * b.n_lms);
  for (const auto& b : blocks) s->max_slots = std::max(s->max_slots, b.n_slots);
  s->stage_len = (s->stage_len + 1) & ~1;
  s->n_lmblocks = (int)blocks.size();
  s->tail_begin = lm_start[n_regular];
  if (slot_pose.empty()) slot_pose.push_back(0);
  if (seg_start.empty()) seg_start.push_back(0);

  FIN_T("lm blocks");
  // ---- dense-block structure ----
  s->dn_row_ptr.assign(1, 0); s->dn_col_ptr.assign(1, 0); s->dn_j_ptr.assign(1, 0);
  s->dn_col_index.clear();
  for (int b = 0; b < s->dn_blocks; ++b) {
    int ncols = 0;
    for (int k = s->dn_pptr[b]; k < s->dn_pptr[b + 1]; ++k) {
      const int kind = s->dn_pkind[k], idx = s->dn_pindex[k];
      int dof, o;
      if (kind == 0) { dof = 6; o = s->se3_off[idx]; }
      else if (kind == 1) { dof = 3; o = s->se2_off[idx]; }
      else if (kind == 2) { dof = 3; o = s->pt_off_user[idx]; }
      else { dof = s->vec_dims[idx]; o = s->vec_off[idx]; }
      for (int c = 0; c < dof; ++c) s->dn_col_index.push_back(o < 0 ? -1 : o + c);
      ncols += dof;
    }
    s->dn_row_ptr.push_back(s->dn_row_ptr.back() + s->dn_rows[b]);
    s->dn_col_ptr.push_back(s->dn_col_ptr.back() + ncols);
    s->dn_j_ptr.push_back(s->dn_j_ptr.back() + (long long)s->dn_rows[b] * ncols);
  }

  FIN_T("dense struct");
  // ---- device memory ----
  cudaStream_t st = s->stream;
  std::vector<double> pts_int(3 * (size_t)s->n_pt + 2, 0.0);    // + 2: a panel's point run is copied in 16-byte units
  for (int q = 0; q < s->n_pt; ++q)
    for (int k = 0; k < 3; ++k) pts_int[3 * (size_t)q + k] = s->h_pts[3 * (size_t)s->pt_iperm[q] + k];
  CU(upload(s->d_se3, s->h_se3, st));
  CU(upload(s->d_se2, s->h_se2, st));
  CU(upload(s->d_so3, s->h_so3, st));
  CU(upload(s->d_so3_off, s->so3_off, st));
  CU(s->b_so3.alloc(s->d_so3.n));
  CU(upload(s->d_pts, pts_int, st));
  CU(upload(s->d_vec, s->h_vec, st));
  CU(s->d_stage.alloc(3 * (size_t)s->n_pt));
  CU(upload(s->d_pt_perm, s->pt_perm, st));
  CU(upload(s->d_used, used, st));
  CU(upload(s->d_se3_off, s->se3_off, st));
  CU(upload(s->d_se2_off, s->se2_off, st));
  CU(upload(s->d_vec_entry_off, s->vec_entry_off, st));
  CU(upload(s->d_ptred_entry_off, s->pt_red_entry_off, st));
  CU(upload(s->d_ou, ou, st)); CU(upload(s->d_ov, ov, st)); CU(upload(s->d_od, od, st));
  CU(upload(s->d_opose, opose, st)); CU(upload(s->d_opt, opt, st)); CU(upload(s->d_ogrp, ogrp, st));
  CU(upload(s->d_lm_start, lm_start, st));
  CU(upload(s->d_blocks, blocks, st));
  CU(upload(s->d_slot_pose, slot_pose, st));
  CU(upload(s->d_lm_obs, lm_obs, st));
  CU(upload(s->d_lm_obs_local, lm_obs_local, st));
  CU(upload(s->d_seg_start, seg_start, st));
  CU(upload(s->d_obs_code, obs_code, st));
  CU(upload(s->d_sch_descs, sch_descs, st));
  CU(upload(s->d_sch_pairs, sch_pairs, st));
  CU(upload(s->d_sch_combos, sch_combos, st));
  CU(upload(s->d_slot_off, slot_off, st));
  CU(s->d_slot_poses.alloc(12 * slot_pose.size()));
  CU(s->d_slot_dx.alloc(6 * slot_pose.size()));
  {                        // blocks with many poses: staging (dynamic) + rows (static) may exceed the 48 KB default
    const void* fn = nullptr;
    switch (s->loss_kind) {
      case 0: fn = (const void*)bs::reproj_block_kernel<0>; break;
      case 1: fn = (const void*)bs::reproj_block_kernel<1>; break;
      case 2: fn = (const void*)bs::reproj_block_kernel<2>; break;
      case 3: fn = (const void*)bs::reproj_block_kernel<3>; break;
      case 4: fn = (const void*)bs::reproj_block_kernel<4>; break;
      case 5: fn = (const void*)bs::reproj_block_kernel<5>; break;
      default: fn = (const void*)bs::reproj_block_kernel<-1>; break;
    }
    CU(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * (size_t)s->stage_len * sizeof(double))));
  }
  if (s->schur_smem > 0)   // static + dynamic shared memory may exceed the 48 KB default
    CU(cudaFuncSetAttribute(bs::schur_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->schur_smem));
  {
    int per_sm = 1, sms = 148;
    if (s->schur_smem > 0)
      CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bs::schur_block_kernel, bs::kSchurThreads, s->schur_smem));
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device));
    s->schur_grid = std::max(1, per_sm) * sms;       // persistent CTAs, all resident
  }
  CU(upload(s->d_groups, s->groups, st));
  for (auto* b : s->edges) {
    CU(upload(b->d_i1, b->i1, st));
    CU(upload(b->d_i2, b->i2, st));
    CU(upload(b->d_Tobs, b->Tobs, st));
    CU(upload(b->d_stiff, b->stiff, st));
  }
  for (auto* b : s->motions) {
    CU(upload(b->d_pts1, b->pts1, st));
    CU(upload(b->d_obs2, b->obs2, st));
  }
  for (auto* b : s->photos) {
    CU(upload(b->d_uvd, b->uvd, st));
    CU(upload(b->d_im_ref, b->im_ref, st));
    CU(upload(b->d_im_jac, b->im_jac, st));
    CU(upload(b->d_im_track, b->im_track, st));
  }
  CU(upload(s->d_dn_row_ptr, s->dn_row_ptr, st));
  CU(upload(s->d_dn_col_ptr, s->dn_col_ptr, st));
  CU(upload(s->d_dn_j_ptr, s->dn_j_ptr, st));
  CU(upload(s->d_dn_col_index, s->dn_col_index, st));
  CU(s->d_dn_J.alloc((size_t)s->dn_j_ptr.back()));
  CU(s->d_dn_e.alloc((size_t)s->dn_row_ptr.back()));
  s->d_W.release();                             // allocated on first use (ensure_W)
  CU(upload(s->d_pdescs, pdescs, st));
  CU(upload(s->d_pobs, pobs, st));
  CU(upload(s->d_pgrp, pgrp, st));
  CU(s->d_se3_prev.alloc(std::max<size_t>(1, s->d_se3.n)));
  if (s->n_panels > 0) {
    const size_t psm = bs::panel_smem_bytes(s->panel_max_var), fsm = 0;
    const void *fk = nullptr, *ff = nullptr;
    switch (s->loss_kind) {
      case 0: fk = (const void*)bs::fused_panel_kernel<0>; ff = (const void*)bs::panel_finish_kernel<0>; break;
      case 1: fk = (const void*)bs::fused_panel_kernel<1>; ff = (const void*)bs::panel_finish_kernel<1>; break;
      case 2: fk = (const void*)bs::fused_panel_kernel<2>; ff = (const void*)bs::panel_finish_kernel<2>; break;
      case 3: fk = (const void*)bs::fused_panel_kernel<3>; ff = (const void*)bs::panel_finish_kernel<3>; break;
      case 4: fk = (const void*)bs::fused_panel_kernel<4>; ff = (const void*)bs::panel_finish_kernel<4>; break;
      case 5: fk = (const void*)bs::fused_panel_kernel<5>; ff = (const void*)bs::panel_finish_kernel<5>; break;
      default: fk = (const void*)bs::fused_panel_kernel<-1>; ff = (const void*)bs::panel_finish_kernel<-1>; break;
    }
    CU(cudaFuncSetAttribute(fk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psm));
    int per_sm = 1, per_sm_f = 1, sms = 148;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fk, bs::kPanelThreads, psm));
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_f, ff, bs::kFinishThreads, fsm));
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device));
    NEED(per_sm >= 1 && per_sm_f >= 1, "panel kernels do not fit on an SM (%zu bytes of shared memory)", psm);
    s->panel_grid = per_sm * sms;               // persistent CTAs, all resident
    s->finish_grid = per_sm_f * sms;
  }
  CU(s->d_Vg.alloc(9 * (size_t)s->n_lm));
  CU(s->d_Vinv.alloc(6 * (size_t)s->n_lm));
  CU(s->d_red.alloc(s->red_len()));
  CU(s->d_dx.alloc((size_t)s->n_pad + 3 * (size_t)s->n_lm));
  CU(s->d_Linv.alloc((size_t)s->nblk * bs::kNB * bs::kNB));
  CU(s->b_se3.alloc(s->d_se3.n)); CU(s->b_se2.alloc(s->d_se2.n));
  CU(s->b_pts.alloc(s->d_pts.n)); CU(s->b_vec.alloc(s->d_vec.n));
  CU(cudaMemsetAsync(s->d_red.p, 0, s->red_len() * sizeof(double), st));
  CU(cudaMemsetAsync(s->d_dx.p, 0, s->d_dx.n * sizeof(double), st));
  CU(cudaStreamSynchronize(st));
  FIN_T("uploads");
  build_tile_mask(s, opose_lm, lm_start);
  FIN_T("tile mask");
  s->h_opose_lm = opose_lm;
  s->h_lm_start = lm_start;
  s->finalized = true;
  s->dn_uploaded = false;
  return BSLAM_OK;
}

int bslam_get_layout(bslam_solver* s, int32_t* se3_off, int32_t* se2_off, int32_t* pt_off, int32_t* vec_off, int32_t* dim,
                     int32_t* n_reduced) {
  NEED(s && s->finalized, "bslam_get_layout: solver not finalized");
  if (se3_off) std::copy(s->se3_off.begin(), s->se3_off.end(), se3_off);
  if (se2_off) std::copy(s->se2_off.begin(), s->se2_off.end(), se2_off);
  if (pt_off) std::copy(s->pt_off_user.begin(), s->pt_off_user.end(), pt_off);
  if (vec_off) std::copy(s->vec_off.begin(), s->vec_off.end(), vec_off);
  if (dim) *dim = s->dim;
  if (n_reduced) *n_reduced = s->n_red;
  return BSLAM_OK;
}

int bslam_get_layout_so3(bslam_solver* s, int32_t* so3_off) {
  NEED(s && s->finalized, "bslam_get_layout_so3: solver not finalized");
  if (so3_off) std::copy(s->so3_off.begin(), s->so3_off.end(), so3_off);
  return BSLAM_OK;
}

// ------------------------------------------------------------------ hot path

int bslam_eval_cost(bslam_solver* s, double* cost) {
  NEED(s && s->finalized, "bslam_eval_cost: solver not finalized");
  CU(cudaSetDevice(s->device));
  CU(cudaMemsetAsync(s->scalars() + BSLAM_S_COST_EVAL, 0, sizeof(double), s->stream));
  launch_cost(s, BSLAM_S_COST_EVAL);
  CU(cudaGetLastError());
  const bool t = s->timing;
  s->timing = false;
  int rc = fetch_scalars(s);
  s->timing = t;
  if (rc) return rc;
  if (cost) *cost = s->h_scalars[BSLAM_S_COST_EVAL];
  return BSLAM_OK;
}

int bslam_linearize(bslam_solver* s, double* cost_lin) {
  NEED(s, "NULL solver");
  CU(cudaSetDevice(s->device));
  int rc = do_linearize(s, false);
  if (rc) return rc;
  if (cost_lin) {
    CU(cudaMemcpyAsync(s->h_scalars, s->scalars(), BSLAM_N_SCALARS * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    *cost_lin = s->h_scalars[BSLAM_S_COST_LIN];
  }
  return BSLAM_OK;
}

int bslam_reduce(bslam_solver* s, double lambda) {
  NEED(s && s->finalized, "bslam_reduce: solver not finalized");
  NEED(lambda >= 0.0, "bslam_reduce: lambda must be >= 0");
  CU(cudaSetDevice(s->device));
  return do_reduce(s, lambda, false);
}

int bslam_solve_reduced(bslam_solver* s) {
  NEED(s && s->finalized, "bslam_solve_reduced: solver not finalized");
  CU(cudaSetDevice(s->device));
  return do_solve_reduced(s);
}

int bslam_retract(bslam_solver* s, int eval_new_cost) {
  NEED(s && s->finalized, "bslam_retract: solver not finalized");
  CU(cudaSetDevice(s->device));
  return do_retract(s, eval_new_cost, false);
}

int bslam_linearize_reduce(bslam_solver* s, double lambda) {
  NEED(s && s->finalized, "bslam_linearize_reduce: solver not finalized");
  NEED(lambda >= 0.0, "bslam_linearize_reduce: lambda must be >= 0");
  CU(cudaSetDevice(s->device));
  int rc;
  if ((rc = prepare_iterate(s))) return rc;
  if ((rc = do_linearize(s, true))) return rc;
  return do_reduce(s, lambda, true);
}

int bslam_retract_iterate(bslam_solver* s, int eval_new_cost) {
  NEED(s && s->finalized, "bslam_retract_iterate: solver not finalized");
  CU(cudaSetDevice(s->device));
  return do_retract(s, eval_new_cost, true);
}

int bslam_set_fused(bslam_solver* s, int mode) {
  NEED(s, "NULL solver");
  NEED(!s->finalized, "bslam_set_fused after finalize; call bslam_clear_blocks first");
  NEED(mode >= 0 && mode <= 2, "bslam_set_fused: mode must be 0, 1 or 2");
  s->fused_mode = mode;
  return BSLAM_OK;
}

int bslam_get_fused(bslam_solver* s, int32_t* n_panels, int32_t* n_landmarks) {
  NEED(s && s->finalized, "bslam_get_fused: solver not finalized");
  if (n_panels) *n_panels = s->n_panels;
  if (n_landmarks) *n_landmarks = s->n_fused;
  return BSLAM_OK;
}

int bslam_get_scalars(bslam_solver* s, double* out) {
  NEED(s && s->finalized, "bslam_get_scalars: solver not finalized");
  CU(cudaSetDevice(s->device));
  int rc = fetch_scalars(s);
  if (rc) return rc;
  if (out) std::memcpy(out, s->h_scalars, BSLAM_N_SCALARS * sizeof(double));
  return BSLAM_OK;
}

int bslam_last_scalars(bslam_solver* s, double* out) {
  NEED(s && s->finalized && out, "bslam_last_scalars: bad arguments");
  std::memcpy(out, s->h_scalars, BSLAM_N_SCALARS * sizeof(double));     // pinned mirror filled by the last iterate / get_scalars
  return BSLAM_OK;
}

// One full iteration enqueued on the handle's stream, scalars copied to the pinned host mirror, NO synchronisation.
static int iterate_enqueue(bslam_solver* s, double lambda, int eval_new_cost) {
  int rc;
  const bool graphable = s->use_graph && !s->timing && s->dn_blocks == 0 && s->d_trace.p == nullptr;
  if (graphable) {
    // the whole iteration (6 kernels + the scalar read-back) is one graph launch
    if ((rc = prepare_iterate(s))) return rc;
    if (s->graph_exec && (s->graph_lambda != lambda || s->graph_eval != eval_new_cost)) {
      cudaGraphExecDestroy(s->graph_exec);
      s->graph_exec = nullptr;
    }
    s->graph_lambda = lambda;
    s->graph_eval = eval_new_cost;
    return run_graphed(s, &s->graph_exec, &s->graph_launches, [&]() {
      int r = do_linearize(s, true);
      if (!r) r = do_reduce(s, lambda, true);
      if (!r && s->world > 1) r = do_peer_publish(s);
      if (!r) r = do_solve_reduced(s, fuse_retract(s, true));
      if (!r) r = do_retract(s, eval_new_cost, true, fuse_retract(s, true));
      if (!r && s->world > 1) r = do_peer_scalars(s);
      if (!r && cudaMemcpyAsync(s->h_scalars, s->scalars(), BSLAM_N_SCALARS * sizeof(double), cudaMemcpyDeviceToHost,
                                s->stream) != cudaSuccess)
        r = fail(s, BSLAM_E_CUDA, "scalar read-back could not be captured");
      return r;
    });
  }
  if ((rc = do_linearize(s, true))) return rc;
  if ((rc = do_reduce(s, lambda, true))) return rc;
  if (s->world > 1 && (rc = do_peer_publish(s))) return rc;
  if ((rc = do_solve_reduced(s, fuse_retract(s, true)))) return rc;
  if ((rc = do_retract(s, eval_new_cost, true, fuse_retract(s, true)))) return rc;
  if (s->world > 1 && (rc = do_peer_scalars(s))) return rc;
  CU(cudaMemcpyAsync(s->h_scalars, s->scalars(), BSLAM_N_SCALARS * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  return BSLAM_OK;
}

static int iterate_finish(bslam_solver* s, double* cost_lin, double* cost_new, double* dx_norm) {
  int rc;
  if ((rc = sync_and_timings(s))) return rc;       // one synchronisation (the scalar read-back is already enqueued)
  if (s->h_scalars[BSLAM_S_PEER_TIMEOUT] != 0.0)
    return fail(s, BSLAM_E_CUDA, "peer exchange timed out: a rank of the sharded iteration did not arrive");
  if (cost_lin) *cost_lin = s->h_scalars[BSLAM_S_COST_LIN];
  if (cost_new) *cost_new = s->h_scalars[BSLAM_S_COST_NEW];
  if (dx_norm) *dx_norm = std::sqrt(s->h_scalars[BSLAM_S_DX_NORM2]);
  return BSLAM_OK;
}

int bslam_iterate(bslam_solver* s, double lambda, int eval_new_cost, double* cost_lin, double* cost_new, double* dx_norm) {
  NEED(s && s->finalized, "bslam_iterate: solver not finalized");
  NEED(lambda >= 0.0, "bslam_iterate: lambda must be >= 0");
  CU(cudaSetDevice(s->device));
  int rc;
  if ((rc = iterate_enqueue(s, lambda, eval_new_cost))) return rc;
  return iterate_finish(s, cost_lin, cost_new, dx_norm);
}

int bslam_iterate_async(bslam_solver* s, double lambda, int eval_new_cost) {
  NEED(s && s->finalized, "bslam_iterate_async: solver not finalized");
  NEED(lambda >= 0.0, "bslam_iterate_async: lambda must be >= 0");
  CU(cudaSetDevice(s->device));
  return iterate_enqueue(s, lambda, eval_new_cost);
}

int bslam_iterate_wait(bslam_solver* s, double* cost_lin, double* cost_new, double* dx_norm) {
  NEED(s && s->finalized, "bslam_iterate_wait: solver not finalized");
  CU(cudaSetDevice(s->device));
  return iterate_finish(s, cost_lin, cost_new, dx_norm);
}

int bslam_add_coupling(bslam_solver* s, int group, int n, const int32_t* idx1, const int32_t* idx2) {
  NEED(s, "NULL solver");
  NEED(!s->finalized, "bslam_add_coupling after finalize; call bslam_clear_blocks first");
  NEED(group == BSLAM_SE2 || group == BSLAM_SE3, "bslam_add_coupling: group must be BSLAM_SE2 or BSLAM_SE3");
  NEED(n >= 0 && (n == 0 || (idx1 && idx2)), "bslam_add_coupling: bad arguments");
  const int table = group == 3 ? s->n_se3 : s->n_se2;
  for (int i = 0; i < n; ++i)
    NEED(idx1[i] >= 0 && idx1[i] < table && idx2[i] >= 0 && idx2[i] < table, "bslam_add_coupling %d: pose index outside the table (%d)", i, table);
  for (int i = 0; i < n; ++i) { s->cp_group.push_back(group); s->cp_i1.push_back(idx1[i]); s->cp_i2.push_back(idx2[i]); }
  return BSLAM_OK;
}

int bslam_layout_hash(bslam_solver* s, uint64_t* hash) {
  NEED(s && s->finalized && hash, "bslam_layout_hash: bad arguments");
  uint64_t h = 1469598103934665603ULL;                       // FNV-1a over the reduced offsets and the tile structure
  auto mix = [&](const void* p, size_t n) {
    const unsigned char* b = static_cast<const unsigned char*>(p);
    for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ULL; }
  };
  mix(s->se3_off.data(), s->se3_off.size() * sizeof(int));
  mix(s->se2_off.data(), s->se2_off.size() * sizeof(int));
  mix(s->vec_off.data(), s->vec_off.size() * sizeof(int));
  mix(s->so3_off.data(), s->so3_off.size() * sizeof(int));
  mix(s->tile_mask.data(), s->tile_mask.size());
  mix(&s->n_pad, sizeof(int));
  *hash = h;
  return BSLAM_OK;
}

int bslam_peer_region(bslam_solver* s, void** dev_ptr, size_t* n_bytes, uint8_t* ipc_handle) {
  NEED(s && s->finalized, "bslam_peer_region: solver not finalized");
  CU(cudaSetDevice(s->device));
  if (!s->plan_valid) { int rc = build_chol_plan(s); if (rc) return rc; }
  CU(cudaStreamSynchronize(s->stream));                      // the region is zeroed before anybody can map it
  if (dev_ptr) *dev_ptr = s->d_pack.p;
  if (n_bytes) *n_bytes = s->d_pack.n * sizeof(double);
  if (ipc_handle) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, s->d_pack.p));
    std::memcpy(ipc_handle, &h, sizeof h);
  }
  return BSLAM_OK;
}

int bslam_peer_connect(bslam_solver* s, int world, int rank, const uint8_t* ipc_handles, void* const* dev_ptrs) {
  NEED(s && s->finalized, "bslam_peer_connect: solver not finalized");
  NEED(world >= 1 && world <= bs::kMaxPeers && rank >= 0 && rank < world, "bslam_peer_connect: world %d / rank %d invalid (max %d ranks)",
       world, rank, bs::kMaxPeers);
  NEED(world == 1 || ipc_handles || dev_ptrs, "bslam_peer_connect: neither IPC handles nor device pointers given");
  CU(cudaSetDevice(s->device));
  if (!s->plan_valid) { int rc = build_chol_plan(s); if (rc) return rc; }
  CU(cudaStreamSynchronize(s->stream));
  for (int r = 0; r < world; ++r) {
    s->mc_region = nullptr;
    if (r == rank) { s->peer_region[r] = s->d_pack.p; continue; }
    if (dev_ptrs && dev_ptrs[r]) { s->peer_region[r] = static_cast<double*>(dev_ptrs[r]); continue; }
    NEED(ipc_handles, "bslam_peer_connect: no handle for rank %d", r);
    cudaIpcMemHandle_t h;
    std::memcpy(&h, ipc_handles + 64 * (size_t)r, sizeof h);
    void* p = nullptr;
    CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    s->peer_region[r] = static_cast<double*>(p);
    s->peer_opened[r] = true;
  }
  if (world > 1) preload_iteration_kernels(s);
  CU(s->d_peer_ctl.alloc(4));
  CU(cudaMemsetAsync(s->d_peer_ctl.p, 0, 4 * sizeof(long long), s->stream));
  CU(cudaStreamSynchronize(s->stream));
  s->world = world;
  s->shard_rank = rank;
  drop_graph(s);
  return BSLAM_OK;
}

int bslam_peer_connect_symmetric(bslam_solver* s, int world, int rank, void* const* region_ptrs, void* multicast_ptr, size_t n_bytes) {
  NEED(s && s->finalized, "bslam_peer_connect_symmetric: solver not finalized");
  NEED(world >= 2 && world <= bs::kMaxPeers && rank >= 0 && rank < world && region_ptrs, "bslam_peer_connect_symmetric: bad arguments");
  CU(cudaSetDevice(s->device));
  if (!s->plan_valid) { int rc = build_chol_plan(s); if (rc) return rc; }
  NEED(n_bytes >= s->d_pack.n * sizeof(double), "bslam_peer_connect_symmetric: regions of %zu bytes, need %zu", n_bytes,
       s->d_pack.n * sizeof(double));
  for (int r = 0; r < world; ++r) {
    NEED(region_ptrs[r], "bslam_peer_connect_symmetric: no region for rank %d", r);
    s->peer_region[r] = static_cast<double*>(region_ptrs[r]);       // the own region too: the caller's symmetric allocation
  }
  s->mc_region = static_cast<double*>(multicast_ptr);
  preload_iteration_kernels(s);
  CU(s->d_peer_ctl.alloc(4));
  CU(cudaMemsetAsync(s->d_peer_ctl.p, 0, 4 * sizeof(long long), s->stream));
  CU(cudaStreamSynchronize(s->stream));
  s->world = world;
  s->shard_rank = rank;
  drop_graph(s);
  return BSLAM_OK;
}

// Tiles (slots of the packed payload) THIS handle's residual blocks can write: its landmarks' co-visibility tiles, and on
// shard 0 the blocks that are not sharded.  The ranks exchange these once; bslam_peer_set_contributors then tells the
// Cholesky kernel which ranks to read for every tile.
int bslam_peer_local_slots(bslam_solver* s, uint8_t* flags, size_t n) {
  NEED(s && s->finalized && flags, "bslam_peer_local_slots: bad arguments");
  CU(cudaSetDevice(s->device));
  if (!s->plan_valid) { int rc = build_chol_plan(s); if (rc) return rc; }
  NEED(n == (size_t)s->n_nz_tiles, "bslam_peer_local_slots: expected %d slots, got %zu", s->n_nz_tiles, n);
  const int nt = s->nblk;
  std::vector<uint8_t> local((size_t)nt * nt, 0);
  std::vector<int> tiles;
  auto mark = [&]() {
    for (int a2 : tiles)
      for (int b2 : tiles)
        if (a2 >= b2) local[(size_t)a2 * nt + b2] = 1;
  };
  for (int q = 0; q < s->n_lm; ++q) {
    tiles.clear();
    for (int k = s->h_lm_start[q]; k < s->h_lm_start[q + 1]; ++k) tiles_of(s->se3_off[s->h_opose_lm[k]], 6, tiles);
    mark();
  }
  if (s->shard_rank == 0) {
    for (int t = 0; t < nt; ++t) local[(size_t)t * nt + t] = 1;        // priors, photometric / motion-only blocks: diagonal tiles
    for (auto* b : s->edges) {
      if (!b->binary) continue;
      const std::vector<int>& off = b->group == 3 ? s->se3_off : s->se2_off;
      const int dof = b->group == 3 ? 6 : 3;
      for (int e = 0; e < b->n; ++e) {
        tiles.clear();
        tiles_of(off[b->i1[e]], dof, tiles);
        tiles_of(off[b->i2[e]], dof, tiles);
        mark();
      }
    }
    for (auto* b : s->photos) {
      if (b->rot_idx < 0) continue;
      tiles.clear();
      tiles_of(s->so3_off[b->rot_idx], 3, tiles);
      tiles_of(s->vec_off[b->vec_idx], 3, tiles);
      mark();
    }
    for (int b = 0; b < s->dn_blocks; ++b) {
      tiles.clear();
      for (int c = s->dn_col_ptr[b]; c < s->dn_col_ptr[b + 1]; ++c) tiles_of(s->dn_col_index[c], 1, tiles);
      mark();
    }
  }
  int slot = 0;
  for (int i = 0; i < nt; ++i)
    for (int j = 0; j <= i; ++j)
      if (s->tile_mask[(size_t)i * nt + j]) flags[slot++] = local[(size_t)i * nt + j];
  return BSLAM_OK;
}

int bslam_peer_set_contributors(bslam_solver* s, const uint8_t* masks, size_t n) {
  NEED(s && s->finalized && masks, "bslam_peer_set_contributors: bad arguments");
  NEED(s->plan_valid && n == (size_t)s->n_nz_tiles, "bslam_peer_set_contributors: expected %d slots, got %zu", s->n_nz_tiles, n);
  CU(cudaSetDevice(s->device));
  const int nt = s->nblk;
  std::vector<int> diag_slot(nt, -1);
  {
    int slot = 0;
    for (int i = 0; i < nt; ++i)
      for (int j = 0; j <= i; ++j)
        if (s->tile_mask[(size_t)i * nt + j]) { if (i == j) diag_slot[i] = slot; ++slot; }
  }
  for (bs::CholTask& t : s->h_tasks) {
    if (t.slot >= 0) t.ranks = masks[t.slot];
    else if (t.slot == -2) t.ranks = diag_slot[t.j] >= 0 ? masks[diag_slot[t.j]] : 0xff;     // rhs of the poses of tile column j
    else t.ranks = 0;
  }
  CU(cudaStreamSynchronize(s->stream));
  CU(upload(s->d_tasks, s->h_tasks, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  drop_graph(s);
  return BSLAM_OK;
}

int bslam_peer_barrier(bslam_solver* s) {
  NEED(s && s->finalized, "bslam_peer_barrier: solver not finalized");
  if (s->world <= 1) return BSLAM_OK;
  CU(cudaSetDevice(s->device));
  LAUNCH(s, bs::peer_scalar_exchange_kernel, 1, 32, 0, s->scalars(), peer_ctx(s), 0);
  CU(cudaGetLastError());
  return BSLAM_OK;
}

int bslam_iterate_host(bslam_solver* s, double lambda, int eval_new_cost, const double* Rt_in, const double* xyz_in,
                       double* Rt_out, double* xyz_out, double* cost_lin, double* cost_new, double* dx_norm) {
  NEED(s && s->finalized, "bslam_iterate_host: solver not finalized");
  NEED(lambda >= 0.0, "bslam_iterate_host: lambda must be >= 0");
  CU(cudaSetDevice(s->device));
  int rc;
  // host parameters -> device (stream-ordered; the point table is permuted to the internal order on the device)
  if (Rt_in && s->n_se3 > 0)
    CU(cudaMemcpyAsync(s->d_se3.p, Rt_in, (size_t)s->n_se3 * 12 * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  if (xyz_in && s->n_pt > 0) {
    CU(cudaMemcpyAsync(s->d_stage.p, xyz_in, (size_t)s->n_pt * 3 * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    LAUNCH(s, permute_rows_kernel, cdiv(3LL * s->n_pt, 256), 256, 0, s->n_pt, 3, s->d_stage.p, s->d_pts.p, s->d_pt_perm.p, 1);
  }
  if ((rc = iterate_enqueue(s, lambda, eval_new_cost))) return rc;
  // updated parameters -> host, then ONE synchronisation for the whole step
  if (Rt_out && s->n_se3 > 0)
    CU(cudaMemcpyAsync(Rt_out, s->d_se3.p, (size_t)s->n_se3 * 12 * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  if (xyz_out && s->n_pt > 0) {
    LAUNCH(s, permute_rows_kernel, cdiv(3LL * s->n_pt, 256), 256, 0, s->n_pt, 3, s->d_pts.p, s->d_stage.p, s->d_pt_perm.p, 0);
    CU(cudaMemcpyAsync(xyz_out, s->d_stage.p, (size_t)s->n_pt * 3 * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  }
  CU(cudaGetLastError());
  return iterate_finish(s, cost_lin, cost_new, dx_norm);
}

int bslam_reduced_buffer(bslam_solver* s, void** dev_ptr, size_t* n_doubles, void** scalars_dev_ptr, int32_t* n_pad) {
  NEED(s && s->finalized, "bslam_reduced_buffer: solver not finalized");
  if (dev_ptr) *dev_ptr = s->d_red.p;
  if (n_doubles) *n_doubles = s->red_len();
  if (scalars_dev_ptr) *scalars_dev_ptr = s->scalars();
  if (n_pad) *n_pad = s->n_pad;
  return BSLAM_OK;
}

int bslam_packed_buffer(bslam_solver* s, void** dev_ptr, size_t* n_doubles) {
  NEED(s && s->finalized, "bslam_packed_buffer: solver not finalized");
  CU(cudaSetDevice(s->device));
  if (!s->plan_valid) { int rc = build_chol_plan(s); if (rc) return rc; }
  if (dev_ptr) *dev_ptr = s->d_pack.p;
  if (n_doubles) *n_doubles = s->pack_len;        // the payload [tiles | rhs | scalars]; the exchange tail behind it is private
  return BSLAM_OK;
}

static int do_pack(bslam_solver* s, int unpack);

int bslam_pack_reduced(bslam_solver* s, int unpack) {
  NEED(s && s->finalized, "bslam_pack_reduced: solver not finalized");
  CU(cudaSetDevice(s->device));
  if (!s->plan_valid) { int rc = build_chol_plan(s); if (rc) return rc; }
  return do_pack(s, unpack);
}

static int do_pack(bslam_solver* s, int unpack) {
  const size_t tiles = (size_t)s->n_nz_tiles * bs::kNB * bs::kNB;
  if (s->n_nz_tiles > 0)
    LAUNCH(s, bs::pack_tiles_kernel, s->n_nz_tiles, 256, 0, s->S(), s->n_pad, s->nblk, s->d_nz_tiles.p, s->d_pack.p, unpack);
  const size_t tail = (size_t)s->n_pad + BSLAM_N_SCALARS;      // rhs | scalars are contiguous in both buffers
  if (unpack) CU(cudaMemcpyAsync(s->rhs(), s->d_pack.p + tiles, tail * sizeof(double), cudaMemcpyDeviceToDevice, s->stream));
  else CU(cudaMemcpyAsync(s->d_pack.p + tiles, s->rhs(), tail * sizeof(double), cudaMemcpyDeviceToDevice, s->stream));
  CU(cudaGetLastError());
  return BSLAM_OK;
}

int bslam_iterate_pre(bslam_solver* s, double lambda) {
  NEED(s && s->finalized, "bslam_iterate_pre: solver not finalized");
  NEED(lambda >= 0.0, "bslam_iterate_pre: lambda must be >= 0");
  CU(cudaSetDevice(s->device));
  int rc;
  if ((rc = prepare_iterate(s))) return rc;
  auto body = [&]() {
    int r = do_linearize(s, true);
    if (!r) r = do_reduce(s, lambda, true);
    if (!r) r = do_pack(s, 0);
    return r;
  };
  if (s->use_graph && !s->timing && s->dn_blocks == 0) {
    if (s->graph_pre && s->graph_pre_lambda != lambda) { cudaGraphExecDestroy(s->graph_pre); s->graph_pre = nullptr; }
    s->graph_pre_lambda = lambda;
    return run_graphed(s, &s->graph_pre, &s->graph_pre_launches, body);
  }
  return body();
}

int bslam_iterate_post(bslam_solver* s, int eval_new_cost) {
  NEED(s && s->finalized, "bslam_iterate_post: solver not finalized");
  CU(cudaSetDevice(s->device));
  int rc;
  if ((rc = prepare_iterate(s))) return rc;
  auto body = [&]() {
    int r = do_pack(s, 1);
    if (!r) r = do_solve_reduced(s, fuse_retract(s, true));
    if (!r) r = do_retract(s, eval_new_cost, true, fuse_retract(s, true));
    return r;
  };
  if (s->use_graph && !s->timing && s->dn_blocks == 0 && s->d_trace.p == nullptr) {
    if (s->graph_post && s->graph_post_eval != eval_new_cost) { cudaGraphExecDestroy(s->graph_post); s->graph_post = nullptr; }
    s->graph_post_eval = eval_new_cost;
    return run_graphed(s, &s->graph_post, &s->graph_post_launches, body);
  }
  return body();
}

int bslam_set_shard(bslam_solver* s, int rank) {
  NEED(s, "NULL solver");
  s->shard_rank = rank;
  return BSLAM_OK;
}

int bslam_tile_structure(bslam_solver* s, uint8_t* mask, size_t n, int set) {
  NEED(s && s->finalized && mask, "bslam_tile_structure: bad arguments");
  NEED(n == s->tile_mask.size(), "bslam_tile_structure: expected %zu bytes, got %zu", s->tile_mask.size(), n);
  if (set) {
    NEED(s->world == 1, "bslam_tile_structure: the structure is frozen after bslam_peer_connect");
    for (size_t i = 0; i < n; ++i) s->tile_mask[i] = s->tile_mask[i] || mask[i];
    s->plan_valid = false;
    drop_graph(s);
  } else {
    std::memcpy(mask, s->tile_mask.data(), n);
  }
  return BSLAM_OK;
}

void* bslam_stream(bslam_solver* s) { return s ? (void*)s->stream : nullptr; }

int bslam_snapshot(bslam_solver* s) {
  NEED(s && s->finalized, "bslam_snapshot: solver not finalized");
  CU(cudaSetDevice(s->device));
  auto cp = [&](DevBuf<double>& dst, DevBuf<double>& src) {
    return src.n ? cudaMemcpyAsync(dst.p, src.p, src.n * sizeof(double), cudaMemcpyDeviceToDevice, s->stream) : cudaSuccess;
  };
  CU(cp(s->b_se3, s->d_se3)); CU(cp(s->b_se2, s->d_se2)); CU(cp(s->b_pts, s->d_pts)); CU(cp(s->b_vec, s->d_vec));
  CU(cp(s->b_so3, s->d_so3));
  return BSLAM_OK;
}

int bslam_restore(bslam_solver* s) {
  NEED(s && s->finalized, "bslam_restore: solver not finalized");
  CU(cudaSetDevice(s->device));
  auto cp = [&](DevBuf<double>& dst, DevBuf<double>& src) {
    return src.n ? cudaMemcpyAsync(dst.p, src.p, src.n * sizeof(double), cudaMemcpyDeviceToDevice, s->stream) : cudaSuccess;
  };
  CU(cp(s->d_se3, s->b_se3)); CU(cp(s->d_se2, s->b_se2)); CU(cp(s->d_pts, s->b_pts)); CU(cp(s->d_vec, s->b_vec));
  CU(cp(s->d_so3, s->b_so3));
  return BSLAM_OK;
}

// ---------------------------------------------------------------- inspection

int bslam_get_update(bslam_solver* s, double* dx) {
  NEED(s && s->finalized && dx, "bslam_get_update: bad arguments");
  CU(cudaSetDevice(s->device));
  if (s->n_red) CU(cudaMemcpyAsync(dx, s->d_dx.p, s->n_red * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  if (s->n_lm)
    CU(cudaMemcpyAsync(dx + s->n_red, s->d_dx.p + s->n_pad, 3 * (size_t)s->n_lm * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  return BSLAM_OK;
}

int bslam_get_reduced_system(bslam_solver* s, double* Sout, double* rhs) {
  NEED(s && s->finalized, "bslam_get_reduced_system: solver not finalized");
  CU(cudaSetDevice(s->device));
  const int n = s->n_red, ld = s->n_pad;
  std::vector<double> tmp((size_t)ld * ld);
  CU(cudaMemcpyAsync(tmp.data(), s->S(), tmp.size() * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  if (rhs && n) CU(cudaMemcpyAsync(rhs, s->rhs(), n * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  if (Sout)
    for (int i = 0; i < n; ++i)
      for (int j = 0; j <= i; ++j) Sout[(size_t)i * n + j] = Sout[(size_t)j * n + i] = tmp[(size_t)i * ld + j];
  return BSLAM_OK;
}

int bslam_get_normal_equations(bslam_solver* s, double* H, double* b) {
  NEED(s && s->finalized && H && b, "bslam_get_normal_equations: bad arguments");
  NEED(s->dim <= 20000, "bslam_get_normal_equations: D = %d too large for a dense export", s->dim);
  NEED(s->n_obs == 0 || s->d_W.p, "bslam_get_normal_equations: call bslam_linearize first");
  CU(cudaSetDevice(s->device));
  const int D = s->dim, n = s->n_red, ld = s->n_pad, N = s->n_obs;
  std::vector<double> Sd((size_t)ld * ld), W(bs::w_alloc_len(N)), Vg(9 * (size_t)s->n_lm);
  std::vector<int> opose(N), opt(N);
  CU(cudaMemcpyAsync(Sd.data(), s->S(), Sd.size() * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  if (n) CU(cudaMemcpyAsync(b, s->rhs(), n * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  if (N) {
    CU(cudaMemcpyAsync(W.data(), s->d_W.p, W.size() * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaMemcpyAsync(opose.data(), s->d_opose.p, N * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaMemcpyAsync(opt.data(), s->d_opt.p, N * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
  }
  if (s->n_lm) CU(cudaMemcpyAsync(Vg.data(), s->d_Vg.p, Vg.size() * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  std::fill(H, H + (size_t)D * D, 0.0);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j <= i; ++j) H[(size_t)i * D + j] = H[(size_t)j * D + i] = Sd[(size_t)i * ld + j];
  static const int vi[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
  for (int q = 0; q < s->n_lm; ++q) {
    const int o = n + 3 * q;
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) H[(size_t)(o + r) * D + o + c] = Vg[9 * (size_t)q + vi[r][c]];
      b[o + r] = Vg[9 * (size_t)q + 6 + r];
    }
  }
  for (int k = 0; k < N; ++k) {
    const int po = s->se3_off[opose[k]];
    if (po < 0 || opt[k] >= s->n_lm) continue;
    const int o = n + 3 * opt[k];
    for (int r = 0; r < 6; ++r)
      for (int c = 0; c < 3; ++c) {
        H[(size_t)(po + r) * D + o + c] += W[bs::w_index(k, 3 * r + c)];
        H[(size_t)(o + c) * D + po + r] += W[bs::w_index(k, 3 * r + c)];
      }
  }
  return BSLAM_OK;
}

int bslam_covariance(bslam_solver* s, double* cov) {
  NEED(s && s->finalized && cov, "bslam_covariance: bad arguments");
  NEED(s->dim <= 8192, "bslam_covariance: D = %d too large for a dense covariance (limit 8192)", s->dim);
  CU(cudaSetDevice(s->device));
  int rc;
  if ((rc = do_linearize(s, false))) return rc;
  if ((rc = do_reduce(s, 0.0, false))) return rc;
  if ((rc = do_solve_reduced(s))) return rc;          // factor L (in S), diagonal inverses (Linv)
  const size_t D = (size_t)s->dim;
  const int nt = s->nblk, ld = s->n_pad;
  DevBuf<double> d_cov, d_G;
  CU(d_cov.alloc(D * D));
  CU(d_G.alloc((size_t)ld * ld));
  CU(cudaMemsetAsync(d_cov.p, 0, D * D * sizeof(double), s->stream));
  CU(cudaMemsetAsync(d_G.p, 0, (size_t)ld * ld * sizeof(double), s->stream));
  const size_t smem = 2 * bs::kNB * bs::kLd * sizeof(double);
  CU(cudaFuncSetAttribute(bs::cov_linv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CU(cudaFuncSetAttribute(bs::cov_gtg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  LAUNCH(s, bs::cov_linv_kernel, nt, bs::kCholThreads, smem, s->S(), ld, s->d_Linv.p, nt, s->d_fill_mask.p, d_G.p);
  LAUNCH(s, bs::cov_gtg_kernel, dim3(nt, nt), bs::kCholThreads, smem, d_G.p, ld, nt, d_cov.p, D);
  if (s->n_lm > 0) {
    bs::CovLmArgs a;
    a.n_lm = s->n_lm; a.n_obs = s->n_obs; a.n_pad = s->n_pad; a.D = D;
    a.obs_pose = s->d_opose.p; a.lm_start = s->d_lm_start.p; a.lm_obs = s->d_lm_obs.p; a.pose_off = s->d_se3_off.p;
    a.W = s->d_W.p; a.Vinv = s->d_Vinv.p; a.cov = d_cov.p;
    LAUNCH(s, bs::cov_lm_pose_kernel, dim3(cdiv(s->n_pad, 128), 3 * s->n_lm), 128, 0, a);
    LAUNCH(s, bs::cov_lm_lm_kernel, dim3(cdiv(s->n_lm, 128), 3 * s->n_lm), 128, 0, a);
  }
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(cov, d_cov.p, D * D * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  return BSLAM_OK;
}

int bslam_debug_chol_trace(bslam_solver* s, int64_t* out, int max_tasks, int* n_tasks) {
  NEED(s && s->finalized && out && n_tasks, "bslam_debug_chol_trace: bad arguments");
  CU(cudaSetDevice(s->device));
  if (!s->plan_valid) { int rc = build_chol_plan(s); if (rc) return rc; }
  const int n = s->n_tile_tasks + s->nblk;
  *n_tasks = n;
  NEED(max_tasks >= n, "bslam_debug_chol_trace: need room for %d tasks", n);
  CU(s->d_trace.alloc(4 * (size_t)n));
  int rc = do_solve_reduced(s);
  if (rc) return rc;
  std::vector<long long> h(4 * (size_t)n);
  CU(cudaMemcpyAsync(h.data(), s->d_trace.p, h.size() * sizeof(long long), cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  s->d_trace.release();
  for (int t = 0; t < n; ++t) {
    out[6 * t + 0] = t < s->n_tile_tasks ? s->h_tasks[t].i : -1;
    out[6 * t + 1] = t < s->n_tile_tasks ? s->h_tasks[t].j : s->nblk - 1 - (t - s->n_tile_tasks);
    for (int k = 0; k < 4; ++k) out[6 * t + 2 + k] = h[4 * (size_t)t + k];
  }
  return BSLAM_OK;
}

int bslam_enable_timing(bslam_solver* s, int on) {
  NEED(s, "NULL solver");
  s->timing = on != 0;
  return BSLAM_OK;
}

int bslam_get_timings(bslam_solver* s, double* ms) {
  NEED(s && ms, "bslam_get_timings: bad arguments");
  std::memcpy(ms, s->timings, sizeof s->timings);
  return BSLAM_OK;
}

int64_t bslam_launch_count(const bslam_solver* s) { return s ? s->launches : 0; }

// ---------------------------------------------------------------- stand-alone device routines (no solver handle)

#define CUG(call)                                                                                   \
  do {                                                                                              \
    cudaError_t e__ = (call);                                                                       \
    if (e__ != cudaSuccess) return fail(nullptr, BSLAM_E_CUDA, "%s failed: %s", #call, cudaGetErrorString(e__)); \
  } while (0)

int bslam_ransac(int device, int n_hyp, int n_min, const int32_t* idx, const double* T_21_in, int n_pts, const double* pts_1,
                 const double* pts_2, const double* obs_2, const double intr[5], double thresh, double* T_21_out,
                 int32_t* counts, int32_t* best, uint8_t* best_mask) {
  bslam_solver* s = nullptr;
  if (n_hyp <= 0 || n_pts <= 0 || !pts_1 || !obs_2 || !intr || (!idx && !T_21_in) || (idx && (!pts_2 || n_min < 3)))
    return fail(s, BSLAM_E_INVALID, "bslam_ransac: bad arguments");
  if (idx)
    for (size_t k = 0; k < (size_t)n_hyp * n_min; ++k)
      if (idx[k] < 0 || idx[k] >= n_pts) return fail(s, BSLAM_E_INVALID, "bslam_ransac: sample index %d outside [0,%d)", idx[k], n_pts);
  CUG(cudaSetDevice(device));
  DevBuf<double> d_p1, d_p2, d_o2, d_T;
  DevBuf<int> d_idx, d_cnt, d_best;
  DevBuf<unsigned char> d_mask;
  CUG(d_p1.alloc(3 * (size_t)n_pts)); CUG(d_o2.alloc(3 * (size_t)n_pts)); CUG(d_T.alloc(16 * (size_t)n_hyp));
  CUG(d_cnt.alloc(n_hyp)); CUG(d_best.alloc(2)); CUG(d_mask.alloc(n_pts));
  CUG(cudaMemcpy(d_p1.p, pts_1, 3 * (size_t)n_pts * sizeof(double), cudaMemcpyHostToDevice));
  CUG(cudaMemcpy(d_o2.p, obs_2, 3 * (size_t)n_pts * sizeof(double), cudaMemcpyHostToDevice));
  if (idx) {
    CUG(d_p2.alloc(3 * (size_t)n_pts)); CUG(d_idx.alloc((size_t)n_hyp * n_min));
    CUG(cudaMemcpy(d_p2.p, pts_2, 3 * (size_t)n_pts * sizeof(double), cudaMemcpyHostToDevice));
    CUG(cudaMemcpy(d_idx.p, idx, (size_t)n_hyp * n_min * sizeof(int), cudaMemcpyHostToDevice));
    bs::ransac_transform_kernel<<<cdiv(n_hyp, 128), 128>>>(n_hyp, n_min, d_idx.p, d_p1.p, d_p2.p, d_T.p);
  } else {
    CUG(cudaMemcpy(d_T.p, T_21_in, 16 * (size_t)n_hyp * sizeof(double), cudaMemcpyHostToDevice));
  }
  bs::ransac_count_kernel<<<n_hyp, 256>>>(n_pts, d_T.p, d_p1.p, d_o2.p, intr[0], intr[1], intr[2], intr[3], intr[4], thresh, d_cnt.p);
  bs::ransac_best_kernel<<<1, 256>>>(n_hyp, n_pts, d_cnt.p, d_T.p, d_p1.p, d_o2.p, intr[0], intr[1], intr[2], intr[3], intr[4], thresh,
                                     d_best.p, d_mask.p);
  CUG(cudaGetLastError());
  if (T_21_out) CUG(cudaMemcpy(T_21_out, d_T.p, 16 * (size_t)n_hyp * sizeof(double), cudaMemcpyDeviceToHost));
  if (counts) CUG(cudaMemcpy(counts, d_cnt.p, n_hyp * sizeof(int), cudaMemcpyDeviceToHost));
  if (best) CUG(cudaMemcpy(best, d_best.p, 2 * sizeof(int), cudaMemcpyDeviceToHost));
  if (best_mask) CUG(cudaMemcpy(best_mask, d_mask.p, n_pts, cudaMemcpyDeviceToHost));
  CUG(cudaDeviceSynchronize());
  return BSLAM_OK;
}

int bslam_image_pyramid(int device, const uint8_t* image, int width, int height, int levels, double* im_out, double* gx_out,
                        double* gy_out) {
  bslam_solver* s = nullptr;
  if (!image || width <= 0 || height <= 0 || levels <= 0 || !im_out) return fail(s, BSLAM_E_INVALID, "bslam_image_pyramid: bad arguments");
  CUG(cudaSetDevice(device));
  size_t total = 0;
  { int w = width, h = height; for (int l = 0; l < levels; ++l) { total += (size_t)w * h; w = (w + 1) / 2; h = (h + 1) / 2; } }
  DevBuf<unsigned char> d_a, d_b;
  DevBuf<double> d_im, d_gx, d_gy;
  CUG(d_a.alloc((size_t)width * height)); CUG(d_b.alloc((size_t)width * height));
  CUG(d_im.alloc(total)); CUG(d_gx.alloc(total)); CUG(d_gy.alloc(total));
  CUG(cudaMemcpy(d_a.p, image, (size_t)width * height, cudaMemcpyHostToDevice));
  unsigned char *cur = d_a.p, *nxt = d_b.p;
  size_t off = 0;
  int w = width, h = height;
  for (int l = 0; l < levels; ++l) {
    const size_t n = (size_t)w * h;
    bs::u8_to_unit_kernel<<<cdiv((long long)n, 256), 256>>>(cur, n, d_im.p + off);
    if (gx_out && gy_out)
      bs::sobel_half_kernel<<<dim3(cdiv(w, 16), cdiv(h, 16)), 256>>>(d_im.p + off, w, h, d_gx.p + off, d_gy.p + off);
    off += n;
    if (l + 1 < levels) {
      const int wo = (w + 1) / 2, ho = (h + 1) / 2;
      bs::pyr_down_u8_kernel<<<dim3(cdiv(wo, 16), cdiv(ho, 16)), 256>>>(cur, w, h, nxt, wo, ho);
      std::swap(cur, nxt);
      w = wo; h = ho;
    }
  }
  CUG(cudaGetLastError());
  CUG(cudaMemcpy(im_out, d_im.p, total * sizeof(double), cudaMemcpyDeviceToHost));
  if (gx_out && gy_out) {
    CUG(cudaMemcpy(gx_out, d_gx.p, total * sizeof(double), cudaMemcpyDeviceToHost));
    CUG(cudaMemcpy(gy_out, d_gy.p, total * sizeof(double), cudaMemcpyDeviceToHost));
  }
  CUG(cudaDeviceSynchronize());
  return BSLAM_OK;
}

int bslam_subsample_pyramid(int device, const double* map, int width, int height, int levels, double scale_per_level, double* out) {
  bslam_solver* s = nullptr;
  if (!map || width <= 0 || height <= 0 || levels <= 0 || !out) return fail(s, BSLAM_E_INVALID, "bslam_subsample_pyramid: bad arguments");
  CUG(cudaSetDevice(device));
  size_t total = 0;
  { int w = width, h = height; for (int l = 0; l < levels; ++l) { total += (size_t)w * h; w = (w + 1) / 2; h = (h + 1) / 2; } }
  // keyframes.py:100-112: the UNSCALED map is sub-sampled level after level ([0::2, 0::2]); level l is stored times scale^l
  DevBuf<double> d_raw, d_out;
  CUG(d_raw.alloc(total)); CUG(d_out.alloc(total));
  CUG(cudaMemcpy(d_raw.p, map, (size_t)width * height * sizeof(double), cudaMemcpyHostToDevice));
  CUG(cudaMemcpy(d_out.p, d_raw.p, (size_t)width * height * sizeof(double), cudaMemcpyDeviceToDevice));
  size_t off = 0;
  int w = width, h = height;
  double scale = 1.0;
  for (int l = 1; l < levels; ++l) {
    const int wo = (w + 1) / 2, ho = (h + 1) / 2;
    const size_t n = (size_t)w * h;
    scale *= scale_per_level;
    const dim3 grid(cdiv(wo, 16), cdiv(ho, 16));
    bs::subsample2_kernel<<<grid, 256>>>(d_raw.p + off, w, h, d_raw.p + off + n, wo, ho, 1.0);
    bs::subsample2_kernel<<<grid, 256>>>(d_raw.p + off, w, h, d_out.p + off + n, wo, ho, scale);
    off += n;
    w = wo; h = ho;
  }
  CUG(cudaGetLastError());
  CUG(cudaMemcpy(out, d_out.p, total * sizeof(double), cudaMemcpyDeviceToHost));
  CUG(cudaDeviceSynchronize());
  return BSLAM_OK;
}

}  // extern "C"

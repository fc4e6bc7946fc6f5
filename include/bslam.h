/* bslam.h -- C ABI of libbslam.so: the B200-native Gauss-Newton / LM inner loop
 * behind pyslam's `Problem.solve()`.
 *
 * The reference (utiasSTARS/pyslam) has NO FFI of its own: its boundary is the
 * Python `Problem` class plus three duck-typed protocols.  Each
 * entry point below therefore cites the reference *Python* code whose work it
 * replaces; pyslam_b200/problem.py (the drop-in `Problem`) is the only caller
 * and binds these with ctypes (pyslam_b200/engine.py; binding stub for a
 * reference maintainer in INTEGRATION.md).
 *
 * Rules of the ABI
 *   - every function returns 0 on success, <0 on error (BSLAM_E_*); the message
 *     is available from bslam_last_error(); nothing throws across the boundary;
 *   - the caller owns every host buffer; the library copies during the call and
 *     never retains a host pointer;  the library owns all device memory;
 *   - one CUDA stream per handle; a handle is not thread-safe, distinct handles
 *     are independent;
 *   - all floating point data is IEEE double; indices are int32;
 *   - matrices are row-major.  An SE(3) pose is 12 doubles [R(3x3) | t(3)]
 *     (R first, row-major, then t); an SE(2) pose is 6 doubles [R(2x2) | t(2)].
 *     Tangent vectors are ordered [rho; phi] and perturbations are applied on
 *     the left, T <- exp(xi) T, as liegroups/pyslam do.
 *
 * Life cycle:  create -> set_* parameter tables -> add_* blocks -> finalize ->
 *              { iterate | linearize/solve/retract | eval_cost }* -> get_* -> destroy
 */
#ifndef BSLAM_H_
#define BSLAM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bslam_solver bslam_solver;

#if defined(__GNUC__)
#define BSLAM_API __attribute__((visibility("default")))
#else
#define BSLAM_API
#endif

/* ---- status codes ------------------------------------------------------ */
#define BSLAM_OK              0
#define BSLAM_E_INVALID      -1   /* bad argument / call order            */
#define BSLAM_E_CUDA         -2   /* CUDA runtime error (see last_error)  */
#define BSLAM_E_STRUCTURE    -3   /* problem structure unsupported        */
#define BSLAM_E_NUMERIC      -4   /* reduced system not positive definite */

/* ---- pose groups --------------------------------------------------------- */
#define BSLAM_SE2 2
#define BSLAM_SE3 3

/* ---- robust losses: pyslam/losses.py:8-214 (element-wise IRLS) -- */
#define BSLAM_LOSS_L2     0
#define BSLAM_LOSS_L1     1
#define BSLAM_LOSS_CAUCHY 2
#define BSLAM_LOSS_HUBER  3
#define BSLAM_LOSS_TUKEY  4
#define BSLAM_LOSS_TDIST  5

/* ---- scalar slots written by bslam_iterate / read by bslam_get_scalars ---- */
#define BSLAM_S_COST_LIN   0   /* sum rho(r) at the linearisation point (blocks with >=1 variable param) */
#define BSLAM_S_COST_NEW   1   /* sum rho(r) over ALL blocks at x [+] dx                                 */
#define BSLAM_S_DX_NORM2   2   /* ||dx||^2 over the whole update vector                                   */
#define BSLAM_S_CHOL_FAIL  3   /* >0 if a non-positive pivot was met                                       */
#define BSLAM_S_COST_EVAL  4   /* result of bslam_eval_cost                                                */
#define BSLAM_S_PEER_TIMEOUT 5 /* >0 if a rendezvous of the sharded iteration timed out (bslam_peer_connect) */
#define BSLAM_N_SCALARS   16

/* ---- timing slots (milliseconds, CUDA events on the handle's stream) ------ */
#define BSLAM_T_LINEARIZE  0   /* zero + all linearisation kernels                  */
#define BSLAM_T_REPROJ     1   /* the reprojection linearisation kernel alone       */
#define BSLAM_T_SCHUR      2   /* landmark inverse + Schur complement               */
#define BSLAM_T_CHOLESKY   3   /* dense factorisation of the reduced system         */
#define BSLAM_T_TRSV       4   /* forward / backward substitution                   */
#define BSLAM_T_BACKSUB    5   /* landmark back-substitution                        */
#define BSLAM_T_RETRACT    6   /* exp / retract + ||dx||                            */
#define BSLAM_T_COST       7   /* cost at the new point                             */
#define BSLAM_T_TOTAL      8
#define BSLAM_T_FUSED      9   /* fused_panel_kernel alone (linearise + eliminate the panels) */
#define BSLAM_T_PEER_PUBLISH 10 /* sharded iteration: pack + rendezvous of the partial reduced systems  */
#define BSLAM_T_PEER_SCALARS 11 /* sharded iteration: scalar exchange + rendezvous at the end            */
#define BSLAM_N_TIMINGS   16

/* ---- life cycle ------------------------------------------------------------ */

/* Library/ABI version (major*10000 + minor*100 + patch). */
BSLAM_API int bslam_version(void);
/* Edge of the square tiles the reduced camera system is stored and factorised in (32): the unit of
 * bslam_tile_structure masks and of the packed multi-GPU payload. */
BSLAM_API int bslam_tile_edge(void);

/* Create a solver bound to CUDA device `device`.  Replaces `Problem.__init__`
 * (pyslam/problem.py:43-69).  Fails with BSLAM_E_CUDA when no usable GPU exists:
 * there is no CPU fallback. */
BSLAM_API int bslam_create(bslam_solver** out, int device);
BSLAM_API void bslam_destroy(bslam_solver* s);

/* Message of the most recent error on `s` (or of a failed bslam_create when s
 * is NULL).  The pointer stays valid until the next call on the same handle. */
BSLAM_API const char* bslam_last_error(const bslam_solver* s);

/* ---- parameter tables: `Problem.initialize_params` +
 *      `set_parameters_constant/variable` (pyslam/problem.py:83-108) ----------
 * Each call (re)defines one table: n entries, values, and one byte per entry
 * that is non-zero for parameters held constant (may be NULL = all variable).
 * After bslam_finalize, calling a setter again with the same n updates VALUES
 * only (is_const must be NULL or unchanged). */
BSLAM_API int bslam_set_poses_se3(bslam_solver* s, int n, const double* Rt /* n x 12 */, const uint8_t* is_const);
BSLAM_API int bslam_set_poses_se2(bslam_solver* s, int n, const double* Rt /* n x 6  */, const uint8_t* is_const);
BSLAM_API int bslam_set_points(bslam_solver* s, int n, const double* xyz /* n x 3 */, const uint8_t* is_const);
/* Generic vector parameters (anything that is `+=`-updated, problem.py:405-409):
 * n vectors, dims[i] entries each, values concatenated. */
BSLAM_API int bslam_set_vectors(bslam_solver* s, int n, const int32_t* dims, const double* values, const uint8_t* is_const);

/* SO(3) parameters (9 doubles, row-major; liegroups SO3: dof 3, perturb R <- exp(phi) R): the rotation half of the
 * (SO3, t) parameter form of the photometric residual (photometric_residual.py:83-84, pipelines/dense.py:185-190). */
BSLAM_API int bslam_set_rotations_so3(bslam_solver* s, int n, const double* R /* n x 9 */, const uint8_t* is_const);
BSLAM_API int bslam_get_rotations_so3(bslam_solver* s, double* R /* n x 9 */);

BSLAM_API int bslam_get_poses_se3(bslam_solver* s, double* Rt /* n x 12 */);
BSLAM_API int bslam_get_poses_se2(bslam_solver* s, double* Rt /* n x 6  */);
BSLAM_API int bslam_get_points(bslam_solver* s, double* xyz /* n x 3 */);
BSLAM_API int bslam_get_vectors(bslam_solver* s, double* values);

/* ---- residual blocks: `Problem.add_residual_block` (pyslam/problem.py:72-81) --
 * `stiffness` holds ONE matrix shared by the n blocks when per_block == 0, or
 * n matrices when per_block != 0. */

/* ReprojectionResidual + StereoCamera.project (pyslam/residuals/
 * reprojection_residual.py:13-37, pyslam/sensors/stereo_camera.py:100-134).
 * pose_idx -> SE3 table, pt_idx -> point table, obs = (u,v,d),
 * stiffness 3x3, intr = (cu,cv,fu,fv,b).  b > 0: StereoCamera (d = disparity); b <= 0 selects the RGB-D pinhole
 * model RGBDCamera (pyslam/sensors/rgbd_camera.py:7-168: third measurement = depth z) in every kernel that takes intr. */
BSLAM_API int bslam_add_reprojection_blocks(bslam_solver* s, int n,
                                  const int32_t* pose_idx, const int32_t* pt_idx,
                                  const double* obs /* n x 3 */,
                                  const double* stiffness, int per_block,
                                  const double intr[5], int loss_kind, double loss_k);

/* PoseResidual (pyslam/residuals/pose_residual.py:4-27): r = S log(T T_obs^-1),
 * J = S.  group = BSLAM_SE2 | BSLAM_SE3; T_obs n x (6|12); stiffness dof x dof. */
BSLAM_API int bslam_add_pose_blocks(bslam_solver* s, int group, int n, const int32_t* pose_idx,
                          const double* T_obs, const double* stiffness, int per_block,
                          int loss_kind, double loss_k);

/* PoseToPoseResidual (pyslam/residuals/pose_to_pose_residual.py:4-32):
 * r = S log(T2 T1^-1 T21_obs^-1), J1 = -S Ad(T2 T1^-1), J2 = S. */
BSLAM_API int bslam_add_pose_to_pose_blocks(bslam_solver* s, int group, int n,
                                  const int32_t* idx1, const int32_t* idx2,
                                  const double* T21_obs, const double* stiffness, int per_block,
                                  int loss_kind, double loss_k);

/* PhotometricResidualSE3 with a single SE3 parameter (pyslam/residuals/
 * photometric_residual.py:38-161): dense direct alignment of a reference image
 * (pixels already filtered by the constructor: valid disparity, gradient >= min_grad)
 * against `im_track` (height x width, row-major).  uvd_ref n x 3 (pixel u, v,
 * disparity), im_ref n, im_jac n x 2 (dI/du, dI/dv); stiffness = 1/sigma of the
 * intensity and of the disparity.  pose_idx -> SE3 table (T_track_ref). */
BSLAM_API int bslam_add_photometric_block(bslam_solver* s, int pose_idx, int n_px, const double* uvd_ref,
                                          const double* im_ref, const double* im_jac, const double* im_track,
                                          int width, int height, const double intr[5],
                                          double intensity_stiffness, double depth_stiffness,
                                          int loss_kind, double loss_k);
/* The same block on the two-parameter form ['R_1_0', 't_1_0_1'] (photometric_residual.py:83-84,147-157: T = SE3(R, t),
 * Jacobians J[:, 3:6] for the rotation and J[:, 0:3] for the translation; either may be constant, e.g. the
 * translation on the coarse pyramid levels, pipelines/dense.py:188-189).  rot_idx -> SO3 table, vec_idx -> a
 * 3-vector of the generic vector table. */
BSLAM_API int bslam_add_photometric_block_split(bslam_solver* s, int rot_idx, int vec_idx, int n_px, const double* uvd_ref,
                                      const double* im_ref, const double* im_jac, const double* im_track, int width,
                                      int height, const double intr[5], double intensity_stiffness, double depth_stiffness,
                                      int loss_kind, double loss_k);

/* Host-evaluated blocks (user-defined Python residuals, the plug-in surface of
 * pyslam/problem.py:338-360).  Declares the STRUCTURE once: n_blocks blocks,
 * block b has rows[b] residual rows and uses parameters
 * param_kind/param_index[param_ptr[b] .. param_ptr[b+1]) where kind is
 * 0 = SE3 table, 1 = SE2 table, 2 = point table, 3 = vector table.
 * Values are uploaded every iteration with bslam_upload_dense_values. */
BSLAM_API int bslam_set_dense_blocks(bslam_solver* s, int n_blocks, const int32_t* rows,
                           const int32_t* param_ptr, const int32_t* param_kind,
                           const int32_t* param_index);
/* e: concatenated sqrt(w)*r of all blocks; J: for every block, for every one of
 * its parameters in order, the row-major (rows x dof) matrix sqrt(w)*J (zeros
 * for constant parameters); cost = sum rho(r) over these blocks (added to
 * BSLAM_S_COST_LIN).  Must be called before each bslam_linearize/iterate when
 * dense blocks exist. */
BSLAM_API int bslam_upload_dense_values(bslam_solver* s, const double* e, size_t n_e,
                              const double* J, size_t n_J, double cost);

BSLAM_API int bslam_clear_blocks(bslam_solver* s);

/* How bslam_finalize lowers the reprojection blocks for bslam_iterate* (call before bslam_finalize):
 *   0  landmark-block kernels only: W = J_T^T w J_p is materialised (144 B per observation);
 *   1  (default) dense landmark panels -- runs of <= 64 landmarks seen by <= 8 poses whose
 *      (pose, landmark) grid is at least half full -- are linearised AND eliminated by one fused
 *      kernel that never writes W; everything else as in mode 0;
 *   2  every run of landmarks that fits a panel takes the fused kernel (tests).
 * The result of an iteration is the same up to rounding in all modes. */
BSLAM_API int bslam_set_fused(bslam_solver* s, int mode);
/* Number of panels / of landmarks inside panels chosen by the last bslam_finalize. */
BSLAM_API int bslam_get_fused(bslam_solver* s, int32_t* n_panels, int32_t* n_landmarks);

/* Lower the problem: orders observations by landmark, decides which points are
 * eliminated by the Schur complement, lays out the reduced system, allocates
 * device memory.  Replaces `_get_update_partition_dict` (problem.py:252-277)
 * for the internal ordering; see bslam_get_layout for the mapping back. */
BSLAM_API int bslam_finalize(bslam_solver* s);

/* Internal update-vector layout: for every table entry the offset of its
 * tangent slice in the internal dx (or -1 if constant).  The host maps this to
 * the reference ordering (param_dict insertion order, problem.py:252-277).
 * Any pointer may be NULL.  *dim = total length D of dx, *n_reduced = size of
 * the reduced (camera) system. */
BSLAM_API int bslam_get_layout(bslam_solver* s, int32_t* se3_off, int32_t* se2_off, int32_t* pt_off,
                     int32_t* vec_off, int32_t* dim, int32_t* n_reduced);
BSLAM_API int bslam_get_layout_so3(bslam_solver* s, int32_t* so3_off);   /* offsets of the SO3 table entries (-1: constant) */

/* ---- the hot path -------------------------------------------------------------- */

/* `Problem.eval_cost` (problem.py:110-128): sum of rho(r) over ALL built-in
 * blocks at the current parameters (dense blocks are the host's to add). */
BSLAM_API int bslam_eval_cost(bslam_solver* s, double* cost);

/* One Gauss-Newton / LM iteration, entirely on the device, one host sync:
 *   linearize  (problem.py:279-360)  -> block-sparse J^T W J, -J^T W r, cost
 *   solve      (problem.py:186)      -> Schur complement + dense Cholesky
 *   retract    (problem.py:155-156, 400-409) -> x <- x [+] dx
 *   cost       (problem.py:189-190, 362-398 net effect) if eval_new_cost
 * lambda = 0 is the reference's Gauss-Newton; lambda > 0 adds lambda*diag(H)
 * (extension, not in the reference).  Any out pointer may be NULL. */
BSLAM_API int bslam_iterate(bslam_solver* s, double lambda, int eval_new_cost,
                  double* cost_lin, double* cost_new, double* dx_norm);
/* The same iteration for a caller whose parameters live in HOST memory, as the reference's param_dict does
 * (pyslam/problem.py:143-156): upload the SE3 pose table (n x 12) and the point table (n x 3, user order),
 * iterate, download the updated tables, with ONE stream synchronisation for the whole step.  Pinned buffers
 * make the copies asynchronous; in == out is allowed; any pointer may be NULL (that transfer is skipped). */
BSLAM_API int bslam_iterate_host(bslam_solver* s, double lambda, int eval_new_cost,
                       const double* Rt_in, const double* xyz_in, double* Rt_out, double* xyz_out,
                       double* cost_lin, double* cost_new, double* dx_norm);

/* The same iteration in separately callable phases (parity tests, multi-GPU). */
BSLAM_API int bslam_linearize(bslam_solver* s, double* cost_lin);          /* no Schur yet          */
BSLAM_API int bslam_reduce(bslam_solver* s, double lambda);                /* damping + Schur        */
BSLAM_API int bslam_solve_reduced(bslam_solver* s);                        /* Cholesky + substitution + landmark back-substitution */
BSLAM_API int bslam_retract(bslam_solver* s, int eval_new_cost);           /* x <- x [+] dx (+ cost) */
BSLAM_API int bslam_get_scalars(bslam_solver* s, double* out /* BSLAM_N_SCALARS */);   /* syncs */
/* The scalars as the last bslam_iterate* / bslam_get_scalars call read them back (host mirror: no device work, no sync). */
BSLAM_API int bslam_last_scalars(bslam_solver* s, double* out /* BSLAM_N_SCALARS */);
/* Exactly what bslam_iterate runs before the reduced solve (fused panels included): linearise + damping +
 * landmark elimination.  bslam_get_reduced_system then returns the Schur complement the iteration
 * factorises; bslam_solve_reduced and bslam_retract_iterate complete the iteration. */
BSLAM_API int bslam_linearize_reduce(bslam_solver* s, double lambda);
BSLAM_API int bslam_retract_iterate(bslam_solver* s, int eval_new_cost);

/* Multi-GPU plumbing: device address/length (in doubles) of the contiguous
 * [S (n_pad x n_pad) | rhs (n_pad) | scalars (BSLAM_N_SCALARS)] buffer that the
 * host all-reduces (NCCL, sum) between bslam_reduce and bslam_solve_reduced,
 * and of the scalar tail alone (second, 16-double all-reduce after retract). */
BSLAM_API int bslam_reduced_buffer(bslam_solver* s, void** dev_ptr, size_t* n_doubles,
                         void** scalars_dev_ptr, int32_t* n_pad);
/* Compact all-reduce payload: only the structurally non-zero tiles of S (before fill-in),
 * then rhs and the scalars.  bslam_pack_reduced(s, 0) gathers it from the dense buffer after
 * bslam_reduce, the host all-reduces bslam_packed_buffer, bslam_pack_reduced(s, 1) scatters it back
 * before bslam_solve_reduced.  (The tile structures of all ranks must have been merged first.) */
BSLAM_API int bslam_packed_buffer(bslam_solver* s, void** dev_ptr, size_t* n_doubles);
BSLAM_API int bslam_pack_reduced(bslam_solver* s, int unpack);
/* The two halves of a sharded iteration as single calls (each recorded once as a CUDA graph and replayed):
 *   bslam_iterate_pre  = bslam_linearize + bslam_reduce(lambda) + bslam_pack_reduced(0)
 *   [the host all-reduces bslam_packed_buffer over the ranks]
 *   bslam_iterate_post = bslam_pack_reduced(1) + bslam_solve_reduced + bslam_retract(eval_new_cost)
 * Neither synchronises; read the scalars with bslam_get_scalars after the ranks' scalar all-reduce. */
BSLAM_API int bslam_iterate_pre(bslam_solver* s, double lambda);
BSLAM_API int bslam_iterate_post(bslam_solver* s, int eval_new_cost);
/* Tile structure (tiles of bslam_tile_edge(), (nt + 1) x nt bytes with nt = n_pad / edge, row-major) of
 * the reduced system as seen by THIS handle's residual blocks.  set == 0 copies it
 * out; set != 0 ORs `mask` into it -- with sharded landmarks the host ORs the
 * masks of all ranks (all-reduce MAX) once after bslam_finalize, because the
 * summed matrix has the union structure. */
BSLAM_API int bslam_tile_structure(bslam_solver* s, uint8_t* mask, size_t n, int set);
/* Shard rank of this handle when landmarks are partitioned over several GPUs
 * (rank 0 alone counts the replicated reduced part in ||dx||^2). */
BSLAM_API int bslam_set_shard(bslam_solver* s, int rank);
/* ---- sharded iteration over NVLink peer memory (pyslam_b200/csrc/peer.cuh; the reference has no distributed
 * path, SURVEY 2.2 -- this is the B200-native replacement for the single-process spsolve on the full system) ----
 *
 * bslam_add_coupling: declare (before bslam_finalize) that poses idx1[k] and idx2[k] of `group` are coupled in the
 * reduced system WITHOUT adding a residual.  A rank that holds only a shard of the landmarks declares the
 * co-visibility pairs of ALL shards, so that every rank derives the same ordering, offsets and tile structure
 * (bslam_layout_hash must then agree across ranks).
 * bslam_peer_region: device address, size and CUDA-IPC handle (64 bytes, may be NULL) of this handle's exchange
 * region [packed non-zero tiles | rhs | scalars | mailbox | flags].
 * bslam_peer_connect: map the regions of all `world` ranks -- from their IPC handles (world x 64 bytes, other
 * processes) or from plain device pointers (dev_ptrs[r] != NULL: handles of the same process) -- and switch
 * bslam_iterate / bslam_iterate_async / bslam_iterate_host to the sharded schedule: linearise + eliminate the
 * local landmarks, publish the partial reduced system, rendezvous, factorise sum_r S_r read tile by tile from
 * the peers' regions (the all-reduce is fused into the Cholesky kernel's operand loads), back-substitute the
 * local landmarks, exchange the partial scalars -- ONE CUDA graph, no host code or library collective inside.
 * The returned scalars are the sums over all ranks; every rank must call iterate the same number of times. */
BSLAM_API int bslam_add_coupling(bslam_solver* s, int group, int n, const int32_t* idx1, const int32_t* idx2);
BSLAM_API int bslam_layout_hash(bslam_solver* s, uint64_t* hash);
BSLAM_API int bslam_peer_region(bslam_solver* s, void** dev_ptr, size_t* n_bytes, uint8_t* ipc_handle /* 64 bytes */);
BSLAM_API int bslam_peer_connect(bslam_solver* s, int world, int rank, const uint8_t* ipc_handles, void* const* dev_ptrs);
/* The same with caller-provided SYMMETRIC memory (one allocation of >= bslam_peer_region's size per rank, zero-filled,
 * e.g. torch.distributed._symmetric_memory): region_ptrs[r] = rank r's allocation as mapped in this process (the own
 * one included -- the handle publishes its partial system there), multicast_ptr = the NVLS multicast mapping of all of
 * them or NULL.  With a multicast mapping the Cholesky kernel reads every element of sum_r S_r with ONE
 * multimem.ld_reduce (reduced inside the NVSwitch) instead of one load per rank. */
BSLAM_API int bslam_peer_connect_symmetric(bslam_solver* s, int world, int rank, void* const* region_ptrs, void* multicast_ptr,
                                 size_t n_bytes);
/* Which ranks contribute to which tile.  bslam_peer_local_slots: flags[k] != 0 when THIS handle's residual blocks can
 * write slot k of the packed payload (n = number of packed tiles, bslam_packed_buffer's tile count); the ranks exchange
 * the flags once and call bslam_peer_set_contributors with masks[k] = OR over ranks r of (flags_r[k] != 0) << r.  The
 * Cholesky kernel then reads a tile only from the ranks that can have written it: with time-contiguous landmark shards
 * the NVLink traffic of the fused all-reduce drops from (world - 1) x payload to ~1 x payload per rank.  Optional. */
BSLAM_API int bslam_peer_local_slots(bslam_solver* s, uint8_t* flags, size_t n);
BSLAM_API int bslam_peer_set_contributors(bslam_solver* s, const uint8_t* masks, size_t n);
/* Device-side rendezvous of all connected ranks, enqueued on the handle's stream (no host synchronisation): the
 * kernels enqueued after it start on every rank within a few microseconds of each other (bench.py aligns the ranks
 * with it before each timed step).  No-op for a single rank. */
BSLAM_API int bslam_peer_barrier(bslam_solver* s);
/* bslam_iterate split in two: enqueue the iteration (no synchronisation) / wait for it and read the scalars.
 * Several handles of one process (e.g. the shards of a sharded iteration on different streams) are enqueued
 * first and waited for afterwards. */
BSLAM_API int bslam_iterate_async(bslam_solver* s, double lambda, int eval_new_cost);
BSLAM_API int bslam_iterate_wait(bslam_solver* s, double* cost_lin, double* cost_new, double* dx_norm);
/* The handle's cudaStream_t, so the host can order collectives and record
 * events on it. */
BSLAM_API void* bslam_stream(bslam_solver* s);

/* ---- frame-to-frame motion -------------------------------------------------------------------------
 * ReprojectionMotionOnlyBatchResidual / ReprojectionMotionOnlyResidual (pyslam/residuals/
 * reprojection_motion_only_residual.py:36-113): n points triangulated in frame 1 (pts_1, n x 3) observed in frame 2
 * (obs_2, n x 3), one SE3 parameter T_2_1 = pose_idx; stiffness 3x3 shared; intr as above. */
BSLAM_API int bslam_add_motion_only_blocks(bslam_solver* s, int pose_idx, int n, const double* pts_1, const double* obs_2,
                                 const double* stiffness, const double intr[5], int loss_kind, double loss_k);
/* PoseToPoseOrientationResidual (pyslam/residuals/pose_to_pose_orientation_residual.py:4-38): binary factor on two SE3
 * poses from a relative rotation measurement C_2_1_obs (n x 9, row-major), stiffness 3x3. */
BSLAM_API int bslam_add_orientation_blocks(bslam_solver* s, int n, const int32_t* idx1, const int32_t* idx2, const double* C21_obs,
                                 const double* stiffness, int per_block, int loss_kind, double loss_k);
/* FrameToFrameRANSAC.perform_ransac on the device (pyslam/pipelines/ransac.py:12-56,107-165), no solver handle:
 * n_hyp hypotheses.  idx != NULL: minimal sets idx[n_hyp][n_min] (rows of pts_1 / pts_2, n_pts x 3 each) -> rigid
 * transforms by the SVD method (compute_transform_fast), returned in T_21_out (n_hyp x 16, 4x4 row-major);
 * idx == NULL: the hypotheses are given in T_21_in.  Then compute_ransac_cost: inlier counts
 * |project(T_21 pts_1) - obs_2|^2 < thresh per hypothesis (counts, n_hyp), the first arg-max and its count
 * (best[0], best[1]) and the winner's inlier mask (best_mask, n_pts bytes).  Outputs may be NULL. */
BSLAM_API int bslam_ransac(int device, int n_hyp, int n_min, const int32_t* idx, const double* T_21_in, int n_pts,
                 const double* pts_1, const double* pts_2, const double* obs_2, const double intr[5], double thresh,
                 double* T_21_out, int32_t* counts, int32_t* best, uint8_t* best_mask);
/* ---- pyramids of the dense pipeline (pyslam/pipelines/keyframes.py:30-46,59-72,92-114), no handle ----
 * bslam_image_pyramid: 8-bit image -> `levels` levels (cv2.pyrDown chain on the 8-bit data), each as float / 255
 * (im_out) with 0.5 * Sobel gradients (gx_out, gy_out; may both be NULL); levels concatenated, level l has
 * ceil(w / 2^l) x ceil(h / 2^l) pixels.  bslam_subsample_pyramid: disparity / depth maps: level l = map[::2^l, ::2^l]
 * * scale_per_level^l (0.5 for disparities in pixels, 1 for depths). */
BSLAM_API int bslam_image_pyramid(int device, const uint8_t* image, int width, int height, int levels, double* im_out,
                        double* gx_out, double* gy_out);
BSLAM_API int bslam_subsample_pyramid(int device, const double* map, int width, int height, int levels,
                            double scale_per_level, double* out);

/* Best-parameter snapshot used by `allow_nondecreasing_steps` (problem.py:163-175). */
BSLAM_API int bslam_snapshot(bslam_solver* s);
BSLAM_API int bslam_restore(bslam_solver* s);

/* ---- inspection (parity tests, covariance) ---------------------------------------- */

/* Update vector of the last solve, internal ordering, length D. */
BSLAM_API int bslam_get_update(bslam_solver* s, double* dx);
/* Dense D x D normal matrix H = J^T W J and right-hand side b = -J^T W r of the
 * last bslam_linearize, internal ordering (small problems only: D <= 20000). */
BSLAM_API int bslam_get_normal_equations(bslam_solver* s, double* H, double* b);
/* Reduced system after bslam_reduce: S (n_reduced x n_reduced, symmetric, full) and rhs. */
BSLAM_API int bslam_get_reduced_system(bslam_solver* s, double* S, double* rhs);
/* Dense covariance (H^-1, D x D, internal ordering) at the current parameters:
 * `Problem.compute_covariance` (problem.py:196-203). */
BSLAM_API int bslam_covariance(bslam_solver* s, double* cov);

/* Debug aid: run one reduced solve with per-task tracing of the Cholesky kernel.
 * out[6*t..] = (tile row | -1, tile column, start ns, dependencies-ready ns, end ns, SM id). */
BSLAM_API int bslam_debug_chol_trace(bslam_solver* s, int64_t* out, int max_tasks, int* n_tasks);

/* Phase timings of the last bslam_iterate with timing enabled (ms). */
BSLAM_API int bslam_enable_timing(bslam_solver* s, int on);
BSLAM_API int bslam_get_timings(bslam_solver* s, double* ms /* BSLAM_N_TIMINGS */);
/* Number of kernels launched by this handle since creation (bench.py's gpu_launches). */
BSLAM_API int64_t bslam_launch_count(const bslam_solver* s);

#ifdef __cplusplus
}
#endif
#endif /* BSLAM_H_ */

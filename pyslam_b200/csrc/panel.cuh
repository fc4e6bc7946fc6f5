// panel.cuh -- fused linearisation + landmark elimination for DENSE landmark panels:
// the product path of bundle adjustment.  W = J_T^T w J_p is never written to HBM.
//
// Replaces, for the reprojection blocks of a panel, in ONE kernel per iteration
//   ReprojectionResidual.evaluate / StereoCamera.project   pyslam/residuals/reprojection_residual.py:13-37,
//                                                           pyslam/sensors/stereo_camera.py:100-134
//   sqrt(loss.weight) scaling, cost                         pyslam/problem.py:349-360
//   HT.HT^T, -HT.e                                          pyslam/problem.py:329-333
//   the landmark part of spsolve(H, g)                      pyslam/problem.py:186   (Schur complement)
// and, in panel_finish_kernel, the landmark part of the update + the cost at x [+] dx
//   (pyslam/problem.py:155-156, 189-190, 400-409).
//
// A PANEL is a run of <= 64 consecutive landmarks (internal order: sorted by the first pose that
// sees them) together with the <= 8 distinct poses ("rows") that observe them; its observations are
// stored as a padded grid  cell = (row, landmark)  [n_rows][64]  with a 64-bit presence mask per row
// (bslam_finalize builds panels only where the grid is well filled; everything else takes the
// landmark-block kernels of reproj.cuh / schur.cuh).  A CTA of 8 warps walks panels:
//
//   P1  warp = row, lane = two landmarks: structured linearisation (reproj_blocks), cost;
//       camera values U_c, b_c (27) summed over the row by a register butterfly (no shared memory)
//       -> one fp64 atomic per value; landmark values V_p, b_p (9) -> shared partials (one slot per row);
//       W (6x3) -> shared memory, already in the operand layout of P4.
//   P2  thread = landmark: V_p = sum of partials, (V_p + lambda diag)^-1 -> HBM (back-substitution),
//       Cholesky V = L L^T, c = L^-1 b_p.
//   P3  Z = W L^-T in place (so that W V^-1 W'^T = Z Z'^T and W V^-1 b_p = Z c).
//   P4  S -= Zbig Zbig^T, rhs -= Zbig c: a symmetric rank-192 update of the panel's
//       (6 n_var + 1)-row operand [Z_0; ...; Z_{n_var-1}; c] on 8x8x4 fp64 tensor-core tiles
//       (mma.sync.m8n8k4.f64 = DMMA; tcgen05 has no fp64 kind), K = 4 packs landmark/coordinate
//       pairs without padding, lower-triangular tiles only, one fp64 atomic per element.
//
// Algorithmic HBM bytes per iteration: 32 B per observation (u, v, d + padding of the grid) read,
// 24 B per landmark read, 120 B per landmark written (V_p | b_p | V^-1), 432 B per pose of atomics
// -> SURVEY 8(d)'s fused figure 32 N + 432 K + 120 L.  The kernel is bound by the fp64 pipe
// (DFMA + DMMA share it on sm_100a), not by HBM: see DESIGN.md section 3.
#pragma once
#include "cholesky.cuh"
#include "common.cuh"
#include "reproj.cuh"
#include "schur.cuh"

namespace bs {

constexpr int kPanelLm = 64;        // landmarks per panel (two per lane)
constexpr int kPanelRows = 8;       // poses per panel = warps per CTA
constexpr int kPanelMaxVar = 7;     // variable poses per panel (unless a single landmark needs 8): two CTAs of 110 KB per SM
constexpr int kPanelThreads = 32 * kPanelRows;
constexpr int kZGroup = 76;         // doubles per group of 4 landmarks of one row: 6 x 12 + 4 (bank shift)
constexpr int kZRow = (kPanelLm / 4) * kZGroup;
constexpr int kCGroup = 12;

struct PanelRow {
  int pose;                 // index into the SE3 table
  int off;                  // reduced offset of the pose, -1: constant
  unsigned mask_lo, mask_hi;  // landmark j of the panel is observed by this pose
};
struct Panel {
  int lm_begin, n_lms;      // landmarks [lm_begin, lm_begin + n_lms)
  int n_rows;               // rows of the panel: the variable poses first
  int n_var;                // number of variable rows
  int pad0, pad1, pad2, pad3;
};
// Everything a CTA needs to know about a panel, 160 contiguous bytes at a FIXED stride: fetched with one bulk copy
// (cp.async.bulk + mbarrier), and nothing else that is prefetched depends on its contents:
//   observations   pobs[panel][u|v|d][8 rows][64]   fixed stride (zeros where a pose does not see a landmark)
//   points         pts + 3 * lm_begin               (lm_begin / n_lms requested one panel earlier)
//   poses          poses + 12 * rows[warp].pose     (pose index requested one panel earlier)
struct PanelDesc {
  Panel hdr;
  PanelRow rows[kPanelRows];
};
static_assert(sizeof(PanelDesc) == 160, "PanelDesc is one 160-byte bulk copy");
constexpr int kPanelObs = 3 * kPanelRows * kPanelLm;      // doubles of observations per panel

struct PanelArgs {
  int n_panels;
  const PanelDesc* __restrict__ descs;      // [n_panels]
  const double* __restrict__ pobs;          // [n_panels][3][8][64] observations of the cells (0 where absent)
  const unsigned short* __restrict__ pgrp;  // [n_panels][8][64] group per cell (kLoss < 0 only)
  const ReprojGroup* __restrict__ groups;
  ReprojGroup g0;
  const double* __restrict__ poses;         // [K][12] at the linearisation point
  const double* __restrict__ poses_new;     // finish: retracted poses
  const double* __restrict__ pts_in;        // [P][3]
  double* __restrict__ pts;                 // finish: updated in place (same buffer as pts_in)
  double lambda;
  double* __restrict__ Vg;                  // [n_lm][9] V_p | b_p
  double* __restrict__ Vinv;                // [n_lm][6]
  double* __restrict__ S;
  int ldS;
  double* __restrict__ rhs;
  double* __restrict__ scalars;
  const double* __restrict__ dx_red;        // finish: reduced update (dx_c at the poses' offsets)
  double* __restrict__ dx_lm;               // finish: landmark part of the update vector
  int eval_cost;
  int max_var;                              // shared-memory carve-up (rows of Z)
};

BS_HD size_t panel_smem_bytes(int max_var) {
  return sizeof(double) * ((size_t)max_var * kZRow + (kPanelLm / 4) * kCGroup + kPanelRows * 9 * kPanelLm + 6 * kPanelLm + 3 * kPanelLm + 2 +
                           12 * kPanelRows) +
         sizeof(int) * 64 + sizeof(PanelDesc);
}

// Sum over the 32 lanes of 32 values per lane; on return v[0] of lane L holds the total of value L.
// 31 exchanges instead of 32 x 5 for a plain xor reduction of every value.
BS_D void warp_transpose_sum32(double (&v)[32], int lane) {
#pragma unroll
  for (int h = 16; h >= 1; h >>= 1) {
    const bool up = (lane & h) != 0;
#pragma unroll
    for (int i = 0; i < h; ++i) {
      const double send = up ? v[i] : v[i + h];
      const double keep = up ? v[i + h] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, h);
    }
  }
}

// index t of the lower-triangular tile list (row-major) -> (I, J)
BS_D void tile_ij(int t, int& I, int& J) {
  int i = 0;
  while ((i + 1) * (i + 2) / 2 <= t) ++i;
  I = i;
  J = t - i * (i + 1) / 2;
}

constexpr int kMaxTiles = 28;       // lower-triangular 8x8 tiles of a (6 * 8 + 1)-row operand
constexpr int kTilesPerWarp = 4;    // ceil(28 / 8)

// ---- TMA bulk copies (cp.async.bulk, UBLKCP in SASS) completing on an mbarrier (SYNCS): the panel descriptor and the
// panel's points are staged global -> shared by the copy engine, one elected thread issues them, nobody blocks on them
// until the next panel starts.
BS_D unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
BS_D void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
BS_D void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
BS_D void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
BS_D void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// The point run of a panel starts at byte 24 * lm_begin: 16-byte aligned for even lm_begin, else the copy starts one
// double earlier (and the staged points are read at an offset of one double).
BS_HD unsigned panel_pts_bytes(int lm_begin, int n_lms) { return (unsigned)((24 * n_lms + 8 * (lm_begin & 1) + 15) & ~15); }

template <int kLoss>
__global__ void __launch_bounds__(kPanelThreads, 2) fused_panel_kernel(const PanelArgs a) {
  extern __shared__ __align__(16) double psm[];
  __shared__ double sred[kPanelThreads / 32];
  __shared__ unsigned char sTileI[kMaxTiles], sTileJ[kMaxTiles];
  __shared__ int sZoff[64];                                   // operand row -> offset of its first element in psm
  __shared__ __align__(8) unsigned long long s_mbar;          // completion of the staged descriptor + points
  double* sZ = psm;                                           // [max_var][16 groups][76]
  double* sC = sZ + (size_t)a.max_var * kZRow;                // [16 groups][12]
  double* sVp = sC + (kPanelLm / 4) * kCGroup;                // [8 rows][9][64]
  double* sL = sVp + kPanelRows * 9 * kPanelLm;               // [6][64]: 1/l00, l10, 1/l11, l20, l21, 1/l22
  double* sPtsBase = sL + 6 * kPanelLm;                       // [64][3] + 2 (bulk copy, 16-byte aligned)
  double* sPose = sPtsBase + 3 * kPanelLm + 2;                // [8 rows][12]  (bulk copies, one per warp)
  int* sIdx = reinterpret_cast<int*>(sPose + 12 * kPanelRows);   // operand row -> index in S / rhs
  PanelDesc* sDesc = reinterpret_cast<PanelDesc*>(sIdx + 64); // header + rows of the current panel (bulk copy)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int c_off = a.max_var * kZRow;                        // offset of sC
  double cost = 0.0;
  // lane -> position of its camera value inside the pose's diagonal block of S (lower triangle, row-major)
  int tri_off;
  {
    int r = 0;
    while ((r + 1) * (r + 2) / 2 <= lane) ++r;
    tri_off = lane < 21 ? r * a.ldS + (lane - r * (r + 1) / 2) : 0;
  }
  if (tid < kMaxTiles) {
    int I, J;
    tile_ij(tid, I, J);
    sTileI[tid] = (unsigned char)I; sTileJ[tid] = (unsigned char)J;
  }
  const double damp = 1.0 + a.lambda;

  int pn = blockIdx.x;
  if (pn >= a.n_panels) return;
  if (tid == 0) mbar_init(&s_mbar, 1 + kPanelRows);   // thread 0 (descriptor + points) and lane 0 of every warp (its pose)
  __syncthreads();

  // What is prefetched for a panel, and from where:
  //   descriptor + points  -> shared memory, bulk copies issued by thread 0 (lm_begin / n_lms come from `hdr`, which
  //                           was requested with a plain load a whole panel earlier)
  //   observations         -> registers, fixed-stride addresses (no dependency on anything loaded)
  //   pose of the row      -> shared memory, one 96-byte bulk copy per warp, index requested a whole panel earlier
  struct Obs { double u[2], v[2], d[2]; };
  auto stage = [&](int panel, int lm_begin, int n_lms, int pose) {
    if (lane == 0) {
      mbar_expect_tx(&s_mbar, 96u);
      bulk_g2s(sPose + 12 * warp, a.poses + 12 * (size_t)pose, 96u, &s_mbar);
    }
    if (tid == 0) {
      Panel hdr; hdr.lm_begin = lm_begin; hdr.n_lms = n_lms;
      const unsigned nb = panel_pts_bytes(hdr.lm_begin, hdr.n_lms);
      mbar_expect_tx(&s_mbar, (unsigned)sizeof(PanelDesc) + nb);
      bulk_g2s(sDesc, a.descs + panel, (unsigned)sizeof(PanelDesc), &s_mbar);
      bulk_g2s(sPtsBase, a.pts_in + 3 * (size_t)hdr.lm_begin - (hdr.lm_begin & 1), nb, &s_mbar);
    }
  };
  auto load_obs = [&](int panel, Obs& ob) {
    const double* base = a.pobs + (size_t)panel * kPanelObs + warp * kPanelLm + lane;
#pragma unroll
    for (int sub = 0; sub < 2; ++sub) {
      ob.u[sub] = ld_stream(base + 32 * sub);
      ob.v[sub] = ld_stream(base + kPanelRows * kPanelLm + 32 * sub);
      ob.d[sub] = ld_stream(base + 2 * kPanelRows * kPanelLm + 32 * sub);
    }
  };
  Obs ob;
  {   // first panel of the CTA: nothing was requested ahead
    const Panel hdr = a.descs[pn].hdr;
    stage(pn, hdr.lm_begin, hdr.n_lms, a.descs[pn].rows[warp].pose);
    load_obs(pn, ob);
  }
  unsigned parity = 0;

  for (;;) {
    // scalars of the NEXT panel: requested now, first used after P3 (never waited for)
    const int pn_next = pn + gridDim.x;
    const bool has_next = pn_next < a.n_panels;
    int n_lm_begin = 0, n_n_lms = 0, npose = 0;
    if (has_next) {
      n_lm_begin = a.descs[pn_next].hdr.lm_begin;
      n_n_lms = a.descs[pn_next].hdr.n_lms;
      npose = a.descs[pn_next].rows[warp].pose;
    }
    __syncthreads();                 // the tile phase of the previous panel is over
    mbar_wait(&s_mbar, parity);      // descriptor + points of this panel have landed
    parity ^= 1u;
    const Panel pan = sDesc->hdr;
    const double* sPts = sPtsBase + (pan.lm_begin & 1);

    // ---------------------------------------------------------------- P1: one row per warp
    if (warp < pan.n_rows) {
      const PanelRow row = sDesc->rows[warp];
      const bool isvar = warp < pan.n_var;
      double P[12];
#pragma unroll
      for (int k = 0; k < 12; ++k) P[k] = sPose[12 * warp + k];
      double U[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) U[k] = 0.0;
#pragma unroll
      for (int sub = 0; sub < 2; ++sub) {
        const int j = lane + 32 * sub;
        const bool present = (((sub ? row.mask_hi : row.mask_lo) >> lane) & 1u) != 0;
        double* zc = sZ + warp * kZRow + (j >> 2) * kZGroup + 3 * (j & 3);
        double* vq = sVp + warp * 9 * kPanelLm + j;            // landmark partials of this row: [9][64]
        if (present) {
          const size_t cell = ((size_t)pn * kPanelRows + warp) * kPanelLm + j;
          const ReprojGroup& grp = kLoss >= 0 ? a.g0 : a.groups[a.pgrp[cell]];
          double X[3];
#pragma unroll
          for (int k = 0; k < 3; ++k) X[k] = sPts[3 * j + k];
          ReprojBlocks o;
          reproj_blocks<kLoss>(grp, P, X, ob.u[sub], ob.v[sub], ob.d[sub], o);
          cost += o.cost;
          // V_p = R^T (M R) (xx xy xz yy yz zz), b_p = -R^T t
          vq[0] = P[0] * o.MR[0] + P[3] * o.MR[3] + P[6] * o.MR[6];
          vq[kPanelLm] = P[0] * o.MR[1] + P[3] * o.MR[4] + P[6] * o.MR[7];
          vq[2 * kPanelLm] = P[0] * o.MR[2] + P[3] * o.MR[5] + P[6] * o.MR[8];
          vq[3 * kPanelLm] = P[1] * o.MR[1] + P[4] * o.MR[4] + P[7] * o.MR[7];
          vq[4 * kPanelLm] = P[1] * o.MR[2] + P[4] * o.MR[5] + P[7] * o.MR[8];
          vq[5 * kPanelLm] = P[2] * o.MR[2] + P[5] * o.MR[5] + P[8] * o.MR[8];
          vq[6 * kPanelLm] = -(P[0] * o.t[0] + P[3] * o.t[1] + P[6] * o.t[2]);
          vq[7 * kPanelLm] = -(P[1] * o.t[0] + P[4] * o.t[1] + P[7] * o.t[2]);
          vq[8 * kPanelLm] = -(P[2] * o.t[0] + P[5] * o.t[1] + P[8] * o.t[2]);
          if (isvar) {
            // U_c lower triangle (row-major): rows 0-2 M; rows 3-5 [(M B)^T | B^T M B]; then b_c = -[t; B^T t]
            U[0] += o.M[0]; U[1] += o.M[1]; U[2] += o.M[3]; U[3] += o.M[2]; U[4] += o.M[4]; U[5] += o.M[5];
            U[6] += o.MB[0]; U[7] += o.MB[3]; U[8] += o.MB[6]; U[9] += o.BMB[0];
            U[10] += o.MB[1]; U[11] += o.MB[4]; U[12] += o.MB[7]; U[13] += o.BMB[1]; U[14] += o.BMB[3];
            U[15] += o.MB[2]; U[16] += o.MB[5]; U[17] += o.MB[8]; U[18] += o.BMB[2]; U[19] += o.BMB[4]; U[20] += o.BMB[5];
            U[21] -= o.t[0]; U[22] -= o.t[1]; U[23] -= o.t[2];
            U[24] -= o.y * o.t[2] - o.z * o.t[1];
            U[25] -= o.z * o.t[0] - o.x * o.t[2];
            U[26] -= o.x * o.t[1] - o.y * o.t[0];
            // W = [M R; B^T M R], element (r, c) at zc[12 r + c]
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              zc[c] = o.MR[c];
              zc[12 + c] = o.MR[3 + c];
              zc[24 + c] = o.MR[6 + c];
              zc[36 + c] = o.y * o.MR[6 + c] - o.z * o.MR[3 + c];
              zc[48 + c] = o.z * o.MR[c] - o.x * o.MR[6 + c];
              zc[60 + c] = o.x * o.MR[3 + c] - o.y * o.MR[c];
            }
          }
        } else {
#pragma unroll
          for (int k = 0; k < 9; ++k) vq[k * kPanelLm] = 0.0;
          if (isvar) {
#pragma unroll
            for (int r = 0; r < 6; ++r)
#pragma unroll
              for (int c = 0; c < 3; ++c) zc[12 * r + c] = 0.0;
          }
        }
      }
      if (isvar) {
        if (a.lambda > 0.0) { U[0] *= damp; U[2] *= damp; U[5] *= damp; U[9] *= damp; U[14] *= damp; U[20] *= damp; }
        warp_transpose_sum32(U, lane);
        if (lane < 21) red_add(a.S + (size_t)row.off * (a.ldS + 1) + tri_off, U[0]);
        else if (lane < 27) red_add(a.rhs + row.off + (lane - 21), U[0]);
      }
    }
    __syncthreads();

    // ---------------------------------------------------------------- P2: one landmark per thread
    if (tid < kPanelLm) {
      const int j = tid;
      double V[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) V[k] = 0.0;
      for (int rr = 0; rr < pan.n_rows; ++rr) {
#pragma unroll
        for (int k = 0; k < 9; ++k) V[k] += sVp[(rr * 9 + k) * kPanelLm + j];
      }
      double i00 = 1.0, l10 = 0.0, i11 = 1.0, l20 = 0.0, l21 = 0.0, i22 = 1.0, c0 = 0.0, c1 = 0.0, c2 = 0.0;
      if (j < pan.n_lms) {
        const size_t q = (size_t)pan.lm_begin + j;
#pragma unroll
        for (int k = 0; k < 9; ++k) a.Vg[9 * q + k] = V[k];
        double vi[6];
        sym3_inverse(V, a.lambda, vi);
#pragma unroll
        for (int k = 0; k < 6; ++k) a.Vinv[6 * q + k] = vi[k];
        // (V + lambda diag V) = L L^T
        const double v00 = V[0] * damp, v11 = V[3] * damp, v22 = V[5] * damp;
        i00 = fast_rsqrt(v00);
        l10 = V[1] * i00; l20 = V[2] * i00;
        i11 = fast_rsqrt(v11 - l10 * l10);
        l21 = (V[4] - l20 * l10) * i11;
        i22 = fast_rsqrt(v22 - l20 * l20 - l21 * l21);
        c0 = V[6] * i00;
        c1 = (V[7] - l10 * c0) * i11;
        c2 = (V[8] - l20 * c0 - l21 * c1) * i22;
      }
      sL[j] = i00; sL[kPanelLm + j] = l10; sL[2 * kPanelLm + j] = i11;
      sL[3 * kPanelLm + j] = l20; sL[4 * kPanelLm + j] = l21; sL[5 * kPanelLm + j] = i22;
      double* cc = sC + (j >> 2) * kCGroup + 3 * (j & 3);
      cc[0] = c0; cc[1] = c1; cc[2] = c2;
    } else if (tid < kPanelLm + 64) {
      const int zr = tid - kPanelLm;
      int idx = -1, zo = c_off;
      if (zr < 6 * pan.n_var) {
        const int rr = zr / 6, ri = zr - 6 * rr;
        idx = sDesc->rows[rr].off + ri;
        zo = rr * kZRow + ri * 12;
      } else if (zr == 6 * pan.n_var) {
        idx = -2;                                   // the row of c: its products go to the right-hand side
      }
      sIdx[zr] = idx;
      sZoff[zr] = zo;
    }
    __syncthreads();

    // ---------------------------------------------------------------- P3: Z = W L^-T in place
    if (warp < pan.n_var) {
      const PanelRow row = sDesc->rows[warp];
#pragma unroll
      for (int sub = 0; sub < 2; ++sub) {
        const int j = lane + 32 * sub;
        if (!((((sub ? row.mask_hi : row.mask_lo) >> lane) & 1u))) continue;
        double* zc = sZ + warp * kZRow + (j >> 2) * kZGroup + 3 * (j & 3);
        const double i00 = sL[j], l10 = sL[kPanelLm + j], i11 = sL[2 * kPanelLm + j];
        const double l20 = sL[3 * kPanelLm + j], l21 = sL[4 * kPanelLm + j], i22 = sL[5 * kPanelLm + j];
#pragma unroll
        for (int r = 0; r < 6; ++r) {
          const double w0 = zc[12 * r], w1 = zc[12 * r + 1], w2 = zc[12 * r + 2];
          const double z0 = w0 * i00;
          const double z1 = (w1 - z0 * l10) * i11;
          const double z2 = (w2 - z0 * l20 - z1 * l21) * i22;
          zc[12 * r] = z0; zc[12 * r + 1] = z1; zc[12 * r + 2] = z2;
        }
      }
    }
    __syncthreads();

    // ---- everything the NEXT panel needs goes in flight now: sDesc / sPts are not read by the tile phase, the
    //      addresses come from values requested at the top of this panel
    if (has_next) {
      stage(pn_next, n_lm_begin, n_n_lms, npose);
      load_obs(pn_next, ob);
    }

    // ---------------------------------------------------------------- P4: S -= Zbig Zbig^T on DMMA tiles
    {
      const int n_zr = 6 * pan.n_var + 1;
      const int T = (n_zr + 7) >> 3;
      const int ntiles = T * (T + 1) / 2;
      const int ng = (pan.n_lms + 3) >> 2;
      // consecutive tiles per warp (they mostly share their row operand)
      const int base = ntiles / kPanelRows, rem = ntiles % kPanelRows;
      const int t_first = warp * base + min(warp, rem);
      const int nt_w = base + (warp < rem ? 1 : 0);
      int tI[kTilesPerWarp], tJ[kTilesPerWarp];
      const double* pa[kTilesPerWarp];
      const double* pb[kTilesPerWarp];
      int sa[kTilesPerWarp], sb[kTilesPerWarp];
      double acc[kTilesPerWarp][2];
#pragma unroll
      for (int k = 0; k < kTilesPerWarp; ++k) {
        const int ti = min(t_first + k, ntiles - 1);
        tI[k] = sTileI[ti]; tJ[k] = sTileJ[ti];
        const int xa = 8 * tI[k] + g, xb = 8 * tJ[k] + g;
        const int oa = sZoff[xa], ob_ = sZoff[xb];
        pa[k] = psm + oa + t; sa[k] = oa >= c_off ? kCGroup : kZGroup;
        pb[k] = psm + ob_ + t; sb[k] = ob_ >= c_off ? kCGroup : kZGroup;
        acc[k][0] = acc[k][1] = 0.0;
      }
      for (int G = 0; G < ng; ++G) {
#pragma unroll
        for (int s_ = 0; s_ < 3; ++s_) {
          double av[kTilesPerWarp];
#pragma unroll
          for (int k = 0; k < kTilesPerWarp; ++k) {
            if (k < nt_w) {
              av[k] = (k > 0 && tI[k] == tI[k - 1]) ? av[k - 1] : pa[k][G * sa[k] + 4 * s_];
              const double bv = pb[k][G * sb[k] + 4 * s_];
              dmma_8x8x4(acc[k][0], acc[k][1], av[k], bv);
            }
          }
        }
      }
#pragma unroll
      for (int k = 0; k < kTilesPerWarp; ++k) {
        if (k >= nt_w) continue;
        const int x = 8 * tI[k] + g;
        const int ix = sIdx[x];
        if (ix == -1) continue;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int y = 8 * tJ[k] + 2 * t + h;
          if (y > x) continue;
          const int iy = sIdx[y];
          if (iy < 0) continue;
          const double c = acc[k][h];
          if (ix == -2) red_add(a.rhs + iy, -c);
          else {
            const int hi = max(ix, iy), lo = min(ix, iy);
            red_add(a.S + (size_t)hi * a.ldS + lo, -c);
          }
        }
      }
    }
    if (!has_next) break;
    pn = pn_next;
  }
  block_sum_to(cost, a.scalars + 0 /*COST_LIN*/, sred);
}

// ---- the tail of the iteration for panels -------------------------------------------------------
// dx_p = V^-1 (b_p - sum_a W_a^T dx_a) with W_a^T dx_a = R^T (M e), e = d rho + B d phi RECOMPUTED from the
// observation, the pose and the point at the linearisation point (poses_prev = table before retraction, pts = not
// yet updated) -- ~150 flop per observation instead of a 144-byte read of W; then p <- p + dx_p, ||dx_p||^2, and the
// cost at the new point with the retracted poses.
//
// Warp-autonomous, no shared memory, no barrier, no shuffle: the unit of work is HALF a panel, lane = landmark,
// and the lane walks the panel's rows (poses) itself -- the observations of a row are one coalesced 256-byte read
// per array, the pose and its update are warp-uniform (L1 broadcasts), the sum over the poses of a landmark stays in
// registers.  All observations of a unit are requested before the first is used; units are dealt round-robin to the
// resident warps.
constexpr int kFinishThreads = 128;
constexpr int kUnitsPerPanel = kPanelLm / 32;

template <int kLoss>
__global__ void __launch_bounds__(kFinishThreads, 3) panel_finish_kernel(const PanelArgs a) {
  __shared__ double sred[2 * (kFinishThreads / 32)];
  // per warp: the poses of the unit's rows at the linearisation point [8][12], after the update [8][12], and the
  // pose updates of the variable rows [8][6] -- staged once per unit with lane-parallel loads (one latency), then
  // read as broadcasts by the row loops (ncu: a quarter of the kernel's stall samples sat on the first use of a
  // pose loaded inside the loops, another sixth on the row descriptors)
  constexpr int kStage = 2 * 12 * kPanelRows + 6 * kPanelRows;
  __shared__ double sStage[kFinishThreads / 32][kStage];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_warps = gridDim.x * (kFinishThreads / 32);
  const int n_units = a.n_panels * kUnitsPerPanel;
  double* sP = sStage[warp];
  double* sN = sP + 12 * kPanelRows;
  double* sD = sN + 12 * kPanelRows;
  double cost = 0.0, dx2 = 0.0;
  static_assert(sizeof(PanelDesc) == 32 + 16 * kPanelRows && sizeof(PanelRow) == 16, "descriptor read as 16-byte words");

  for (int unit = blockIdx.x * (kFinishThreads / 32) + warp; unit < n_units; unit += n_warps) {
    const int pn = unit / kUnitsPerPanel;
    // descriptor: header by every lane (one broadcast), row r by lane r -- both loads in flight together; the rows
    // reach the other lanes through shuffles
    const uint4* dw = reinterpret_cast<const uint4*>(a.descs + pn);
    const uint4 hw = __ldg(dw);
    uint4 myrow = make_uint4(0u, 0u, 0u, 0u);
    if (lane < kPanelRows) myrow = __ldg(dw + 2 + lane);
    Panel pan;
    pan.lm_begin = (int)hw.x; pan.n_lms = (int)hw.y; pan.n_rows = (int)hw.z; pan.n_var = (int)hw.w;
    const int half = unit % kUnitsPerPanel;
    const int j = 32 * half + lane;
    if (32 * half >= pan.n_lms) continue;                  // warp-uniform
    const bool valid = j < pan.n_lms;
    const size_t q = (size_t)pan.lm_begin + (valid ? j : 0);
    unsigned row_mask[kPanelRows];
#pragma unroll
    for (int r = 0; r < kPanelRows; ++r) row_mask[r] = __shfl_sync(0xffffffffu, half ? myrow.w : myrow.z, r);
    // everything the unit reads from HBM, requested up front
    double ou[kPanelRows], ov[kPanelRows], od[kPanelRows];
    unsigned present = 0;
#pragma unroll
    for (int r = 0; r < kPanelRows; ++r) {
      ou[r] = ov[r] = od[r] = 0.0;
      if (r < pan.n_rows) {
        const unsigned m = row_mask[r];
        if (valid && ((m >> lane) & 1u)) {
          const double* cellp = a.pobs + (size_t)pn * kPanelObs + r * kPanelLm + j;
          ou[r] = ld_stream(cellp); ov[r] = ld_stream(cellp + kPanelRows * kPanelLm); od[r] = ld_stream(cellp + 2 * kPanelRows * kPanelLm);
          present |= 1u << r;
        }
      }
    }
    double X[3] = {0.0, 0.0, 0.0}, bp[3] = {0.0, 0.0, 0.0}, vi[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    if (valid) {
#pragma unroll
      for (int k = 0; k < 3; ++k) { X[k] = a.pts_in[3 * q + k]; bp[k] = ld_stream(a.Vg + 9 * q + 6 + k); }
#pragma unroll
      for (int k = 0; k < 6; ++k) vi[k] = ld_stream(a.Vinv + 6 * q + k);
    }
    // ---- poses and pose updates of the rows -> shared memory (all loads of a lane in flight together)
    __syncwarp();                                          // the previous unit's readers are done
    {
      double vP[3], vN[3], vD[2];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int e = lane + 32 * i, r = min(e / 12, kPanelRows - 1), k = e - 12 * (e / 12);
        const int pose = __shfl_sync(0xffffffffu, (int)myrow.x, r);
        const bool on = e < 12 * pan.n_rows;
        vP[i] = on ? __ldg(a.poses + 12 * (size_t)pose + k) : 0.0;
        vN[i] = on && a.eval_cost ? __ldg(a.poses_new + 12 * (size_t)pose + k) : 0.0;
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int e = lane + 32 * i, r = min(e / 6, kPanelRows - 1), k = e - 6 * (e / 6);
        const int off = __shfl_sync(0xffffffffu, (int)myrow.y, r);
        vD[i] = e < 6 * pan.n_var ? __ldg(a.dx_red + off + k) : 0.0;
      }
#pragma unroll
      for (int i = 0; i < 3; ++i) { sP[lane + 32 * i] = vP[i]; sN[lane + 32 * i] = vN[i]; }
#pragma unroll
      for (int i = 0; i < 2; ++i)
        if (lane + 32 * i < 6 * kPanelRows) sD[lane + 32 * i] = vD[i];
    }
    __syncwarp();
    // ---- sum over the rows of W^T dx_c
    double c0 = 0.0, c1 = 0.0, c2 = 0.0;
#pragma unroll
    for (int r = 0; r < kPanelRows; ++r) {
      if (r >= pan.n_var) break;                           // constant poses carry no update (rows: variable poses first)
      if (!((present >> r) & 1u)) continue;
      const ReprojGroup& grp = kLoss >= 0 ? a.g0 : a.groups[a.pgrp[((size_t)pn * kPanelRows + r) * kPanelLm + j]];
      double P[12], dxa[6], M[6], x, y, z;
#pragma unroll
      for (int k = 0; k < 12; ++k) P[k] = sP[12 * r + k];
#pragma unroll
      for (int k = 0; k < 6; ++k) dxa[k] = sD[6 * r + k];
      reproj_M<kLoss>(grp, P, X, ou[r], ov[r], od[r], M, x, y, z);
      // e = d rho + B d phi,  B = [[0, z, -y], [-z, 0, x], [y, -x, 0]]
      const double e0 = dxa[0] + z * dxa[4] - y * dxa[5];
      const double e1 = dxa[1] - z * dxa[3] + x * dxa[5];
      const double e2 = dxa[2] + y * dxa[3] - x * dxa[4];
      const double f0 = M[0] * e0 + M[1] * e1 + M[2] * e2;
      const double f1 = M[1] * e0 + M[3] * e1 + M[4] * e2;
      const double f2 = M[2] * e0 + M[4] * e1 + M[5] * e2;
      c0 += P[0] * f0 + P[3] * f1 + P[6] * f2;
      c1 += P[1] * f0 + P[4] * f1 + P[7] * f2;
      c2 += P[2] * f0 + P[5] * f1 + P[8] * f2;
    }
    // ---- back-substitution and retraction
    double Xn[3] = {0.0, 0.0, 0.0};
    if (valid) {
      const double s0 = bp[0] - c0, s1 = bp[1] - c1, s2 = bp[2] - c2;
      const double d0 = vi[0] * s0 + vi[1] * s1 + vi[2] * s2;
      const double d1 = vi[1] * s0 + vi[3] * s1 + vi[4] * s2;
      const double d2 = vi[2] * s0 + vi[4] * s1 + vi[5] * s2;
      a.dx_lm[3 * q] = d0; a.dx_lm[3 * q + 1] = d1; a.dx_lm[3 * q + 2] = d2;
      dx2 += d0 * d0 + d1 * d1 + d2 * d2;
      Xn[0] = X[0] + d0; Xn[1] = X[1] + d1; Xn[2] = X[2] + d2;
      a.pts[3 * q] = Xn[0]; a.pts[3 * q + 1] = Xn[1]; a.pts[3 * q + 2] = Xn[2];
    }
    // ---- cost at the new point (all rows, constant poses included)
    if (a.eval_cost) {
#pragma unroll
      for (int r = 0; r < kPanelRows; ++r) {
        if (r >= pan.n_rows) break;
        if (!((present >> r) & 1u)) continue;
        const ReprojGroup& grp = kLoss >= 0 ? a.g0 : a.groups[a.pgrp[((size_t)pn * kPanelRows + r) * kPanelLm + j]];
        double P[12], rr[3];
#pragma unroll
        for (int k = 0; k < 12; ++k) P[k] = sN[12 * r + k];
        reproj_residual_only(grp, P, Xn, ou[r], ov[r], od[r], rr);
#pragma unroll
        for (int k = 0; k < 3; ++k) cost += loss_rho_t<kLoss>(grp.loss, rr[k]);
      }
    }
  }
  cost = warp_sum(cost);
  dx2 = warp_sum(dx2);
  if (lane == 0) { sred[warp] = cost; sred[kFinishThreads / 32 + warp] = dx2; }
  __syncthreads();
  if (tid == 0) {
    double c = 0.0, d = 0.0;
#pragma unroll
    for (int w = 0; w < kFinishThreads / 32; ++w) { c += sred[w]; d += sred[kFinishThreads / 32 + w]; }
    if (a.eval_cost) red_add(a.scalars + 1 /*COST_NEW*/, c);
    red_add(a.scalars + 2 /*DX_NORM2*/, d);
  }
}

}  // namespace bs

// schur.cuh -- landmark elimination.  The reference solves the full sparse
// system with SuperLU (`splinalg.spsolve`, pyslam/problem.py:186); here the
// 3x3 landmark blocks are inverted in a batch, the reduced camera system
//     (U - sum_p W_p V_p^-1 W_p^T) dx_c = b_c - sum_p W_p V_p^-1 b_p
// is formed in the dense lower triangle of S, and after the dense solve the
// landmark updates follow from  dx_p = V_p^-1 (b_p - W_p^T dx_c).
#pragma once
#include "cholesky.cuh"
#include "common.cuh"
#include "reproj.cuh"
#include "lie.cuh"

namespace bs {

// Vg[q] = (xx,xy,xz,yy,yz,zz | b0,b1,b2)  ->  Vinv[q] = (xx,xy,xz,yy,yz,zz) of (V + lambda diag V)^-1
BS_D void sym3_inverse(const double* __restrict__ v, double lambda, double* __restrict__ o) {
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
  o[0] = c00 * id; o[1] = c01 * id; o[2] = c02 * id;
  o[3] = c11 * id; o[4] = c12 * id; o[5] = c22 * id;
}

// landmarks [q_begin, n_lm)
__global__ void __launch_bounds__(256) landmark_invert_kernel(int q_begin, int n_lm, const double* __restrict__ Vg,
                                                              double lambda, double* __restrict__ Vinv) {
  const int q = q_begin + blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_lm) return;
  sym3_inverse(Vg + 9 * (size_t)q, lambda, Vinv + 6 * (size_t)q);
}

struct SchurArgs {
  int n_obs;
  int n_lm;
  int obs_begin;                      // generic kernel: first observation it handles
  double lambda;
  int n_blocks;
  const LmBlock* __restrict__ blocks;
  const int* __restrict__ slot_pose;
  const unsigned* __restrict__ obs_code;        // [N] slot (bits 0-7, 255: constant pose) | block-local landmark (8-15)
  double* __restrict__ Vinv_out;
  const int* __restrict__ obs_pose;
  const int* __restrict__ obs_pt;
  const int* __restrict__ lm_start;   // [n_lm+1] CSR over landmarks
  const int* __restrict__ lm_obs;     // [N] CSR position -> observation index
  const int* __restrict__ pose_off;
  const double* __restrict__ W;       // tiled, see w_index()
  const double* __restrict__ Vg;      // [n_lm][9]
  const double* __restrict__ Vinv;    // [n_lm][6]
  double* __restrict__ S;
  int ldS;
  double* __restrict__ rhs;
  // static structure of the block kernel (built at finalize)
  const int* __restrict__ slot_off;         // reduced offset per slot entry
  const struct SchurPair* __restrict__ pairs;
  const unsigned* __restrict__ combos;      // runs: row observation | col observation << 8 | length << 16 (block-local)
  const struct SchurDesc* __restrict__ descs;  // [n_blocks]
  int max_lms, max_pairs, max_runs;         // shared-memory carve-up
};

// One thread per observation i: Y_i = W_i V^-1, then for every observation j of
// the same landmark whose pose block does not lie above i's:  S(i,j) -= Y_i W_j^T.
__global__ void __launch_bounds__(128) schur_generic_kernel(const SchurArgs a) {
  const int i = a.obs_begin + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n_obs) return;
  const int q = a.obs_pt[i];
  if (q >= a.n_lm) return;
  const int oi = a.pose_off[a.obs_pose[i]];
  if (oi < 0) return;
  const double* vi = a.Vinv + 6 * (size_t)q;
  const double m00 = vi[0], m01 = vi[1], m02 = vi[2], m11 = vi[3], m12 = vi[4], m22 = vi[5];
  double Y[18];
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    const double w0 = a.W[w_index(i, 3 * r)], w1 = a.W[w_index(i, 3 * r + 1)], w2 = a.W[w_index(i, 3 * r + 2)];
    Y[3 * r + 0] = w0 * m00 + w1 * m01 + w2 * m02;
    Y[3 * r + 1] = w0 * m01 + w1 * m11 + w2 * m12;
    Y[3 * r + 2] = w0 * m02 + w1 * m12 + w2 * m22;
  }
  const double* g = a.Vg + 9 * (size_t)q + 6;
  const double g0 = g[0], g1 = g[1], g2 = g[2];
#pragma unroll
  for (int r = 0; r < 6; ++r) red_add(a.rhs + oi + r, -(Y[3 * r] * g0 + Y[3 * r + 1] * g1 + Y[3 * r + 2] * g2));

  const int j0 = a.lm_start[q], j1 = a.lm_start[q + 1];
  for (int jj = j0; jj < j1; ++jj) {
    const int j = a.lm_obs[jj];
    const int oj = a.pose_off[a.obs_pose[j]];
    if (oj < 0 || oj > oi) continue;
    double wj[18];
#pragma unroll
    for (int k = 0; k < 18; ++k) wj[k] = a.W[w_index(j, k)];
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


// ---- fast path: one CTA per landmark block, Schur products on the fp64 tensor cores ----
// A landmark l seen by the poses (slots) a, b of its block contributes  S(a,b) -= Y_{l,a} W_{l,b}^T  with
// Y = W V_l^-1 (6x3).  The list of all such (observation of a, observation of b) "combos" is static: it is
// built once at finalize, grouped by slot pair, so a warp that owns a slot pair walks its combos and chains
// ONE mma.sync.m8n8k4.f64 per combo (rows = the 6 rows of Y, columns = the 6 rows of W, K = 3 of 4) into a
// register accumulator; one fp64 atomic per element of the 6x6 block leaves the SM per slot pair.
// b_c -= Y b_p rides along for free on the diagonal pairs: column 6 of the W operand holds b_p.
// Operands live in shared memory in MMA-fragment order, so a fragment load is one conflict-free LDS.64
// at  lane offset + observation * stride:
//     sY[o][3 r + c]           (stride 18; the K = 3 lane reads the next row's entry, times a zero of W)
//     sW[o][4 r + c], c = 3: 0, r = 6: b_p                     (stride 30: conflict-free 16-byte stores)
// Combos of a slot pair are stored as RUNS (row observation o, col observation o', length n) meaning the
// combos (o, o'), (o+1, o'+1), ...: the observations of a slot are sorted by landmark, so the landmarks two
// poses share are usually consecutive in both lists and a pair is one run; the fragment addresses of a run
// advance by constant strides, i.e. the inner loop is two LDS.64 at immediate offsets and one DMMA per combo.
struct SchurPair { unsigned slots_n; int rbeg; };   // row slot | col slot << 8 | #runs << 16;  first run (block-relative)
constexpr int kYS = 18, kWS = 30;

// Per-block descriptor of the Schur kernel (48 bytes: three 16-byte cp.async copies into a ring).
struct SchurDesc {
  int obs_begin, n_obs, lm_begin, n_lms;
  int slot_begin, n_slots, pair_begin, n_pairs;
  int run_begin, n_runs, pad0, pad1;
};

#ifndef BSLAM_SCHUR_CTAS
#define BSLAM_SCHUR_CTAS 4
#endif
constexpr int kSchurCtas = BSLAM_SCHUR_CTAS;
constexpr int kSchurThreads = 2 * kBlkObs;     // two threads per observation (three rows of W / Y each), 8 MMA warps
// small per-block tables, double-buffered: [slot offsets (int) | pairs | runs | V_p, b_p of the landmarks]
BS_HD size_t schur_tab_bytes(int max_lms, int max_pairs, int max_runs) {
  return ((sizeof(int) * kBlkObs + sizeof(SchurPair) * (size_t)max_pairs + sizeof(unsigned) * (size_t)max_runs +
           sizeof(double) * 9 * (size_t)max_lms + 7) & ~(size_t)7) + 8;
}
BS_HD size_t schur_smem_bytes(int max_lms, int max_pairs, int max_runs) {
  return sizeof(double) * ((size_t)kBlkObs * kYS + 8 + (size_t)kBlkObs * kWS + 8 + 10 * (size_t)max_lms) +
         2 * ((schur_tab_bytes(max_lms, max_pairs, max_runs) + 15) & ~(size_t)15);
}

// Persistent, software-pipelined: a CTA walks blocks b, b + grid, ...  The W tile and the observation codes
// of block i+1 are requested (into the registers block i has just finished with) before the MMA phase of
// block i starts; its small tables and landmark blocks V_p | b_p arrive by cp.async.  256 threads: the
// operand phase uses two threads per observation (register budget 64 -> 32 resident warps per SM, which is
// what the latency-bound MMA phase needs), the MMA phase deals the slot pairs to 8 warps.
__global__ void __launch_bounds__(kSchurThreads, kSchurCtas) schur_block_kernel(const SchurArgs a) {
  extern __shared__ __align__(16) double sm[];
  __shared__ __align__(16) SchurDesc sDesc[3];
  double* sY = sm;
  double* sW = sY + kBlkObs * kYS + 8;
  double* sVinv = sW + kBlkObs * kWS + 8;           // [6 max_lms]
  double* sG = sVinv + 6 * a.max_lms;               // [4 max_lms] (padded to pairs)
  char* sTab = reinterpret_cast<char*>(sG + 4 * a.max_lms);
  const size_t tab_bytes = (schur_tab_bytes(a.max_lms, a.max_pairs, a.max_runs) + 15) & ~(size_t)15;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ob = tid & (kBlkObs - 1), half = tid >> 7;      // observation of the block, rows 3 half .. 3 half + 2
  const int stride = gridDim.x;
  int b = blockIdx.x;
  if (b >= a.n_blocks) return;

  auto tab_off = [&](int buf) { return reinterpret_cast<int*>(sTab + buf * tab_bytes); };
  auto tab_pair = [&](int buf) { return reinterpret_cast<SchurPair*>(sTab + buf * tab_bytes + sizeof(int) * kBlkObs); };
  auto tab_run = [&](int buf) {
    return reinterpret_cast<unsigned*>(sTab + buf * tab_bytes + sizeof(int) * kBlkObs + sizeof(SchurPair) * (size_t)a.max_pairs);
  };
  auto tab_vg = [&](int buf) {
    const size_t o = sizeof(int) * kBlkObs + sizeof(SchurPair) * (size_t)a.max_pairs + sizeof(unsigned) * (size_t)a.max_runs;
    return reinterpret_cast<double*>(sTab + buf * tab_bytes + ((o + 7) & ~(size_t)7));
  };
  auto cp_async4 = [&](void* dst, const void* src) {
    const unsigned s_ = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s_), "l"(src) : "memory");
  };
  auto fetch_desc = [&](int blk_id, int slot) {
    if (tid < 3 && blk_id < a.n_blocks)
      cp_async16(reinterpret_cast<char*>(&sDesc[slot]) + 16 * tid, reinterpret_cast<const char*>(a.descs + blk_id) + 16 * tid);
  };
  auto stage_tables = [&](const SchurDesc& d, int buf) {
    int* so = tab_off(buf); SchurPair* sp = tab_pair(buf); unsigned* sr = tab_run(buf); double* sv = tab_vg(buf);
    for (int e = tid; e < d.n_slots; e += kSchurThreads) cp_async4(so + e, a.slot_off + d.slot_begin + e);
    for (int e = tid; e < d.n_pairs; e += kSchurThreads) cp_async8(sp + e, a.pairs + d.pair_begin + e);
    for (int e = tid; e < d.n_runs; e += kSchurThreads) cp_async4(sr + e, a.combos + d.run_begin + e);
    for (int e = tid; e < 9 * d.n_lms; e += kSchurThreads) cp_async8(sv + e, a.Vg + 9 * (size_t)d.lm_begin + e);
  };
  unsigned code = 255u;
  double w10[10];                 // W values k = 8 half .. 8 half + 9 (five (k, k+1) pairs; k = 9 half .. 9 half + 8 are used)
  auto load_inputs = [&](const SchurDesc& d) {
    code = 255u;
    if (ob < d.n_obs) {
      const int i = d.obs_begin + ob;
      code = ld_stream(a.obs_code + i);
      const double* Wp = a.W + w_pair_base(i) + 2 * kWTile * (4 * half);
#pragma unroll
      for (int k = 0; k < 5; ++k) { const double2 t = ld_stream2(Wp + 2 * kWTile * k); w10[2 * k] = t.x; w10[2 * k + 1] = t.y; }
    }
  };

  // ---- prologue
  if (tid < 3) reinterpret_cast<int4*>(&sDesc[0])[tid] = reinterpret_cast<const int4*>(a.descs + b)[tid];
  if (tid < 8) sY[kBlkObs * kYS + tid] = 0.0;
  __syncthreads();
  SchurDesc blk = sDesc[0];
  fetch_desc(b + stride, 1);
  fetch_desc(b + 2 * stride, 2);
  stage_tables(blk, 0);
  cp_async_commit();
  load_inputs(blk);

  const int g = lane >> 2, t = lane & 3;
  const char* pY = reinterpret_cast<const char*>(sY + 3 * g + t);      // rows 6-7: discarded outputs
  const char* pW = reinterpret_cast<const char*>(sW + lane);           // 4 n + t == lane
  const int lane_S = g * a.ldS + 2 * t;
  int it = 0;
  for (;;) {
    const int buf = it & 1;
    const int bn = b + stride;
    const bool has_next = bn < a.n_blocks;
    cp_async_wait_all();
    __syncthreads();               // tables / V_p of this block and the next descriptor have landed;
                                   // the MMA phase of the previous block is over (sY / sW / sVinv are free)
    // ---- phase A1: V^-1 and b_p of the block's landmarks (one thread each; blocks hold <= 128 landmarks)
    if (tid < blk.n_lms) {
      const double* vg = tab_vg(buf) + 9 * tid;
      double vi[6];
      sym3_inverse(vg, a.lambda, vi);
      const int q = blk.lm_begin + tid;
#pragma unroll
      for (int k = 0; k < 6; ++k) { sVinv[6 * tid + k] = vi[k]; a.Vinv_out[6 * (size_t)q + k] = vi[k]; }
      sG[4 * tid] = vg[6]; sG[4 * tid + 1] = vg[7]; sG[4 * tid + 2] = vg[8]; sG[4 * tid + 3] = 0.0;
    }
    __syncthreads();
    // ---- phase A2: operands of this block, rows 3 half .. 3 half + 2 of W and Y
    {
      const int sl = code & 255;
      if (sl != 255) {
        const int l = (code >> 8) & 255;
        const double* vi = sVinv + 6 * l;
        const double m00 = vi[0], m01 = vi[1], m02 = vi[2], m11 = vi[3], m12 = vi[4], m22 = vi[5];
        double* yr = sY + ob * kYS + 9 * half;
        double2* wr = reinterpret_cast<double2*>(sW + ob * kWS + 12 * half);
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          // k = 9 half + 3 r + c  ->  w10 index k - 8 half = half + 3 r + c
          const double w0 = half ? w10[1 + 3 * r] : w10[3 * r];
          const double w1 = half ? w10[2 + 3 * r] : w10[3 * r + 1];
          const double w2 = half ? w10[3 + 3 * r] : w10[3 * r + 2];
          yr[3 * r] = w0 * m00 + w1 * m01 + w2 * m02;
          yr[3 * r + 1] = w0 * m01 + w1 * m11 + w2 * m12;
          yr[3 * r + 2] = w0 * m02 + w1 * m12 + w2 * m22;
          wr[2 * r] = make_double2(w0, w1);
          wr[2 * r + 1] = make_double2(w2, 0.0);
        }
        if (half) {
          const double2* gl = reinterpret_cast<const double2*>(sG + 4 * l);
          wr[6] = gl[0];           // operand row 6: b_p
          wr[7] = gl[1];
        }
      } else if (!half) {
        sY[ob * kYS] = 0.0;        // read (times zero) by the K = 3 lane of the previous row
      }
    }
    // ---- everything the NEXT block needs goes in flight now (registers of this block are free)
    SchurDesc nblk = blk;
    if (has_next) {
      nblk = sDesc[(it + 1) % 3];
      load_inputs(nblk);
      stage_tables(nblk, buf ^ 1);
    }
    fetch_desc(b + 3 * stride, it % 3);
    cp_async_commit();
    __syncthreads();               // operands complete

    // ---- phase B: slot pairs, dealt round-robin to the warps (sorted by decreasing length at finalize)
    const int* sOff = tab_off(buf);
    const SchurPair* sPair = tab_pair(buf);
    const unsigned* sRun = tab_run(buf);
    for (int p = warp; p < blk.n_pairs; p += kSchurThreads / 32) {
      const SchurPair P = sPair[p];
      const int n_runs = P.slots_n >> 16;
      const unsigned* rl = sRun + P.rbeg;
      double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;
      for (int r = 0; r < n_runs; ++r) {
        const unsigned run = rl[r];
        const double* ya = reinterpret_cast<const double*>(pY + (run & 255u) * (kYS * 8));
        const double* wb = reinterpret_cast<const double*>(pW + ((run >> 8) & 255u) * (kWS * 8));
        int n = run >> 16;
        for (; n >= 4; n -= 4, ya += 4 * kYS, wb += 4 * kWS) {       // two independent accumulator chains
          dmma_8x8x4(c0, c1, ya[0], wb[0]);
          dmma_8x8x4(d0, d1, ya[kYS], wb[kWS]);
          dmma_8x8x4(c0, c1, ya[2 * kYS], wb[2 * kWS]);
          dmma_8x8x4(d0, d1, ya[3 * kYS], wb[3 * kWS]);
        }
        if (n & 2) {
          dmma_8x8x4(c0, c1, ya[0], wb[0]);
          dmma_8x8x4(d0, d1, ya[kYS], wb[kWS]);
          ya += 2 * kYS; wb += 2 * kWS;
        }
        if (n & 1) dmma_8x8x4(c0, c1, ya[0], wb[0]);
      }
      c0 += d0; c1 += d1;
      // C[g][2t], C[g][2t+1] = sum Y_row[g][.] W_col[2t(+1)][.]
      const int oa = sOff[P.slots_n & 255u], ob_ = sOff[(P.slots_n >> 8) & 255u];
      const bool diag = oa == ob_;
      if (g < 6) {
        if (t < 3) {
          double* Sd = a.S + ((size_t)oa * a.ldS + ob_ + lane_S);
          if (!diag || 2 * t <= g) red_add(Sd, -c0);
          if (!diag || 2 * t + 1 <= g) red_add(Sd + 1, -c1);
        } else if (diag) {
          red_add(a.rhs + oa + g, -c0);             // column 6: Y b_p
        }
      }
    }
    if (!has_next) break;
    b = bn; blk = nblk; ++it;
  }
  cp_async_wait_all();
}

struct BacksubArgs {
  int n_lm;
  int q_begin;                        // first landmark handled by the generic kernel
  int n_obs;
  int lm_off;                         // landmark slice of dx starts here
  const int* __restrict__ obs_pose;
  const int* __restrict__ lm_start;
  const int* __restrict__ lm_obs;
  const int* __restrict__ pose_off;
  const double* __restrict__ W;
  const double* __restrict__ Vg;
  const double* __restrict__ Vinv;
  double* __restrict__ dx;            // [reduced (padded) | 3 n_lm]
};

// One thread per landmark: dx_p = V^-1 (b_p - sum_j W_j^T dx_c(j)).
__global__ void __launch_bounds__(128) backsub_kernel(const BacksubArgs a) {
  const int q = a.q_begin + blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= a.n_lm) return;
  const double* g = a.Vg + 9 * (size_t)q + 6;
  double s0 = g[0], s1 = g[1], s2 = g[2];
  for (int jj = a.lm_start[q]; jj < a.lm_start[q + 1]; ++jj) {
    const int j = a.lm_obs[jj];
    const int oj = a.pose_off[a.obs_pose[j]];
    if (oj < 0) continue;
    const double* d = a.dx + oj;
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      const double dr = d[r];
      s0 -= a.W[w_index(j, 3 * r)] * dr;
      s1 -= a.W[w_index(j, 3 * r + 1)] * dr;
      s2 -= a.W[w_index(j, 3 * r + 2)] * dr;
    }
  }
  const double* vi = a.Vinv + 6 * (size_t)q;
  double* o = a.dx + a.lm_off + 3 * (size_t)q;
  o[0] = vi[0] * s0 + vi[1] * s1 + vi[2] * s2;
  o[1] = vi[1] * s0 + vi[3] * s1 + vi[4] * s2;
  o[2] = vi[2] * s0 + vi[4] * s1 + vi[5] * s2;
}

// ---- fused tail of the iteration for the landmark blocks ------------------------------
// back-substitution dx_p = V^-1 (b_p - W^T dx_c), retraction p <- p + dx_p, ||dx_p||^2 and the
// cost at the new point (pyslam/problem.py:155-156, 189-190, 400-409) in one pass over W and
// the observations.  The poses must already be retracted (retract_poses_kernel) and gathered
// per slot together with dx_c (retract_se3_slots_kernel).
struct FinishArgs {
  int n_obs;
  int lm_off;
  int eval_cost;
  int n_blocks;
  int max_slots;                        // shared-memory carve-up of the slot tables
  const LmBlock* __restrict__ blocks;
  const unsigned* __restrict__ obs_code;
  const unsigned char* __restrict__ lm_obs_local;
  const int* __restrict__ obs_pose;
  const ReprojGroup* __restrict__ groups;
  ReprojGroup g0;
  const int* __restrict__ lm_start;
  const double* __restrict__ obs_u;
  const double* __restrict__ obs_v;
  const double* __restrict__ obs_d;
  const double* __restrict__ poses;       // retracted
  const double* __restrict__ slot_poses;  // [n_slot_entries][12] retracted poses per slot entry
  const double* __restrict__ slot_dx;     // [n_slot_entries][6]  dx_c per slot entry
  double* __restrict__ pts;
  const double* __restrict__ W;
  const double* __restrict__ Vg;
  const double* __restrict__ Vinv;
  double* __restrict__ dx;
  double* __restrict__ scalars;
};

#ifndef BSLAM_FINISH_CTAS
#define BSLAM_FINISH_CTAS 4
#endif
constexpr int kFinishCtas = BSLAM_FINISH_CTAS;
BS_HD size_t finish_smem_bytes(int max_slots) { return 2 * sizeof(double) * 18 * (size_t)max_slots; }

// Persistent, software-pipelined like the assembly kernel: the W tile, the observations and the landmark
// data of block i+1 are requested into the registers block i has just finished with, the slot tables arrive
// by cp.async into the other half of a double buffer.  One atomic per CTA for the cost and for ||dx_p||^2.
template <int kLoss>
__global__ void __launch_bounds__(kBlkObs, kFinishCtas) lm_finish_kernel(const FinishArgs a) {
  extern __shared__ __align__(16) double sSlot[];      // 2 x [12 max_slots poses | 6 max_slots dx_c]
  __shared__ double sPts[3 * kBlkObs];
  __shared__ double sAcc[3 * kBlkObs];
  __shared__ double sred[2 * (kBlkObs / 32)];
  __shared__ __align__(16) LmBlock sDesc[3];
  __shared__ unsigned char sLmObs[kBlkObs];
  const int tid = threadIdx.x;
  const int stride = gridDim.x;
  int b = blockIdx.x;
  if (b >= a.n_blocks) return;
  const int tab = 18 * a.max_slots;

  auto fetch_desc = [&](int blk_id, int slot) {
    if (tid < 2 && blk_id < a.n_blocks)
      cp_async16(reinterpret_cast<char*>(&sDesc[slot]) + 16 * tid, reinterpret_cast<const char*>(a.blocks + blk_id) + 16 * tid);
  };
  auto stage_slots = [&](const LmBlock& d, int buf) {
    double* dp = sSlot + buf * tab;
    double* dd = dp + 12 * a.max_slots;
    const double* gp = a.slot_poses + 12 * (size_t)d.slot_begin;
    const double* gd = a.slot_dx + 6 * (size_t)d.slot_begin;
    if (a.eval_cost)
      for (int e = tid; e < 6 * d.n_slots; e += kBlkObs) cp_async16(dp + 2 * e, gp + 2 * e);
    for (int e = tid; e < 3 * d.n_slots; e += kBlkObs) cp_async16(dd + 2 * e, gd + 2 * e);
  };
  // registers that carry a block's inputs
  double w18[18];
  unsigned code = 255u;
  int lmobs = 0;
  double ou = 0.0, ov = 0.0, od = 0.0;
  double g0 = 0.0, g1 = 0.0, g2 = 0.0, vi[6] = {0, 0, 0, 0, 0, 0}, p0 = 0.0, p1 = 0.0, p2 = 0.0;
  int k0 = 0, k1 = 0;
  auto load_obs = [&](const LmBlock& d) {
    code = 255u;
    if (tid < d.n_obs) {
      const int i = d.obs_begin + tid;
      code = ld_stream(a.obs_code + i);
      lmobs = a.lm_obs_local[i];
      const double* Wp = a.W + w_pair_base(i);
#pragma unroll
      for (int k = 0; k < 9; ++k) { const double2 t = ld_stream2(Wp + 2 * kWTile * k); w18[2 * k] = t.x; w18[2 * k + 1] = t.y; }
    }
  };
  auto load_uvd = [&](const LmBlock& d) {
    if (a.eval_cost && tid < d.n_obs) {
      const int i = d.obs_begin + tid;
      ou = ld_stream(a.obs_u + i); ov = ld_stream(a.obs_v + i); od = ld_stream(a.obs_d + i);
    }
  };
  auto load_lm = [&](const LmBlock& d) {
    if (tid < d.n_lms) {
      const int q = d.lm_begin + tid;
      const double* g = a.Vg + 9 * (size_t)q + 6;
      g0 = ld_stream(g); g1 = ld_stream(g + 1); g2 = ld_stream(g + 2);
#pragma unroll
      for (int k = 0; k < 6; ++k) vi[k] = ld_stream(a.Vinv + 6 * (size_t)q + k);
      k0 = a.lm_start[q]; k1 = a.lm_start[q + 1];
      const double* P = a.pts + 3 * (size_t)q;
      p0 = P[0]; p1 = P[1]; p2 = P[2];
    }
  };

  // ---- prologue
  LmBlock blk = a.blocks[b];
  fetch_desc(b + stride, 1);
  fetch_desc(b + 2 * stride, 2);
  stage_slots(blk, 0);
  cp_async_commit();
  load_obs(blk);
  load_uvd(blk);
  load_lm(blk);

  double cost = 0.0, dx2 = 0.0;
  int it = 0;
  for (;;) {
    const int buf = it & 1;
    const int bn = b + stride;
    const bool has_next = bn < a.n_blocks;
    const double* sPose = sSlot + buf * tab;
    const double* sDx = sPose + 12 * a.max_slots;
    sLmObs[tid] = (unsigned char)lmobs;
    cp_async_wait_all();
    __syncthreads();                 // slot tables of this block, next descriptor; previous block fully done
    LmBlock nblk = blk;
    if (has_next) nblk = sDesc[(it + 1) % 3];
    const int sl = code & 255, ql = (code >> 8) & 255, grp_id = code >> 16;
    const int i = blk.obs_begin + tid;
    // ---- W^T dx_c per observation
    if (tid < blk.n_obs) {
      double c0 = 0.0, c1 = 0.0, c2 = 0.0;
      if (sl != 255) {
#pragma unroll
        for (int r = 0; r < 6; ++r) {
          const double d = sDx[6 * sl + r];
          c0 = fma(w18[3 * r], d, c0); c1 = fma(w18[3 * r + 1], d, c1); c2 = fma(w18[3 * r + 2], d, c2);
        }
      }
      sAcc[3 * tid] = c0; sAcc[3 * tid + 1] = c1; sAcc[3 * tid + 2] = c2;
    }
    if (has_next) load_obs(nblk);    // W / code / landmark map of the next block (code, sl, ql of this block are copies)
    __syncthreads();
    // ---- per landmark: back-substitution, retraction
    if (tid < blk.n_lms) {
      const int q = blk.lm_begin + tid;
      double s0 = g0, s1 = g1, s2 = g2;
      for (int k = k0 - blk.obs_begin; k < k1 - blk.obs_begin; ++k) {
        const int o = sLmObs[k];
        s0 -= sAcc[3 * o]; s1 -= sAcc[3 * o + 1]; s2 -= sAcc[3 * o + 2];
      }
      const double d0 = vi[0] * s0 + vi[1] * s1 + vi[2] * s2;
      const double d1 = vi[1] * s0 + vi[3] * s1 + vi[4] * s2;
      const double d2 = vi[2] * s0 + vi[4] * s1 + vi[5] * s2;
      double* o = a.dx + a.lm_off + 3 * (size_t)q;
      o[0] = d0; o[1] = d1; o[2] = d2;
      dx2 += d0 * d0 + d1 * d1 + d2 * d2;
      const double n0 = p0 + d0, n1 = p1 + d1, n2 = p2 + d2;
      sPts[3 * tid] = n0; sPts[3 * tid + 1] = n1; sPts[3 * tid + 2] = n2;
      double* P = a.pts + 3 * (size_t)q;
      P[0] = n0; P[1] = n1; P[2] = n2;
    }
    if (has_next) load_lm(nblk);
    __syncthreads();
    // ---- cost at the new point
    if (a.eval_cost && tid < blk.n_obs) {
      const ReprojGroup& g = kLoss >= 0 ? a.g0 : a.groups[grp_id];
      double P[12];
      if (sl != 255) {
        const double2* Ps = reinterpret_cast<const double2*>(sPose + 12 * sl);
#pragma unroll
        for (int k = 0; k < 6; ++k) { const double2 t = Ps[k]; P[2 * k] = t.x; P[2 * k + 1] = t.y; }
      } else {
        const double* Pg = a.poses + 12 * (size_t)a.obs_pose[i];
#pragma unroll
        for (int k = 0; k < 12; ++k) P[k] = Pg[k];
      }
      double r[3];
      reproj_residual_only(g, P, sPts + 3 * ql, ou, ov, od, r);
#pragma unroll
      for (int k = 0; k < 3; ++k) cost += loss_rho_t<kLoss>(g.loss, r[k]);
    }
    if (has_next) {
      load_uvd(nblk);
      stage_slots(nblk, buf ^ 1);
    }
    fetch_desc(b + 3 * stride, it % 3);
    cp_async_commit();
    if (!has_next) break;
    b = bn; blk = nblk; ++it;
  }
  cp_async_wait_all();
  // two block sums -> two atomics per CTA
  cost = warp_sum(cost);
  dx2 = warp_sum(dx2);
  const int lane = tid & 31, warp = tid >> 5;
  if (lane == 0) { sred[warp] = cost; sred[kBlkObs / 32 + warp] = dx2; }
  __syncthreads();
  if (tid == 0) {
    double c = 0.0, d = 0.0;
#pragma unroll
    for (int w = 0; w < kBlkObs / 32; ++w) { c += sred[w]; d += sred[kBlkObs / 32 + w]; }
    if (a.eval_cost) red_add(a.scalars + 1 /*COST_NEW*/, c);
    red_add(a.scalars + 2 /*DX_NORM2*/, d);
  }
}

}  // namespace bs

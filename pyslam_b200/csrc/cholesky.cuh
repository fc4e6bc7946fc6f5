// cholesky.cuh -- dense/tile-sparse fp64 Cholesky of the reduced camera system
// plus both triangular solves, as ONE persistent kernel.  Together with
// schur.cuh this replaces `splinalg.spsolve(precision, information)`
// (pyslam/problem.py:186).
//
// Layout: S is n_pad x n_pad row-major (n_pad = kNB * nt, kNB = 32, identity on the
// padding diagonal, lower triangle referenced); the right-hand side is stored as row
// n_pad of the same buffer, i.e. it is tile-row `nt` of a (nt+1) x nt grid of
// kNB x kNB tiles.  Carrying b as an extra row makes the forward substitution part
// of the factorisation:  [L; y^T] [L^T] = [S; b^T].
//
// Schedule: left-looking, one task per non-zero tile (i,j) of L, dispatched through an
// atomic ticket in DEPENDENCY-LEVEL order (solver.cu: build_chol_plan); the structure
// (which tiles are non-zero after fill-in) is computed once on the host from the
// co-visibility graph, so banded / sparse camera systems skip their zero tiles.
//     C   = S_ij - sum_{k<j} L_ik L_jk^T       fp64 tensor-core MMAs (DMMA m8n8k4), producers
//                                              consumed in readiness order, waiting on their flags
//     i==j: L_jj = chol(C), X_jj = L_jj^-1      (two 16x16 register-resident warp factorisations
//                                              whose idle half-warp forms the inverse in the same
//                                              pivot loop + DMMA for the off-diagonal block)
//     i>j : L_ij = C X_jj^T                      DMMA
// followed by nt backward-substitution tasks x_k = X_kk^T (y_k - sum_{i>k} L_ik^T x_i).
// The run time is the latency of the chain diag(k) -> L_ik -> diag(i) along the elimination
// tree, not flops: diagonal tasks therefore take their (up to two) critical producers
// through early-published C_ik tiles and form L_ik themselves (CholPlan::cll), and every
// transfer on that chain -- X_kk to the parent's diagonal task and to the column's
// off-diagonal tasks, the early C tiles, x_k in the backward substitution -- travels as
// (value, tag) RECORDS written and polled with single 16-byte accesses (the LL protocol
// of NCCL; tag = launch epoch XOR the value's bits, so that even a torn 16-byte access
// cannot pair a stale value with a fresh tag): one store and one polling load per hop
// instead of store, fence, flag store, flag poll, load.  A diagonal task whose two children finish together takes both in one
// pass.  (C4: 68 -> 54 us, profiles/r2_chol_micro.md.)
// The other dependencies are flags in global memory (st.release / ld.acquire at gpu scope) compared
// against a launch epoch the kernel advances itself (no memsets between launches); all
// CTAs are co-resident and take tickets in increasing order of a topological order, so a
// waiting CTA always waits on a ticket held by a running CTA (no deadlock).
//
// tcgen05 has no fp64 kind: DMMA (mma.sync.m8n8k4.f64) is the tensor path an
// fp64 factorisation can use on sm_100a.
#pragma once
#include "common.cuh"
#include "lie.cuh"

namespace bs {

#ifndef BSLAM_POLL_NS
#define BSLAM_POLL_NS 64     // back-off between two polls of a record
#endif
#ifndef BSLAM_TILE
#define BSLAM_TILE 32
#endif
constexpr int kNB = BSLAM_TILE;   // tile edge (32 or 64).  The factorisation of sparse (trajectory-like) reduced
                                  // systems is bound by the latency of the pivot chain along the elimination tree's
                                  // critical path, not by flops: 32-wide supernodes/separators halve that chain.
constexpr int kLd = kNB + 4;      // smem leading dimension (doubles): rows shift by 8 banks
constexpr int kCholThreads = 256;
constexpr int kNB16 = kNB / 16;   // 16x16 blocks per tile edge
static_assert(kNB == 32 || kNB == 64, "tile edge must be 32 or 64");
constexpr int kAccMI = kNB / 32;  // 8-row MMA blocks per warp   (warp w: rows (kNB/4) * (w / 2) ...)
constexpr int kAccNJ = kNB / 16;  // 8-col MMA blocks per warp   (        cols (kNB/2) * (w % 2) ...)

struct CholTask {
  int i, j;        // tile row / column (i == nt: right-hand-side row)
  int kbeg, kend;  // range in klist: columns k < j with L_ik and L_jk both non-zero
  int early;       // off-diagonal (i, j): e + 1 = also publish C_ij = S_ij - sum L_ik L_jk^T (before the triangular
                   //                    solve) into early slot e of row i
                   // diagonal (i, i): take the LAST `early` (0..kEarly) producer columns of klist through early C tiles
  int slot;       // multi-GPU: index of the tile in the packed exchange payload (structurally non-zero tiles of S
                   // before fill-in); -1: fill-in only (no initial data); -2: right-hand-side row
  int ranks;       // multi-GPU: bit r set = rank r's landmark shard can contribute to this tile; the other ranks' copies are
                   // structurally zero and are not read (time-contiguous shards touch ~1/N of the tiles each)
};

constexpr int kEarly = 2;
constexpr int kCholMaxPeers = 8;   // early C tiles per diagonal task (a separator has two children in the nested-dissection tree)

struct CholPlan {
  int nt;
  int n_tile_tasks;
  const CholTask* tasks;
  const int* klist;
  const int* bwd_ptr;    // [nt+1]
  const int* bwd_rows;   // rows i > k with a non-zero tile (i,k), descending
  int* ready;            // [(nt+1)*nt] epoch flags: tile final
  int* xready;           // [nt]       epoch flags: x_k final
  double* cll;           // [kEarly nt][kNB*kNB] records (value, launch epoch): early C tiles.  The critical chain is
                         //   diag(k) -> off(i,k): L_ik = C_ik X_kk^T -> diag(i): acc += L_ik L_ik^T
                         // with a global-memory hop after each arrow.  C_ik is known before X_kk, so the diagonal
                         // task i forms L_ik itself from the early C_ik and X_kk as soon as diag(k) has stored X_kk:
                         // one hop and one tile task leave the chain per tree level.
  double* xll;           // [nt][kNB*kNB] records (value, tag: st_rec) of X_kk = L_kk^-1 for the parents' diagonal tasks, and
  double* yll;           // [nt][kNB] records of x_k for the backward substitution: value and flag travel in ONE 16-byte
                         // store (the LL protocol of NCCL), so the hop  producer -> consumer  on the critical chain is one
                         // store and one (polling) load instead of store, fence, flag store, flag poll, load
  int* ticket;           // [0] ticket counter, [1] epoch of the last completed launch, [2] CTAs finished
                         // (the last CTA to finish resets [0], [2] and advances [1]: no memsets between launches)
  long long* trace;      // optional [n_tasks][4]: start, dependencies satisfied, end (globaltimer ns), SM id
  // Landmark-sharded iteration (peer.cuh): S = sum over the ranks of their partial Schur complements.  The sum is
  // never formed in memory: a tile task reads its own tile as  sum_r peer_pack[r][slot]  straight from the ranks'
  // exchange regions (mapped peer memory over NVLink), in rank order, and writes L into the private S.
  // Fused retraction (iteration path of pure panel problems): the backward task of tile k applies T <- exp(xi) T to the SE3
  // poses whose tangent block lies in tile k as soon as x_k is known, and adds ||x_k||^2 to scalars[DX_NORM2]; the
  // separate retraction kernel and its launch disappear from the iteration.  rt_poses == nullptr: off.
  double* rt_poses;            // [K][12] SE3 table
  const int* rt_pose_off;      // [K] reduced offsets
  const int* rt_tile_ptr;      // [nt + 1] CSR: poses per tile
  const int* rt_tile_pose;
  int rt_norm;                 // add ||x||^2 (the replicated reduced part is counted on shard 0 only)
  int world;             // 1: single GPU, the own tile comes from S
  int rhs_off;           // offset of the right-hand side inside a packed payload (= n_nz_tiles * kNB * kNB)
  const double* peer_pack[kCholMaxPeers];
  const double* mc_pack; // NVLS multicast address of the ranks' exchange regions (symmetric memory), or nullptr: one
                         // multimem.ld_reduce per element returns the sum over all ranks, reduced inside the NVSwitch
};

// sum over all ranks of the double at the multicast address (LDGMC.E.ADD.F64: the reduction happens in the switch)
BS_D double multimem_sum(const double* mc) {
  double v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.f64 %0, [%1];" : "=d"(v) : "l"(mc) : "memory");
  return v;
}

BS_D void dmma_8x8x4(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

BS_D long long gtime() {
  long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

BS_D int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
BS_D void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// (value, tag) records, 16-byte aligned, written and read with single 16-byte accesses; tag = epoch XOR the value's bits.
// A record is accepted when tag XOR value == epoch, which makes the protocol independent of the 16 bytes being transferred
// atomically: a torn read pairs the new value with the old tag (accepted only if that tag also fits the new value -- then
// the value IS the new one) or the old value with the new tag (accepted only if old value == new value).
BS_D void st_rec(double* rec, double value, int epoch) {
  const long long bits = __double_as_longlong(value);
  asm volatile("st.relaxed.gpu.global.v2.b64 [%0], {%1, %2};" ::"l"(rec), "l"(bits), "l"(bits ^ (long long)epoch) : "memory");
}
// Tiles of kNB x kNB records (kLower: only the lower triangle is stored) -> shared memory: every thread polls the
// records it needs, of both tiles when NT == 2, ALL in flight at once -- one L2 round trip after the producer's store lands
template <bool kLower>
BS_D void tile_to_records(double* R, const double* src, int epoch) {      // src: shared memory tile, leading dimension kLd
  for (int q = threadIdx.x; q < kNB * kNB / 2; q += kCholThreads) {
    const int r = q / (kNB / 2), c2 = (q % (kNB / 2)) << 1;
    if (!kLower || c2 <= r) {
      st_rec(R + 2 * (r * kNB + c2), src[r * kLd + c2], epoch);
      st_rec(R + 2 * (r * kNB + c2 + 1), src[r * kLd + c2 + 1], epoch);
    }
  }
}
template <int NT, bool kLower>
BS_D void tile_records(const double* R0, const double* R1, double* d0, double* d1, int epoch) {
  constexpr int kPer = kNB * kNB / 2 / kCholThreads;      // (row, column pair) slots per thread
  long long v[NT][kPer][2], e[NT][kPer][2];
  for (;;) {
    bool ok = true;
#pragma unroll
    for (int t = 0; t < NT; ++t)
#pragma unroll
      for (int h = 0; h < kPer; ++h) {
        const int q = threadIdx.x + h * kCholThreads;
        const int r = q / (kNB / 2), c2 = (q % (kNB / 2)) << 1;
        const double* rec = (t ? R1 : R0) + 2 * (r * kNB + c2);
#pragma unroll
        for (int w = 0; w < 2; ++w) {
          v[t][h][w] = 0; e[t][h][w] = epoch;      // not needed: passes the check below
          if (!kLower || c2 <= r)          // lower triangle (the entry above a diagonal element is stored as 0)
            asm volatile("ld.relaxed.gpu.global.v2.b64 {%0, %1}, [%2];" : "=l"(v[t][h][w]), "=l"(e[t][h][w]) : "l"(rec + 2 * w) : "memory");
        }
      }
#pragma unroll
    for (int t = 0; t < NT; ++t)
#pragma unroll
      for (int h = 0; h < kPer; ++h)
        ok = ok && (e[t][h][0] ^ v[t][h][0]) == (long long)epoch && (e[t][h][1] ^ v[t][h][1]) == (long long)epoch;
    if (ok) break;
    __nanosleep(BSLAM_POLL_NS);
  }
#pragma unroll
  for (int t = 0; t < NT; ++t)
#pragma unroll
    for (int h = 0; h < kPer; ++h) {
      const int q = threadIdx.x + h * kCholThreads;
      const int r = q / (kNB / 2), c2 = (q % (kNB / 2)) << 1;
      double* d = t ? d1 : d0;
      d[r * kLd + c2] = __longlong_as_double(v[t][h][0]);
      d[r * kLd + c2 + 1] = __longlong_as_double(v[t][h][1]);
    }
}

BS_D int rec_epoch(const double* rec) {
  long long v, e;
  asm volatile("ld.relaxed.gpu.global.v2.b64 {%0, %1}, [%2];" : "=l"(v), "=l"(e) : "l"(rec) : "memory");
  return (int)(e ^ v);
}
// N records at rec[stride * i]: all loads in flight at once, repeated until every epoch matches
template <int N>
BS_D void ld_recs(const double* rec, int stride, int epoch, double (&out)[N]) {
  long long v[N], e[N];
  for (;;) {
#pragma unroll
    for (int i = 0; i < N; ++i)
      asm volatile("ld.relaxed.gpu.global.v2.b64 {%0, %1}, [%2];" : "=l"(v[i]), "=l"(e[i]) : "l"(rec + (size_t)stride * i) : "memory");
    bool ok = true;
#pragma unroll
    for (int i = 0; i < N; ++i) ok = ok && (e[i] ^ v[i]) == (long long)epoch;
    if (ok) break;
    __nanosleep(BSLAM_POLL_NS);
  }
#pragma unroll
  for (int i = 0; i < N; ++i) out[i] = __longlong_as_double(v[i]);
}

// Block-wide wait until *flag == epoch.
BS_D void wait_flag(const int* flag, int epoch) {
  if (threadIdx.x == 0) {
    while (ld_acquire(flag) != epoch) __nanosleep(20);
  }
  __syncthreads();
}
// Block-wide publish: every thread's global writes become visible, then the flag.
BS_D void post_flag(int* flag, int epoch) {
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) st_release(flag, epoch);
}

// kNB x kNB accumulator of the CTA: 8 warps, warp w owns rows (kNB/4)*(w/2).., cols (kNB/2)*(w%2)..
// m8n8k4 lane mapping: a = A[g][t], b = B[g][t] (B^T operand), c = C[g][2t..2t+1],
// g = lane/4, t = lane%4.
struct TileAcc {
  double c[kAccMI][kAccNJ][2];
};

BS_D void acc_zero(TileAcc& acc) {
#pragma unroll
  for (int i = 0; i < kAccMI; ++i)
#pragma unroll
    for (int j = 0; j < kAccNJ; ++j) acc.c[i][j][0] = acc.c[i][j][1] = 0.0;
}

// acc += A(kNB x kNB) * B(kNB x kNB)^T, both in shared memory (leading dimension kLd)
BS_D void tile_mma_abt(const double* __restrict__ sA, const double* __restrict__ sB, TileAcc& acc) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const double* pa = sA + ((kNB / 4) * (warp >> 1) + g) * kLd + t;
  const double* pb = sB + ((kNB / 2) * (warp & 1) + g) * kLd + t;
#pragma unroll 4
  for (int k0 = 0; k0 < kNB; k0 += 4) {
    double a[kAccMI], b[kAccNJ];
#pragma unroll
    for (int i = 0; i < kAccMI; ++i) a[i] = pa[8 * i * kLd + k0];
#pragma unroll
    for (int j = 0; j < kAccNJ; ++j) b[j] = pb[8 * j * kLd + k0];
#pragma unroll
    for (int i = 0; i < kAccMI; ++i)
#pragma unroll
      for (int j = 0; j < kAccNJ; ++j) dmma_8x8x4(acc.c[i][j][0], acc.c[i][j][1], a[i], b[j]);
  }
}

// global tile (row-major, leading dimension ld) -> shared tile; rows >= nrows are zero-filled.
// L2-coherent loads (ld.global.cg): the data may have been produced by another CTA of this launch.
BS_D void tile_load(double* __restrict__ s, const double* __restrict__ gsrc, int ld, int nrows) {
  for (int e = threadIdx.x; e < kNB * kNB / 2; e += kCholThreads) {
    const int r = e / (kNB / 2), c2 = (e % (kNB / 2)) << 1;
    double2 v = make_double2(0.0, 0.0);
    if (r < nrows) v = __ldcg(reinterpret_cast<const double2*>(gsrc + (size_t)r * ld + c2));
    s[r * kLd + c2] = v.x;
    s[r * kLd + c2 + 1] = v.y;
  }
}

// accumulator <-> memory in the MMA C layout
template <typename F>
BS_D void acc_foreach(F f) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int m0 = (kNB / 4) * (warp >> 1), n0 = (kNB / 2) * (warp & 1);
#pragma unroll
  for (int i = 0; i < kAccMI; ++i)
#pragma unroll
    for (int j = 0; j < kAccNJ; ++j) f(i, j, m0 + 8 * i + g, n0 + 8 * j + 2 * t);
}

// 1/sqrt(d) to full double precision: hardware approximation (MUFU.RSQ64H, relative error e ~ 2^-22) + ONE third-order
// correction  y (1 + e/2 + 3 e^2 / 8),  e = 1 - d y^2  (error 5/16 e^3 ~ 2^-65): four dependent fp64 operations on the pivot
// critical path instead of the six of two Newton steps (and ~3x fewer than the library rsqrt()).
BS_D double fast_rsqrt(double d) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  const double e = fma(-(d * y), y, 1.0);
  const double p = fma(0.375, e, 0.5);
  return fma(y * e, p, y);
}

// ---- 16x16 building blocks (one warp each) ------------------------------------------------
// C(16x16) = alpha * A(16x16) * op(B) + beta * C by ONE warp (DMMA), shared memory, leading
// dimension kLd.  Safe when C aliases A or B: every operand is read before anything is stored.
template <bool kTransB>
BS_D void gemm16_warp(const double* __restrict__ A, const double* __restrict__ B, double* __restrict__ C, double alpha,
                      double beta) {
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  double c[2][2][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) c[i][j][0] = c[i][j][1] = 0.0;
#pragma unroll
  for (int k0 = 0; k0 < 16; k0 += 4) {
    const double a0 = A[g * kLd + k0 + t], a1 = A[(8 + g) * kLd + k0 + t];
    const double b0 = kTransB ? B[g * kLd + k0 + t] : B[(k0 + t) * kLd + g];
    const double b1 = kTransB ? B[(8 + g) * kLd + k0 + t] : B[(k0 + t) * kLd + 8 + g];
    dmma_8x8x4(c[0][0][0], c[0][0][1], a0, b0);
    dmma_8x8x4(c[0][1][0], c[0][1][1], a0, b1);
    dmma_8x8x4(c[1][0][0], c[1][0][1], a1, b0);
    dmma_8x8x4(c[1][1][0], c[1][1][1], a1, b1);
  }
  double old[2][2][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const double* p = C + (8 * i + g) * kLd + 8 * j + 2 * t;
      old[i][j][0] = beta != 0.0 ? p[0] : 0.0;
      old[i][j][1] = beta != 0.0 ? p[1] : 0.0;
    }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      double* p = C + (8 * i + g) * kLd + 8 * j + 2 * t;
      p[0] = alpha * c[i][j][0] + beta * old[i][j][0];
      p[1] = alpha * c[i][j][1] + beta * old[i][j][1];
    }
}

// Cholesky AND inverse of the 16x16 block at `s` in one pass of ONE warp.  Lanes 0..15 hold the rows of A
// (as potrf16_warp); lanes 16..31, idle there, hold the columns of X = L^-1 as right-hand sides e_c of
// L x = e_c: the column-j update  v[c] -= l * L[c][j]  is the same instruction for both halves (l = L[row][j]
// for a row of A, l = x_j = b[j] / L_jj for a column of X), so the inverse costs no extra latency on the
// pivot chain.  X goes to sX (lower triangular, zeros above the diagonal); 1/L_jj to srcp.
BS_D int potrf16_inv_warp(double* __restrict__ s, double* __restrict__ sX, double* __restrict__ scol,
                          double* __restrict__ srcp) {
  const int lane = threadIdx.x & 31, row = lane & 15;
  const bool hi = lane >= 16;
  double a[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) a[c] = hi ? (c == row ? 1.0 : 0.0) : s[row * kLd + c];
  int bad = 0;
  double d = __shfl_sync(0xffffffffu, a[0], 0);
  double r = fast_rsqrt(d);
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    if (!(d > 0.0)) ++bad;
    const double l = (lane == j) ? d * r : a[j] * r;
    a[j] = l;
    if (lane == j) srcp[j] = r;
    double dn = 0.0, rn = 0.0;
    if (j + 1 < 16) {
      dn = __shfl_sync(0xffffffffu, fma(-l, l, a[j + 1]), j + 1);
      rn = fast_rsqrt(dn);
    }
    double* sc = scol + (j & 1) * 16;
    if (!hi) sc[row] = l;
    __syncwarp();
#pragma unroll
    for (int c = j + 1; c < 16; ++c) a[c] = fma(-l, sc[c], a[c]);
    d = dn;
    r = rn;
  }
  if (!hi) {
#pragma unroll
    for (int c = 0; c < 16; ++c) s[row * kLd + c] = (c <= row) ? a[c] : 0.0;
  } else {
#pragma unroll
    for (int c = 0; c < 16; ++c) sX[c * kLd + row] = a[c];       // X[c][row]
  }
  return bad;
}

// Factorise the kNB x kNB tile in sA (lower triangle) in place and build X = L^-1 in sX (lower
// triangular, zeros above the diagonal), blocked by 16: per block column a register-resident
// potrf16 + trtri16 on warp 0, then the panel and trailing updates as one 16x16 DMMA product
// per warp; the off-diagonal blocks of X follow by block back-substitution.  Returns #bad pivots.
BS_D int tile_potrf_inv(double* __restrict__ sA, double* __restrict__ sX, double* __restrict__ scol,
                        double* __restrict__ srcp, int* __restrict__ sbad) {
  const int warp = threadIdx.x >> 5;
  auto Ab = [&](int i, int j) { return sA + (16 * i) * kLd + 16 * j; };
  auto Xb = [&](int i, int j) { return sX + (16 * i) * kLd + 16 * j; };
  if (threadIdx.x == 0) *sbad = 0;
  __syncthreads();
  if constexpr (kNB16 == 2) {
    // 32x32: everything on the pivot chain runs on warp 0; warp 1 forms T = L10 X00 beside the trailing
    // update, warp 2 clears the upper block of X.  The upper block (0,1) of A is scratch (never stored).
    const int lane = threadIdx.x & 31;
    double* T = Ab(0, 1);
    if (warp == 0) {
      const int bad = potrf16_inv_warp(Ab(0, 0), Xb(0, 0), scol, srcp);
      if (lane == 0 && bad) *sbad += bad;
      __syncwarp();
      gemm16_warp<true>(Ab(1, 0), Xb(0, 0), Ab(1, 0), 1.0, 0.0);               // L10 = A10 X00^T
    } else if (warp == 2) {
      for (int e = lane; e < 256; e += 32) Xb(0, 1)[(e >> 4) * kLd + (e & 15)] = 0.0;
    }
    __syncthreads();
    if (warp == 0) {
      gemm16_warp<true>(Ab(1, 0), Ab(1, 0), Ab(1, 1), -1.0, 1.0);              // A11 -= L10 L10^T
      __syncwarp();
      const int bad = potrf16_inv_warp(Ab(1, 1), Xb(1, 1), scol, srcp + 16);
      if (lane == 0 && bad) *sbad += bad;
    } else if (warp == 1) {
      gemm16_warp<false>(Ab(1, 0), Xb(0, 0), T, 1.0, 0.0);                     // T = L10 X00
    }
    __syncthreads();
    if (warp == 0) gemm16_warp<false>(Xb(1, 1), T, Xb(1, 0), -1.0, 0.0);       // X10 = -X11 T
    __syncthreads();
    return *sbad;
  } else {
  for (int b = 0; b < kNB16; ++b) {
    if (warp == 0) {
      const int bad = potrf16_inv_warp(Ab(b, b), Xb(b, b), scol, srcp + 16 * b);
      if ((threadIdx.x & 31) == 0 && bad) *sbad += bad;
    }
    __syncthreads();
    if (b == kNB16 - 1) break;
    // panel: L_ib = A_ib X_bb^T
    if (warp < kNB16 - 1 - b) gemm16_warp<true>(Ab(b + 1 + warp, b), Xb(b, b), Ab(b + 1 + warp, b), 1.0, 0.0);
    __syncthreads();
    // trailing update: A_ij -= L_ib L_jb^T for b < j <= i < kNB16  (at most 6 blocks, one per warp)
    {
      int w = 0;
      for (int i = b + 1; i < kNB16; ++i)
        for (int j = b + 1; j <= i; ++j, ++w)
          if (w == warp) gemm16_warp<true>(Ab(i, b), Ab(j, b), Ab(i, j), -1.0, 1.0);
    }
    __syncthreads();
  }
  // off-diagonal blocks of X by distance from the diagonal; the mirror block (j, i) is scratch
  for (int dist = 1; dist < kNB16; ++dist) {
    if (warp < kNB16 - dist) {
      const int j = warp, i = warp + dist;
      double* T = Xb(j, i);
      gemm16_warp<false>(Ab(i, j), Xb(j, j), T, 1.0, 0.0);                       // L_ij X_jj
      for (int k = j + 1; k < i; ++k) {
        __syncwarp();
        gemm16_warp<false>(Ab(i, k), Xb(k, j), T, 1.0, 1.0);                     // + L_ik X_kj
      }
      __syncwarp();
      gemm16_warp<false>(Xb(i, i), T, Xb(i, j), -1.0, 0.0);                      // X_ij = -X_ii T
    }
    __syncthreads();
  }
  for (int e = threadIdx.x; e < kNB * kNB; e += kCholThreads) {                   // zero the scratch (upper blocks)
    const int r = e / kNB, c = e % kNB;
    if ((c >> 4) > (r >> 4)) sX[r * kLd + c] = 0.0;
  }
  __syncthreads();
  return *sbad;
  }
}

// ---- the persistent kernel ---------------------------------------------------------
__global__ void __launch_bounds__(kCholThreads, 1)
chol_solve_kernel(double* __restrict__ S, int ld, double* __restrict__ Linv, double* __restrict__ x,
                  double* __restrict__ scalars, const CholPlan p) {
  extern __shared__ double smem[];
  double* sA = smem;
  double* sB = smem + kNB * kLd;
  double* sA2 = smem + 2 * kNB * kLd;       // second pair of tiles: both early tiles of a diagonal task at once
  double* sB2 = smem + 3 * kNB * kLd;
  double* scol = smem + 4 * kNB * kLd;      // 64
  double* srcp = scol + 64;                 // 64
  double* sred = srcp + 64;                 // kCholThreads
  __shared__ int s_ticket, s_bad, s_epoch, s_pair;
  const int tid = threadIdx.x;
  const int nt = p.nt;
  if (tid == 0) s_epoch = ld_acquire(p.ticket + 1) + 1;
  __syncthreads();
  const int epoch = s_epoch;        // flags equal to this value are "ready" in this launch

  for (;;) {
    __syncthreads();
    if (tid == 0) s_ticket = atomicAdd(p.ticket, 1);
    __syncthreads();
    const int tk = s_ticket;
    if (tk >= p.n_tile_tasks + nt) break;

    long long t0 = 0, t1 = 0;
    if (p.trace && tid == 0) t0 = gtime();
    if (tk < p.n_tile_tasks) {
      // ------------------------------------------------ tile task (i, j)
      const CholTask task = p.tasks[tk];
      const int ti = task.i, tj = task.j;
      const int rows_i = (ti == nt) ? 1 : kNB;
      TileAcc acc;
      acc_zero(acc);
      // the task's own tile of S is final before the kernel starts: fetch it now, off the dependency chain
      double* Cij = S + (size_t)ti * kNB * ld + (size_t)tj * kNB;
      TileAcc own;
      if (p.world <= 1) {
        acc_foreach([&](int i, int j, int r, int c) {
          double2 v = make_double2(0.0, 0.0);
          if (r < rows_i) v = __ldcg(reinterpret_cast<const double2*>(Cij + (size_t)r * ld + c));
          own.c[i][j][0] = v.x; own.c[i][j][1] = v.y;
        });
      } else {
        // fused all-reduce: the tile of the summed matrix, read from every rank's partial sum over NVLink
        const size_t off0 = task.slot == -2 ? (size_t)p.rhs_off + (size_t)tj * kNB : (size_t)task.slot * kNB * kNB;
        if (p.mc_pack) {
          acc_foreach([&](int i, int j, int r, int c) {
            const bool live = r < rows_i && task.slot != -1;
            const double* src = p.mc_pack + off0 + (size_t)r * kNB + c;
            own.c[i][j][0] = live ? multimem_sum(src) : 0.0;
            own.c[i][j][1] = live ? multimem_sum(src + 1) : 0.0;
          });
        } else
        acc_foreach([&](int i, int j, int r, int c) {
          double2 part[kCholMaxPeers];
          const bool live = r < rows_i && task.slot != -1;
#pragma unroll
          for (int q = 0; q < kCholMaxPeers; ++q) {
            part[q] = make_double2(0.0, 0.0);
            if (live && q < p.world && ((task.ranks >> q) & 1))
              part[q] = __ldcg(reinterpret_cast<const double2*>(p.peer_pack[q] + off0 + (size_t)r * kNB + c));
          }
          double2 v = part[0];
#pragma unroll
          for (int q = 1; q < kCholMaxPeers; ++q) { v.x += part[q].x; v.y += part[q].y; }
          own.c[i][j][0] = v.x; own.c[i][j][1] = v.y;
        });
      }
      const int n_early = (ti == tj) ? task.early : 0;
      const int kend_l = task.kend - n_early;
      for (int kk = task.kbeg; kk < kend_l; ++kk) {
        const int k = p.klist[kk];
        if (tid == 0) {                    // the two producers are polled by two warps at once
          while (ld_acquire(p.ready + ti * nt + k) != epoch) __nanosleep(20);
        } else if (tid == 32) {
          while (ld_acquire(p.ready + tj * nt + k) != epoch) __nanosleep(20);
        }
        __syncthreads();   // also protects sA/sB of the previous round
        tile_load(sA, S + (size_t)ti * kNB * ld + (size_t)k * kNB, ld, rows_i);
        if (ti != tj) tile_load(sB, S + (size_t)tj * kNB * ld + (size_t)k * kNB, ld, kNB);
        __syncthreads();
        tile_mma_abt(sA, ti != tj ? sB : sA, acc);
      }
      // last (critical) producer columns k: L_ik = C_ik X_kk^T formed here from the early C tile and the records of X_kk.
      // On the chain: ONE polling load per thread (all its records of X in flight at once), two MMAs, three barriers.
      if (n_early > 0) {
        const int k0 = p.klist[kend_l], k1 = p.klist[kend_l + n_early - 1];
        const double* R0 = p.xll + (size_t)k0 * kNB * kNB * 2;
        const double* R1 = p.xll + (size_t)k1 * kNB * kNB * 2;
        // the C tiles are normally there before the X tiles: into shared memory first
        __syncthreads();
        {
          const double* C0 = p.cll + (size_t)(kEarly * ti) * kNB * kNB * 2;
          if (n_early == 2) tile_records<2, false>(C0, C0 + (size_t)kNB * kNB * 2, sA, sA2, epoch);
          else tile_records<1, false>(C0, C0, sA, sA, epoch);
        }
        // a separator's two children usually finish together: take both X tiles in one pass unless the first child is clearly ahead
        bool together = false;
        if (n_early == 2) {
          if (tid == 0) {
            bool r0, r1;
            for (;;) {
              r0 = rec_epoch(R0) == epoch; r1 = rec_epoch(R1) == epoch;
              if (r0 || r1) break;
              __nanosleep(BSLAM_POLL_NS);
            }
            s_pair = (r0 && !r1) ? 0 : 1;
          }
          __syncthreads();
          together = s_pair != 0;
        }
        auto form = [&](double* sC, double* sX) {       // acc += (C X^T)(C X^T)^T
          TileAcc t_acc;
          acc_zero(t_acc);
          tile_mma_abt(sC, sX, t_acc);
          __syncthreads();
          acc_foreach([&](int i, int j, int r, int c) {
            sC[r * kLd + c] = t_acc.c[i][j][0];
            sC[r * kLd + c + 1] = t_acc.c[i][j][1];
          });
          __syncthreads();
          tile_mma_abt(sC, sC, acc);
        };
        if (together) {
          tile_records<2, true>(R0, R1, sB, sB2, epoch);
          __syncthreads();
          TileAcc t0_acc, t1_acc;
          acc_zero(t0_acc); acc_zero(t1_acc);
          tile_mma_abt(sA, sB, t0_acc);
          tile_mma_abt(sA2, sB2, t1_acc);
          __syncthreads();
          acc_foreach([&](int i, int j, int r, int c) {
            sA[r * kLd + c] = t0_acc.c[i][j][0]; sA[r * kLd + c + 1] = t0_acc.c[i][j][1];
            sA2[r * kLd + c] = t1_acc.c[i][j][0]; sA2[r * kLd + c + 1] = t1_acc.c[i][j][1];
          });
          __syncthreads();
          tile_mma_abt(sA, sA, acc);
          tile_mma_abt(sA2, sA2, acc);
        } else {
          tile_records<1, true>(R0, R0, sB, sB, epoch);
          __syncthreads();
          form(sA, sB);
          if (n_early == 2) {
            tile_records<1, true>(R1, R1, sB2, sB2, epoch);
            __syncthreads();
            form(sA2, sB2);
          }
        }
      }
      __syncthreads();
      if (p.trace && tid == 0) t1 = gtime();
      // C = S_ij - acc  -> sA
      acc_foreach([&](int i, int j, int r, int c) {
        sA[r * kLd + c] = own.c[i][j][0] - acc.c[i][j][0];
        sA[r * kLd + c + 1] = own.c[i][j][1] - acc.c[i][j][1];
      });
      __syncthreads();
      if (ti != tj && task.early) {          // publish the early C tile for the diagonal task of row ti
        tile_to_records<false>(p.cll + (size_t)(kEarly * ti + task.early - 1) * kNB * kNB * 2, sA, epoch);
      }
      if (ti == tj) {
        const int bad = tile_potrf_inv(sA, sB, scol, srcp, &s_bad);
        if (bad && tid == 0) red_add(scalars + 3 /*CHOL_FAIL*/, (double)bad);
        {   // first out: the records the parent's diagonal task is polling
          tile_to_records<true>(p.xll + (size_t)tj * kNB * kNB * 2, sB, epoch);
        }
        double* Lk = Linv + (size_t)tj * kNB * kNB;
        for (int e = tid; e < kNB * kNB; e += kCholThreads) {
          const int r = e / kNB, c = e % kNB;
          if (c <= r) Cij[(size_t)r * ld + c] = sA[r * kLd + c];
          Lk[e] = sB[r * kLd + c];
        }
      } else {
        {
          const double* Rj = p.xll + (size_t)tj * kNB * kNB * 2;
          tile_records<1, true>(Rj, Rj, sB, sB, epoch);      // X_jj straight from the diagonal task's records
        }
        if (p.trace && tid == 0) t1 = gtime();
        __syncthreads();
        acc_zero(acc);
        tile_mma_abt(sA, sB, acc);            // L_ij = C X_jj^T
        acc_foreach([&](int i, int j, int r, int c) {
          if (r < rows_i)
            *reinterpret_cast<double2*>(Cij + (size_t)r * ld + c) = make_double2(acc.c[i][j][0], acc.c[i][j][1]);
        });
      }
      post_flag(p.ready + ti * nt + tj, epoch);
    } else {
      // ------------------------------------------------ backward substitution, column k
      const int k = nt - 1 - (tk - p.n_tile_tasks);
      constexpr int kG = kCholThreads / kNB, kR = kNB / kG;      // row groups, rows per group
      const int c = tid % kNB, grp = tid / kNB;
      // X_kk rows of this thread, prefetched: x_k[c] = sum_r X_kk[r][c] v[r]
      wait_flag(p.ready + k * nt + k, epoch);
      double xr[kR];
      {
        const double* X = Linv + (size_t)k * kNB * kNB + (size_t)(kR * grp) * kNB + c;
#pragma unroll
        for (int r = 0; r < kR; ++r) xr[r] = __ldcg(X + (size_t)r * kNB);
      }
      double part = 0.0;
      for (int q = p.bwd_ptr[k]; q < p.bwd_ptr[k + 1]; ++q) {
        const int i = p.bwd_rows[q];
        wait_flag(p.ready + i * nt + k, epoch);
        double lr[kR];
        const double* L = S + (size_t)(i * kNB + kR * grp) * ld + (size_t)k * kNB + c;
#pragma unroll
        for (int r = 0; r < kR; ++r) lr[r] = __ldcg(L + (size_t)r * ld);
        const double* xi = p.yll + 2 * ((size_t)i * kNB + kR * grp);      // records of x_i: polled, no flag
        double xv[kR];
        ld_recs<kR>(xi, 2, epoch, xv);
#pragma unroll
        for (int r = 0; r < kR; ++r) part = fma(lr[r], xv[r], part);
      }
      wait_flag(p.ready + nt * nt + k, epoch);     // y_k (row 0 of the right-hand-side tile)
      if (p.trace && tid == 0) t1 = gtime();
      sred[grp * kNB + c] = part;
      __syncthreads();
      if (grp == 0) {
        const double y = __ldcg(S + (size_t)nt * kNB * ld + (size_t)k * kNB + c);
        double sum = 0.0;
#pragma unroll
        for (int q = 0; q < kG; ++q) sum += sred[q * kNB + c];
        scol[c] = y - sum;
      }
      __syncthreads();
      double v = 0.0;
#pragma unroll
      for (int r = 0; r < kR; ++r) v = fma(xr[r], scol[kR * grp + r], v);
      __syncthreads();
      sred[grp * kNB + c] = v;
      __syncthreads();
      if (grp == 0) {
        double sum = 0.0;
#pragma unroll
        for (int q = 0; q < kG; ++q) sum += sred[q * kNB + c];
        st_rec(p.yll + 2 * ((size_t)k * kNB + c), sum, epoch);
        x[(size_t)k * kNB + c] = sum;
        scol[c] = sum;
      }
      post_flag(p.xready + k, epoch);
      if (p.rt_poses) {                      // off the dependency chain: the flag is already posted
        const int q0 = p.rt_tile_ptr[k], np = p.rt_tile_ptr[k + 1] - q0;
        if (tid < np) {
          const int pi = p.rt_tile_pose[q0 + tid];
          const int o = p.rt_pose_off[pi] - k * kNB;
          double xi[6];
#pragma unroll
          for (int e = 0; e < 6; ++e) xi[e] = scol[o + e];
          double* P = p.rt_poses + 12 * (size_t)pi;
          se3_store(P, se3_mul(se3_exp(xi), se3_load(P)));
        }
        if (p.rt_norm && tid >= 32 && tid < 64) {
          const double v = scol[tid - 32];
          const double ss = warp_sum(v * v);
          if (tid == 32 && ss != 0.0) red_add(scalars + 2 /*DX_NORM2*/, ss);
        }
      }
    }
    if (p.trace && tid == 0) {
      unsigned smid;
      asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
      long long* tr = p.trace + 4 * (size_t)tk;
      tr[0] = t0; tr[1] = t1; tr[2] = gtime(); tr[3] = smid;
    }
  }
  // the last CTA out re-arms the control block for the next launch
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(p.ticket + 2, 1) == (int)gridDim.x - 1) {
      p.ticket[0] = 0;
      p.ticket[2] = 0;
      __threadfence();
      st_release(p.ticket + 1, epoch);
    }
  }
}

constexpr size_t kCholSmem = (4 * kNB * kLd + 64 + 64 + kCholThreads) * sizeof(double);

// Everything a linearisation needs cleared or refreshed, in ONE launch (each dependent launch inside the
// iteration's CUDA graph costs 2-3 us of latency):
//   blocks [0, n_tiles)   zero the structurally non-zero tiles; diagonal tiles get 1.0 on padding entries
//   remaining blocks      zero rhs | scalars, zero the V/b_p accumulators of long-track landmarks,
//                         gather the slot poses  slot_poses[e] = poses[slot_pose[e]]
struct PrepareArgs {
  double* S; int ld, nt, n_tiles;
  const int* tiles;
  const unsigned char* used;        // [n_pad] 1: a real unknown, 0: padding
  double* rhs; int n_rhs;           // rhs | scalars, contiguous
  double* vg_tail; int n_vg_tail;
  int n_slot_entries;
  const int* slot_pose;
  const double* poses;
  double* slot_poses;
  double* poses_prev; int n_prev;   // copy of the SE3 table at the linearisation point (panel_finish_kernel re-linearises there)
};
__global__ void __launch_bounds__(256) prepare_kernel(const PrepareArgs a) {
  if ((int)blockIdx.x < a.n_tiles) {
    const int id = a.tiles[blockIdx.x];
    const int ti = id / a.nt, tj = id % a.nt;
    double* T = a.S + (size_t)ti * kNB * a.ld + (size_t)tj * kNB;
    for (int e = threadIdx.x; e < kNB * kNB / 2; e += 256) {
      const int r = e / (kNB / 2), c2 = (e % (kNB / 2)) << 1;
      double2 v = make_double2(0.0, 0.0);
      if (ti == tj && (r >> 1) == (c2 >> 1) && !a.used[ti * kNB + r]) { if (r & 1) v.y = 1.0; else v.x = 1.0; }
      *reinterpret_cast<double2*>(T + (size_t)r * a.ld + c2) = v;
    }
    return;
  }
  const int nb = gridDim.x - a.n_tiles;
  const int t0 = (blockIdx.x - a.n_tiles) * 256 + threadIdx.x, stride = nb * 256;
  for (int e = t0; e < a.n_rhs; e += stride) a.rhs[e] = 0.0;
  for (int e = t0; e < a.n_vg_tail; e += stride) a.vg_tail[e] = 0.0;
  for (int e = t0; e < 12 * a.n_slot_entries; e += stride) a.slot_poses[e] = a.poses[12 * (size_t)a.slot_pose[e / 12] + e % 12];
  for (int e = t0; e < a.n_prev; e += stride) a.poses_prev[e] = a.poses[e];
}

// gather (unpack == 0) / scatter (unpack != 0) the listed tiles between S and a contiguous buffer
__global__ void __launch_bounds__(256) pack_tiles_kernel(double* __restrict__ S, int ld, int nt, const int* __restrict__ tiles,
                                                         double* __restrict__ pack, int unpack) {
  const int id = tiles[blockIdx.x];
  double* T = S + (size_t)(id / nt) * kNB * ld + (size_t)(id % nt) * kNB;
  double* P = pack + (size_t)blockIdx.x * kNB * kNB;
  for (int e = threadIdx.x; e < kNB * kNB / 2; e += 256) {
    const int r = e / (kNB / 2), c2 = (e % (kNB / 2)) << 1;
    double2* t2 = reinterpret_cast<double2*>(T + (size_t)r * ld + c2);
    double2* p2 = reinterpret_cast<double2*>(P + 2 * (size_t)e);
    if (unpack) *t2 = *p2;
    else *p2 = *t2;
  }
}

// S_ii *= (1 + lambda) for i < n  (LM damping lambda * diag(H); extension)
__global__ void damp_diag_kernel(double* __restrict__ S, int ld, int n, double lambda) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) S[(size_t)i * ld + i] *= (1.0 + lambda);
}

}  // namespace bs

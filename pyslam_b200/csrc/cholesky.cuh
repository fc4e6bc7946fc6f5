// cholesky.cuh -- dense fp64 Cholesky of the reduced camera system and the two
// triangular solves.  Together with schur.cuh this replaces
// `splinalg.spsolve(precision, information)` (pyslam/problem.py:186).
//
// S is n_pad x n_pad row-major (n_pad a multiple of NB = 64, padding rows carry
// an identity diagonal), only the lower triangle is referenced.  Right-looking
// blocked factorisation, two launches per 64-column panel:
//   chol_panel_kernel  every CTA re-factorises the 64x64 diagonal tile in shared
//                      memory and forms its triangular inverse; CTA 0 stores L_kk
//                      and L_kk^-1, CTA b>0 turns tile (k+b,k) into L = A L_kk^-T
//                      with fp64 tensor-core MMAs (mma.sync.m8n8k4.f64 -> DMMA);
//   chol_update_kernel A_ij -= L_ik L_jk^T for all tiles i >= j > k, DMMA.
// tcgen05 has no fp64 kind, so DMMA is the tensor path available to an fp64
// factorisation on sm_100a.
#pragma once
#include "common.cuh"

namespace bs {

constexpr int kNB = 64;           // panel width / tile edge
constexpr int kLd = 68;           // smem leading dimension (doubles): rows shift by 8 banks
constexpr int kCholThreads = 256;

BS_D void dmma_8x8x4(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// acc(64x64) = A(64x64) * B(64x64)^T, both tiles in shared memory with leading
// dimension kLd.  8 warps; warp w owns rows 16*(w/2).., cols 32*(w%2)..
// Lane mapping of m8n8k4: a = A[g][t], b = B^T[t][g] = B[g][t], c = C[g][2t..2t+1]
// with g = lane/4, t = lane%4.
struct TileAcc {
  double c[2][4][2];
};

BS_D void tile_mma_abt(const double* __restrict__ sA, const double* __restrict__ sB, TileAcc& acc) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int m0 = 16 * (warp >> 1), n0 = 32 * (warp & 1);
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc.c[i][j][0] = acc.c[i][j][1] = 0.0;
#pragma unroll 4
  for (int k0 = 0; k0 < kNB; k0 += 4) {
    double a[2], b[4];
#pragma unroll
    for (int i = 0; i < 2; ++i) a[i] = sA[(m0 + 8 * i + g) * kLd + k0 + t];
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = sB[(n0 + 8 * j + g) * kLd + k0 + t];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) dmma_8x8x4(acc.c[i][j][0], acc.c[i][j][1], a[i], b[j]);
  }
}

// global (row-major, leading dimension ld) 64x64 tile <-> shared tile
BS_D void tile_load(double* __restrict__ s, const double* __restrict__ gsrc, int ld) {
  for (int e = threadIdx.x; e < kNB * kNB / 2; e += kCholThreads) {
    const int r = e >> 5, c2 = (e & 31) << 1;
    const double2 v = *reinterpret_cast<const double2*>(gsrc + (size_t)r * ld + c2);
    s[r * kLd + c2] = v.x;
    s[r * kLd + c2 + 1] = v.y;
  }
}

// Factorise the 64x64 tile in sA (lower triangle) in place, and write the
// inverse of the factor into sX (lower triangular, zeros above).  Returns the
// number of non-positive pivots seen (same value in every thread).
BS_D int tile_potrf_inv(double* __restrict__ sA, double* __restrict__ sX) {
  __shared__ int s_bad;
  const int tid = threadIdx.x;
  if (tid == 0) s_bad = 0;
  for (int j = 0; j < kNB; ++j) {
    __syncthreads();
    if (tid == 0) {
      const double p = sA[j * kLd + j];
      if (!(p > 0.0)) s_bad += 1;
      sA[j * kLd + j] = sqrt(p);
    }
    __syncthreads();
    const double d = sA[j * kLd + j];
    if (tid > j && tid < kNB) sA[tid * kLd + j] /= d;
    __syncthreads();
    const int rem = kNB - 1 - j;
    for (int e = tid; e < rem * rem; e += kCholThreads) {
      const int i = j + 1 + e / rem, c = j + 1 + e % rem;
      if (c <= i) sA[i * kLd + c] -= sA[i * kLd + j] * sA[c * kLd + j];
    }
  }
  __syncthreads();
  // X = L^-1: thread c owns column c (forward substitution)
  for (int e = tid; e < kNB * kLd; e += kCholThreads) sX[e] = 0.0;
  __syncthreads();
  if (tid < kNB) {
    const int c = tid;
    sX[c * kLd + c] = 1.0 / sA[c * kLd + c];
    for (int i = c + 1; i < kNB; ++i) {
      double s0 = 0.0, s1 = 0.0;
      int m = c;
      for (; m + 1 < i; m += 2) {
        s0 += sA[i * kLd + m] * sX[m * kLd + c];
        s1 += sA[i * kLd + m + 1] * sX[(m + 1) * kLd + c];
      }
      if (m < i) s0 += sA[i * kLd + m] * sX[m * kLd + c];
      sX[i * kLd + c] = -(s0 + s1) / sA[i * kLd + i];
    }
  }
  __syncthreads();
  return s_bad;
}

// Panel k: grid = (#tiles below k) + 1.
__global__ void __launch_bounds__(kCholThreads)
chol_panel_kernel(double* __restrict__ S, int ld, int k, double* __restrict__ Linv, double* __restrict__ scalars) {
  extern __shared__ double smem[];
  double* sA = smem;                 // diagonal tile -> L_kk
  double* sX = smem + kNB * kLd;     // L_kk^-1
  double* sP = smem + 2 * kNB * kLd; // panel tile
  double* Akk = S + (size_t)k * kNB * ld + (size_t)k * kNB;
  tile_load(sA, Akk, ld);
  const int b = blockIdx.x;
  if (b > 0) tile_load(sP, S + (size_t)(k + b) * kNB * ld + (size_t)k * kNB, ld);
  const int bad = tile_potrf_inv(sA, sX);
  if (b == 0) {
    if (bad && threadIdx.x == 0) red_add(scalars + 3 /*CHOL_FAIL*/, (double)bad);
    double* Lk = Linv + (size_t)k * kNB * kNB;
    for (int e = threadIdx.x; e < kNB * kNB; e += kCholThreads) {
      const int r = e >> 6, c = e & 63;
      if (c <= r) Akk[(size_t)r * ld + c] = sA[r * kLd + c];
      Lk[e] = sX[r * kLd + c];
    }
    return;
  }
  // L_ik = A_ik * L_kk^-T   ->  C[m][n] = sum_c A[m][c] X[n][c]
  TileAcc acc;
  tile_mma_abt(sP, sX, acc);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int m0 = 16 * (warp >> 1), n0 = 32 * (warp & 1);
  double* P = S + (size_t)(k + b) * kNB * ld + (size_t)k * kNB;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      double2 v = make_double2(acc.c[i][j][0], acc.c[i][j][1]);
      *reinterpret_cast<double2*>(P + (size_t)(m0 + 8 * i + g) * ld + n0 + 8 * j + 2 * t) = v;
    }
}

// Trailing update after panel k: tile (k+1+by, k+1+bx) -= L_(i,k) L_(j,k)^T, by >= bx.
__global__ void __launch_bounds__(kCholThreads)
chol_update_kernel(double* __restrict__ S, int ld, int k) {
  const int bx = blockIdx.x, by = blockIdx.y;
  if (bx > by) return;
  extern __shared__ double smem[];
  double* sA = smem;
  double* sB = smem + kNB * kLd;
  const int ti = k + 1 + by, tj = k + 1 + bx;
  tile_load(sA, S + (size_t)ti * kNB * ld + (size_t)k * kNB, ld);
  if (bx != by) tile_load(sB, S + (size_t)tj * kNB * ld + (size_t)k * kNB, ld);
  else sB = sA;
  __syncthreads();
  TileAcc acc;
  tile_mma_abt(sA, sB, acc);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int m0 = 16 * (warp >> 1), n0 = 32 * (warp & 1);
  double* C = S + (size_t)ti * kNB * ld + (size_t)tj * kNB;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      double2* p = reinterpret_cast<double2*>(C + (size_t)(m0 + 8 * i + g) * ld + n0 + 8 * j + 2 * t);
      double2 v = *p;
      v.x -= acc.c[i][j][0];
      v.y -= acc.c[i][j][1];
      *p = v;
    }
}

// Forward substitution step k (L y = b), grid = (#tiles below k) + 1:
//   every CTA recomputes y_k = L_kk^-1 b_k; CTA 0 stores it, CTA b>0 applies
//   b_(k+b) -= L_(k+b,k) y_k.
__global__ void __launch_bounds__(kNB) trsv_fwd_kernel(const double* __restrict__ S, int ld, int k,
                                                        const double* __restrict__ Linv,
                                                        double* __restrict__ bvec, double* __restrict__ y) {
  __shared__ double sb[kNB], sy[kNB];
  const int t = threadIdx.x, b = blockIdx.x;
  sb[t] = bvec[k * kNB + t];
  __syncthreads();
  const double* X = Linv + (size_t)k * kNB * kNB + (size_t)t * kNB;
  double v = 0.0;
  for (int c = 0; c <= t; ++c) v += X[c] * sb[c];
  sy[t] = v;
  __syncthreads();
  if (b == 0) { y[k * kNB + t] = v; return; }
  const double* Lr = S + (size_t)((k + b) * kNB + t) * ld + (size_t)k * kNB;
  double acc = 0.0;
#pragma unroll 8
  for (int c = 0; c < kNB; ++c) acc += Lr[c] * sy[c];
  bvec[(k + b) * kNB + t] -= acc;
}

// Backward substitution step k (L^T x = y), grid = k + 1:
//   every CTA recomputes x_k = L_kk^-T y_k; CTA 0 stores it, CTA b>0 applies
//   y_(b-1) -= L_(k,b-1)^T x_k.
__global__ void __launch_bounds__(kNB) trsv_bwd_kernel(const double* __restrict__ S, int ld, int k,
                                                        const double* __restrict__ Linv,
                                                        double* __restrict__ y, double* __restrict__ x) {
  __shared__ double sy[kNB], sx[kNB];
  const int t = threadIdx.x, b = blockIdx.x;
  sy[t] = y[k * kNB + t];
  __syncthreads();
  const double* X = Linv + (size_t)k * kNB * kNB;
  double v = 0.0;
  for (int r = t; r < kNB; ++r) v += X[(size_t)r * kNB + t] * sy[r];
  sx[t] = v;
  __syncthreads();
  if (b == 0) { x[k * kNB + t] = v; return; }
  const int j = b - 1;
  const double* Lt = S + (size_t)k * kNB * ld + (size_t)j * kNB + t;   // column t of tile (k,j)
  double acc = 0.0;
#pragma unroll 8
  for (int r = 0; r < kNB; ++r) acc += Lt[(size_t)r * ld] * sx[r];
  y[j * kNB + t] -= acc;
}

// identity on the padding diagonal so the padded factorisation is well defined
__global__ void pad_diag_kernel(double* __restrict__ S, int ld, int n, int n_pad) {
  const int i = n + blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_pad) S[(size_t)i * ld + i] = 1.0;
}

// S_ii *= (1 + lambda) for i < n  (LM damping lambda * diag(H); extension)
__global__ void damp_diag_kernel(double* __restrict__ S, int ld, int n, double lambda) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) S[(size_t)i * ld + i] *= (1.0 + lambda);
}

}  // namespace bs

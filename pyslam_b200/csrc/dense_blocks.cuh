// dense_blocks.cuh -- residual blocks evaluated by user Python code (the
// duck-typed plug-in surface of pyslam/problem.py:338-360).  The host uploads,
// per block, e' = sqrt(w) r and the row-major J' = sqrt(w) [J_1 | J_2 | ...];
// this kernel does the reference's HT.HT^T and -HT.e for those blocks
// (problem.py:329-333) directly into the dense reduced system.
#pragma once
#include "common.cuh"

namespace bs {

struct DenseArgs {
  int n_blocks;
  const int* __restrict__ row_ptr;   // [n_blocks+1] residual rows
  const int* __restrict__ col_ptr;   // [n_blocks+1] Jacobian columns
  const long long* __restrict__ j_ptr;  // [n_blocks+1] offsets into J
  const int* __restrict__ col_index; // reduced index of each column, -1 if constant
  const double* __restrict__ J;
  const double* __restrict__ e;
  double* __restrict__ S;
  int ldS;
  double* __restrict__ rhs;
};

constexpr int kDenseRowChunk = 2048;

// grid = (n_blocks, row chunks).  A warp owns one (c1, c2) column pair at a
// time and strides its lanes over the rows of the chunk.
__global__ void __launch_bounds__(256) dense_blocks_kernel(const DenseArgs a) {
  const int b = blockIdx.x;
  const int r0 = a.row_ptr[b] + blockIdx.y * kDenseRowChunk;
  const int r1 = min(a.row_ptr[b + 1], r0 + kDenseRowChunk);
  if (r0 >= r1) return;
  const int c0 = a.col_ptr[b], nc = a.col_ptr[b + 1] - c0;
  const int rows0 = a.row_ptr[b];
  const double* J = a.J + a.j_ptr[b];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int p = warp; p < nc * (nc + 1); p += nw) {
    const int c1 = p / (nc + 1), c2 = p % (nc + 1);   // c2 == nc -> right-hand side
    const int i1 = a.col_index[c0 + c1];
    if (i1 < 0) continue;
    if (c2 == nc) {
      double s = 0.0;
      for (int r = r0 + lane; r < r1; r += 32) s += J[(size_t)(r - rows0) * nc + c1] * a.e[r];
      s = warp_sum(s);
      if (lane == 0) red_add(a.rhs + i1, -s);
      continue;
    }
    const int i2 = a.col_index[c0 + c2];
    if (i2 < 0 || i2 > i1) continue;
    double s = 0.0;
    for (int r = r0 + lane; r < r1; r += 32)
      s += J[(size_t)(r - rows0) * nc + c1] * J[(size_t)(r - rows0) * nc + c2];
    s = warp_sum(s);
    if (lane == 0) red_add(a.S + (size_t)i1 * a.ldS + i2, s);
  }
}

}  // namespace bs

// common.cuh -- small device/host helpers shared by all kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define BS_HD __host__ __device__ __forceinline__
#define BS_D __device__ __forceinline__

namespace bs {

constexpr int kWarp = 32;
constexpr double kSmallAngle = 1e-8;   // np.isclose(x, 0.) with default tolerances

// fire-and-forget fp64 reduction into global memory (RED.E.ADD.F64)
BS_D void red_add(double* addr, double v) { atomicAdd(addr, v); }

// 1/a to within an ulp or two: hardware seed (MUFU.RCP64H) + two Newton steps, ~8 instructions instead of
// the ~25-instruction IEEE division sequence with its slow-path branch.  a must be finite and non-zero.
BS_D double fast_rcp(double a) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const double e = fma(-a, y, 1.0);
    y = fma(y, e, y);
  }
  return y;
}

BS_D double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum -> one atomic per block.  `smem` needs blockDim.x/32 doubles.
BS_D void block_sum_to(double v, double* dst, double* smem) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) smem[w] = v;
  __syncthreads();
  if (w == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    double t = lane < nw ? smem[lane] : 0.0;
    t = warp_sum(t);
    if (lane == 0 && t != 0.0) red_add(dst, t);
  }
}

// streaming (read-once) 64-bit load that does not pollute L1
BS_D double ld_stream(const double* p) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
BS_D unsigned ld_stream(const unsigned* p) {
  unsigned v;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
BS_D double2 ld_stream2(const double* p) {     // 16-byte aligned pair
  double2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}
BS_D int ld_stream(const int* p) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

// ---- layout of W = J_T^T w J_p (6x3 per observation, row-major k = 3r + c) ----------------------
// Observations are grouped in tiles of 32; inside a tile the 18 values are stored as nine planes of
// (k, k+1) pairs:  W[tile][k/2][obs % 32][k % 2].  A warp that owns 32 consecutive observations
// moves a plane with one 512-byte, 16-byte-per-lane access (STG.128 / LDG.128) at an immediate
// offset from one base pointer, instead of 18 strided 8-byte accesses.
constexpr int kWTile = 32;
constexpr int kWTileLen = 18 * kWTile;      // doubles per tile
BS_HD size_t w_index(int i, int k) {
  return (size_t)(i >> 5) * kWTileLen + (size_t)(k >> 1) * (2 * kWTile) + ((i & 31) << 1) + (k & 1);
}
BS_HD size_t w_pair_base(int i) { return (size_t)(i >> 5) * kWTileLen + ((i & 31) << 1); }   // + 64 * (k/2)
BS_HD size_t w_alloc_len(int n_obs) { return (size_t)((n_obs + kWTile - 1) / kWTile) * kWTileLen; }

}  // namespace bs

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
BS_D int ld_stream(const int* p) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

}  // namespace bs

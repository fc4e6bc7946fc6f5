// Micro-benchmark (one CTA, clock64) of the diagonal-tile factorisations of cholesky.cuh and of the latencies on their pivot chain.
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I pyslam_b200/csrc tools/chol_micro2.cu -o /tmp/chol_micro2
#include <cstdio>
#include <vector>
#include <cmath>
#include "cholesky.cuh"
using namespace bs;

__global__ void __launch_bounds__(kCholThreads, 1) micro(const double* A, double* out, long long* cyc) {
  extern __shared__ double smem[];
  double* sA = smem; double* sX = smem + kNB * kLd; double* scol = smem + 2 * kNB * kLd; double* srcp = scol + 64;
  __shared__ int sbad;
  const int tid = threadIdx.x;
  for (int rep = 0; rep < 3; ++rep) {
    for (int e = tid; e < kNB * kNB; e += kCholThreads) sA[(e / kNB) * kLd + (e % kNB)] = A[e];
    __syncthreads();
    long long t0 = clock64();
    tile_potrf_inv(sA, sX, scol, srcp, &sbad);
    long long t1 = clock64();
    if (tid == 0) cyc[0] = t1 - t0;
    __syncthreads();
  }
  // latency probes (one warp)
  if (tid < 32) {
    double x = A[tid] + 1.0, y = 1.0000001;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; ++i) x = fma(x, y, 1e-9);
    long long t1 = clock64();
    double z = x;
#pragma unroll 1
    for (int i = 0; i < 256; ++i) z = fast_rsqrt(z + 2.0);
    long long t2 = clock64();
    double w = z;
#pragma unroll 1
    for (int i = 0; i < 256; ++i) w = __shfl_sync(0xffffffffu, w, (i + 1) & 31) + 1.0;
    long long t3 = clock64();
    double q = w;
#pragma unroll 1
    for (int i = 0; i < 256; ++i) { asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(q) : "d"(q + 2.0)); }
    long long t4 = clock64();
    if (tid == 0) { cyc[2] = (t1 - t0) / 256; cyc[3] = (t2 - t1) / 256; cyc[4] = (t3 - t2) / 256; cyc[5] = (t4 - t3) / 256; }
    out[3 * kNB * kNB + tid] = x + z + w + q;
  }
  __syncthreads();
  for (int e = tid; e < kNB * kNB; e += kCholThreads) { out[e] = sA[(e / kNB) * kLd + (e % kNB)]; out[kNB * kNB + e] = sX[(e / kNB) * kLd + (e % kNB)]; }
}

int main() {
  const int n = kNB;
  std::vector<double> A(n * n), B(n * n);
  srand(1);
  for (auto& v : B) v = rand() / (double)RAND_MAX - 0.5;
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) { double s = (i == j) ? 8.0 : 0.0; for (int k = 0; k < n; ++k) s += B[i * n + k] * B[j * n + k]; A[i * n + j] = s; }
  double *dA, *dout; long long* dc;
  cudaMalloc(&dA, n * n * 8); cudaMalloc(&dout, (3 * n * n + 64) * 8); cudaMalloc(&dc, 8 * 8);
  cudaMemcpy(dA, A.data(), n * n * 8, cudaMemcpyHostToDevice);
  size_t smem = kCholSmem;
  cudaFuncSetAttribute(micro, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  micro<<<1, kCholThreads, smem>>>(dA, dout, dc);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
  long long c[8]; std::vector<double> out(3 * n * n);
  cudaMemcpy(c, dc, 64, cudaMemcpyDeviceToHost); cudaMemcpy(out.data(), dout, 2 * n * n * 8, cudaMemcpyDeviceToHost);
  printf("cycles: tile_potrf_inv (32x32, blocked by 16) %lld | per loop iteration: dependent DFMA %lld  fast_rsqrt(+add) %lld  shfl(+add) %lld  rsqrt.approx(+add) %lld\n",
         c[0], c[2], c[3], c[4], c[5]);
  double eL = 0, eX = 0;
  for (int i = 0; i < n; ++i) for (int j = 0; j <= i; ++j) {
    double s = 0; for (int k = 0; k <= j; ++k) s += out[i * n + k] * out[j * n + k];
    eL = fmax(eL, fabs(s - A[i * n + j]));
    double t = 0; for (int k = j; k <= i; ++k) t += out[n * n + i * n + k] * out[k * n + j];
    eX = fmax(eX, fabs(t - (i == j)));
  }
  printf("max |LL^T - A| = %.3e   max |X L - I| = %.3e\n", eL, eX);
  return 0;
}

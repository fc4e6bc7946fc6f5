// Micro-benchmark of the diagonal-tile building blocks of cholesky.cuh (cycles on one SM).
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I pyslam_b200/csrc tools/chol_micro.cu -o /tmp/chol_micro
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
    for (int e = tid; e < kNB * kNB; e += kCholThreads) sA[(e >> 6) * kLd + (e & 63)] = A[e];
    __syncthreads();
    long long t0 = clock64();
    if (tid < 32) potrf16_warp(sA, scol, srcp);
    __syncthreads();
    long long t1 = clock64();
    if (tid < 32) trtri16_warp(sA, srcp, sX);
    __syncthreads();
    long long t2 = clock64();
    if (tid < 32) gemm16_warp<true>(sA + 16 * kLd, sX, sX + 16 * kLd, 1.0, 0.0);
    __syncthreads();
    long long t3 = clock64();
    if (tid == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; }
    __syncthreads();
    for (int e = tid; e < kNB * kNB; e += kCholThreads) sA[(e >> 6) * kLd + (e & 63)] = A[e];
    __syncthreads();
    t0 = clock64();
    tile_potrf_inv(sA, sX, scol, srcp, &sbad);
    t1 = clock64();
    TileAcc acc; acc_zero(acc);
    tile_mma_abt(sA, sX, acc);
    __syncthreads();
    t2 = clock64();
    if (tid == 0) { cyc[3] = t1 - t0; cyc[4] = t2 - t1; }
    acc_foreach([&](int i, int j, int r, int c) { out[2 * 4096 + r * 64 + c] = acc.c[i][j][0]; out[2 * 4096 + r * 64 + c + 1] = acc.c[i][j][1]; });
  }
  for (int e = tid; e < kNB * kNB; e += kCholThreads) { out[e] = sA[(e >> 6) * kLd + (e & 63)]; out[4096 + e] = sX[(e >> 6) * kLd + (e & 63)]; }
}

int main() {
  const int n = 64;
  std::vector<double> A(n * n), B(n * n);
  srand(1);
  for (auto& v : B) v = rand() / (double)RAND_MAX - 0.5;
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) { double s = (i == j) ? 8.0 : 0.0; for (int k = 0; k < n; ++k) s += B[i * n + k] * B[j * n + k]; A[i * n + j] = s; }
  double *dA, *dout; long long* dc;
  cudaMalloc(&dA, n * n * 8); cudaMalloc(&dout, 3 * n * n * 8); cudaMalloc(&dc, 8 * 8);
  cudaMemcpy(dA, A.data(), n * n * 8, cudaMemcpyHostToDevice);
  size_t smem = (2 * kNB * kLd + 128) * sizeof(double);
  cudaFuncSetAttribute(micro, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  micro<<<1, kCholThreads, smem>>>(dA, dout, dc);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
  long long c[8]; std::vector<double> out(3 * n * n);
  cudaMemcpy(c, dc, 64, cudaMemcpyDeviceToHost); cudaMemcpy(out.data(), dout, 3 * n * n * 8, cudaMemcpyDeviceToHost);
  printf("cycles: potrf16 %lld  trtri16 %lld  gemm16 %lld | tile_potrf_inv %lld  tile_mma 64^3 %lld\n", c[0], c[1], c[2], c[3], c[4]);
  // check L L^T = A and X L = I
  double eL = 0, eX = 0;
  for (int i = 0; i < n; ++i) for (int j = 0; j <= i; ++j) {
    double s = 0; for (int k = 0; k <= j; ++k) s += out[i * n + k] * out[j * n + k];
    eL = fmax(eL, fabs(s - A[i * n + j]));
    double t = 0; for (int k = j; k <= i; ++k) t += out[4096 + i * n + k] * out[k * n + j];
    eX = fmax(eX, fabs(t - (i == j)));
  }
  printf("max |LL^T - A| = %.3e   max |X L - I| = %.3e\n", eL, eX);
  return 0;
}

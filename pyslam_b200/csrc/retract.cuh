// retract.cuh -- apply the update vector on the manifold.
// Replaces `_perturb_by_key` + liegroups `perturb` (pyslam/problem.py:155-156,
// 400-409): poses T <- exp(xi) T (left perturbation, xi = [rho; phi]),
// vector-space parameters p <- p + dp; and np.linalg.norm(dx) (problem.py:160).
#pragma once
#include "common.cuh"
#include "lie.cuh"

namespace bs {

template <int G>
__global__ void __launch_bounds__(128) retract_poses_kernel(int n, double* __restrict__ poses,
                                                            const int* __restrict__ off,
                                                            const double* __restrict__ dx) {
  using Gr = Group<G>;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int o = off[i];
  if (o < 0) return;
  double xi[Gr::kDof];
#pragma unroll
  for (int k = 0; k < Gr::kDof; ++k) xi[k] = dx[o + k];
  double* p = poses + (size_t)Gr::kStore * i;
  Gr::store(p, Gr::mul(Gr::exp(xi), Gr::load(p)));
}

// SO(3) parameters (9 doubles, row-major): R <- exp(phi) R  (liegroups SO3.perturb)
__global__ void __launch_bounds__(128) retract_so3_kernel(int n, double* __restrict__ rots, const int* __restrict__ off,
                                                          const double* __restrict__ dx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int o = off[i];
  if (o < 0) return;
  const double xi[6] = {0.0, 0.0, 0.0, dx[o], dx[o + 1], dx[o + 2]};
  SE3 A{};
  double* p = rots + 9 * (size_t)i;
#pragma unroll
  for (int k = 0; k < 9; ++k) A.R[k] = p[k];
  A.t[0] = A.t[1] = A.t[2] = 0.0;
  const SE3 B = se3_mul(se3_exp(xi), A);
#pragma unroll
  for (int k = 0; k < 9; ++k) p[k] = B.R[k];
}

// SE3 poses of a bundle-adjustment problem, one launch: threads [0, n) retract the pose table; threads
// [n, n + n_slot_entries) retract the per-slot copies the landmark-block kernels read (slot_poses holds the
// poses gathered at linearisation time, so the copies never read the table while it is being rewritten) and
// gather dx_c per slot; `prev` (optional) receives the pose table as it was; all threads also reduce ||dx_c||^2 over the first n_red entries of dx.
__global__ void __launch_bounds__(128) retract_se3_slots_kernel(int n, double* __restrict__ poses, const int* __restrict__ off,
                                                                const double* __restrict__ dx, int n_slot_entries,
                                                                const int* __restrict__ slot_off, double* __restrict__ slot_poses,
                                                                double* __restrict__ slot_dx, int n_red, double* __restrict__ dx_norm2,
                                                                double* __restrict__ prev) {
  using Gr = Group<3>;
  __shared__ double sred[4];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const int o = off[i];
    double* p = poses + 12 * (size_t)i;
    if (prev) {             // the panel finish kernel re-linearises at the poses before the update
#pragma unroll
      for (int k = 0; k < 12; ++k) prev[12 * (size_t)i + k] = p[k];
    }
    if (o >= 0) {
      double xi[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) xi[k] = dx[o + k];
      Gr::store(p, Gr::mul(Gr::exp(xi), Gr::load(p)));
    }
  } else if (i < n + n_slot_entries) {
    const int e = i - n;
    const int o = slot_off[e];
    double xi[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) { xi[k] = dx[o + k]; slot_dx[6 * (size_t)e + k] = xi[k]; }
    double* p = slot_poses + 12 * (size_t)e;
    Gr::store(p, Gr::mul(Gr::exp(xi), Gr::load(p)));
  }
  double s = 0.0;
  for (int k = i; k < n_red; k += gridDim.x * blockDim.x) s += dx[k] * dx[k];
  block_sum_to(s, dx_norm2, sred);
}

// entry e of a flat parameter array moves by dx[off[e]] (off < 0: constant)
__global__ void __launch_bounds__(256) retract_flat_kernel(int n, double* __restrict__ vals,
                                                           const int* __restrict__ off,
                                                           const double* __restrict__ dx) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int o = off[e];
  if (o >= 0) vals[e] += dx[o];
}

// landmarks: the first n_lm points, update slice starts at dx + n_red
__global__ void __launch_bounds__(256) retract_landmarks_kernel(int n3, double* __restrict__ pts,
                                                                const double* __restrict__ dxl) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n3) pts[e] += dxl[e];
}

__global__ void __launch_bounds__(256) sumsq_kernel(int n, const double* __restrict__ v, double* __restrict__ dst) {
  __shared__ double sred[8];
  double s = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) s += v[i] * v[i];
  block_sum_to(s, dst, sred);
}

__global__ void add_scalar_kernel(double* dst, double v) { *dst += v; }

}  // namespace bs

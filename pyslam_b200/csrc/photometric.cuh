// photometric.cuh -- dense direct image alignment (BASELINE config 5).
// Replaces PhotometricResidualSE3.evaluate
// (pyslam/residuals/photometric_residual.py:81-161) together with the helpers it
// calls per pixel -- StereoCamera.triangulate(+Jacobian column) / project(+Jacobian)
// / is_valid_measurement (pyslam/sensors/stereo_camera.py:90-174),
// bilinear_interpolate (pyslam/utils.py:16-77), stackmul (utils.py:80-89) and
// fast_se3_odot (photometric_residual.py:14-35) -- plus the IRLS scaling and the
// 6x6 / 6x1 normal-equation reduction of pyslam/problem.py:329-360.
//
// Per reference pixel the kernel streams 48 bytes (u, v, d, I_ref, dI/du, dI/dv);
// the triangulated point and the disparity column of the triangulation Jacobian are
// recomputed from (u, v, d) instead of being stored (the reference's layout streams
// 144 B/pixel).  The tracking image is gathered through L2 (4 taps per pixel).
#pragma once
#include "common.cuh"
#include "loss.cuh"

namespace bs {

struct PhotoArgs {
  int n_px;
  const double* __restrict__ uvd;     // [n][3] reference pixel grid + disparity (already filtered by the constructor)
  const double* __restrict__ im_ref;  // [n]
  const double* __restrict__ im_jac;  // [n][2]
  const double* __restrict__ im_track;  // [h][w]
  int w, h;
  double cu, cv, fu, fv, b;
  double intensity_covar, depth_covar;
  Loss loss;
  const double* __restrict__ R;       // rotation (9, row-major) and translation (3) of T_track_ref: the two halves of an
  const double* __restrict__ t;       // SE3 table entry, or an SO3 parameter + a 3-vector parameter ((SO3, t) form,
                                      // photometric_residual.py:83-84)
  int off_trans, off_rot;             // reduced offsets of the rho columns (J[:, 0:3]) and the phi columns (J[:, 3:6]),
                                      // -1: that parameter is constant (its columns are dropped, problem.py:343-356)
  double* __restrict__ S;
  int ldS;
  double* __restrict__ rhs;
  double* __restrict__ scalars;
};

// utils.py:27-77: truncate toward zero, weights before clamping, clamp = repeat border
BS_D double bilinear(const double* __restrict__ im, int w, int h, double x, double y) {
  int x0 = (int)x, y0 = (int)y;
  int x1 = x0 + 1, y1 = y0 + 1;
  const double wa = (x1 - x) * (y1 - y), wb = (x1 - x) * (y - y0);
  const double wc = (x - x0) * (y1 - y), wd = (x - x0) * (y - y0);
  x0 = min(max(x0, 0), w - 1); x1 = min(max(x1, 0), w - 1);
  y0 = min(max(y0, 0), h - 1); y1 = min(max(y1, 0), h - 1);
  return wa * im[(size_t)y0 * w + x0] + wb * im[(size_t)y1 * w + x0] + wc * im[(size_t)y0 * w + x1] +
         wd * im[(size_t)y1 * w + x1];
}

constexpr int kPhotoThreads = 256;

// kCostOnly: just sum rho(r) into scalars[slot]; otherwise also H (lower 6x6), b.
template <bool kCostOnly>
__global__ void __launch_bounds__(kPhotoThreads) photometric_kernel(const PhotoArgs a, int slot) {
  __shared__ double sred[28][kPhotoThreads / 32];
  double P[12];
#pragma unroll
  for (int k = 0; k < 9; ++k) P[k] = a.R[k];
#pragma unroll
  for (int k = 0; k < 3; ++k) P[9 + k] = a.t[k];
  const double fu_over_fv = a.fu / a.fv;
  const bool stereo = a.b > 0.0;       // b <= 0: RGB-D pinhole model
  double acc[28];            // 21 lower-triangle entries of J^T w J, 6 of -J^T w r, cost
#pragma unroll
  for (int k = 0; k < 28; ++k) acc[k] = 0.0;

  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n_px; i += gridDim.x * blockDim.x) {
    const double u = ld_stream(a.uvd + 3 * (size_t)i), v = ld_stream(a.uvd + 3 * (size_t)i + 1),
                 d = ld_stream(a.uvd + 3 * (size_t)i + 2);
    double X, Y, Z, tj0, tj1, tj2;       // point and the third column of the triangulation Jacobian
    if (stereo) {                        // stereo_camera.py:137-174
      const double b_over_d = a.b / d;
      const double b_over_d2 = b_over_d / d;
      X = (u - a.cu) * b_over_d; Y = (v - a.cv) * b_over_d * fu_over_fv; Z = a.fu * b_over_d;
      tj0 = (a.cu - u) * b_over_d2; tj1 = (a.cv - v) * b_over_d2 * fu_over_fv; tj2 = -a.fu * b_over_d2;
    } else {                             // rgbd_camera.py:149-180: d is the depth
      tj0 = (u - a.cu) / a.fu; tj1 = (v - a.cv) / a.fv; tj2 = 1.0;
      X = tj0 * d; Y = tj1 * d; Z = d;
    }
    // transform + project
    const double x = P[0] * X + P[1] * Y + P[2] * Z + P[9];
    const double y = P[3] * X + P[4] * Y + P[5] * Z + P[10];
    const double z = P[6] * X + P[7] * Y + P[8] * Z + P[11];
    const double iz = 1.0 / z, iz2 = iz * iz;
    const double ut = a.fu * x * iz + a.cu, vt = a.fv * y * iz + a.cv, dt = stereo ? a.fu * a.b * iz : z;
    // stereo_camera.py:90-97 (disparity compared with the image width); rgbd_camera.py:103-110 (depth > 0)
    const bool valid = (dt > 0.0) && (!stereo || dt < a.w) && (vt > 0.0) && (vt < a.h) && (ut > 0.0) && (ut < a.w);
    if (!valid) continue;
    const double r0 = bilinear(a.im_track, a.w, a.h, ut, vt) - ld_stream(a.im_ref + i);
    // image gradient (1x2) times the first two rows of the projection Jacobian -> 1x3
    const double gu = ld_stream(a.im_jac + 2 * (size_t)i), gv = ld_stream(a.im_jac + 2 * (size_t)i + 1);
    const double p0 = gu * (a.fu * iz), p1 = gv * (a.fv * iz);
    const double p2 = gu * (-a.fu * x * iz2) + gv * (-a.fv * y * iz2);
    // d r / d disparity = (p R) . triang_jac[:, 2]
    const double q0 = p0 * P[0] + p1 * P[3] + p2 * P[6];
    const double q1 = p0 * P[1] + p1 * P[4] + p2 * P[7];
    const double q2 = p0 * P[2] + p1 * P[5] + p2 * P[8];
    const double jd = q0 * tj0 + q1 * tj1 + q2 * tj2;
    const double stiff = 1.0 / sqrt(a.intensity_covar + a.depth_covar * jd * jd);
    const double r = stiff * r0;
    acc[27] += loss_rho(a.loss, r);
    if (kCostOnly) continue;
    const double wgt = loss_weight(a.loss, r);
    // J = stiff * p [I | -pt^]
    double J[6];
    J[0] = stiff * p0; J[1] = stiff * p1; J[2] = stiff * p2;
    J[3] = stiff * (p2 * y - p1 * z);
    J[4] = stiff * (p0 * z - p2 * x);
    J[5] = stiff * (p1 * x - p0 * y);
    int k = 0;
#pragma unroll
    for (int rr = 0; rr < 6; ++rr) {
      const double wj = wgt * J[rr];
#pragma unroll
      for (int c = 0; c <= rr; ++c) acc[k++] += wj * J[c];
      acc[21 + rr] -= wj * r;
    }
  }
  // block reduction of the 28 accumulators
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 28; ++k) {
    if (kCostOnly && k != 27) continue;
    const double v = warp_sum(acc[k]);
    if (lane == 0) sred[k][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < 28) {
    const int k = threadIdx.x;
    if (kCostOnly && k != 27) return;
    double v = 0.0;
#pragma unroll
    for (int w8 = 0; w8 < kPhotoThreads / 32; ++w8) v += sred[k][w8];
    if (k == 27) { if (v != 0.0) red_add(a.scalars + slot, v); return; }
    if (v == 0.0) return;
    // tangent index 0..5 = [rho; phi] -> position in the reduced system
    auto pos = [&](int q) { return q < 3 ? (a.off_trans < 0 ? -1 : a.off_trans + q) : (a.off_rot < 0 ? -1 : a.off_rot + q - 3); };
    if (k < 21) {
      // k -> (row, col) of the lower triangle, row-major
      int rr = 0, base = 0;
      while (base + rr + 1 <= k) { base += rr + 1; ++rr; }
      const int gr = pos(rr), gc = pos(k - base);
      if (gr < 0 || gc < 0) return;
      red_add(a.S + (size_t)max(gr, gc) * a.ldS + min(gr, gc), v);
    } else {
      const int g = pos(k - 21);
      if (g >= 0) red_add(a.rhs + g, v);
    }
  }
}

}  // namespace bs

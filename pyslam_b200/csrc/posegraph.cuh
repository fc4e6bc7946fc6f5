// posegraph.cuh -- unary pose priors and binary relative-pose factors on
// SE(2)/SE(3), one thread per block, accumulated straight into the dense
// reduced system (poses are never eliminated).
//
// Replaces PoseResidual.evaluate (pyslam/residuals/pose_residual.py:12-27):
//     r = S log(T T_obs^-1),                       J  = S
// and PoseToPoseResidual.evaluate (pyslam/residuals/pose_to_pose_residual.py:12-32):
//     r = S log(T2 (T1^-1 T21_obs^-1)),            J1 = -S Ad(T2 T1^-1),  J2 = S
// (the reference's approximate Jacobians), followed by the IRLS
// scaling and H += J^T w J, b -= J^T w r of pyslam/problem.py:349-360,329-333.
#pragma once
#include "common.cuh"
#include "lie.cuh"
#include "loss.cuh"

namespace bs {

struct EdgeArgs {
  int n;
  const int* __restrict__ i1;        // first pose (the only pose for unary priors)
  const int* __restrict__ i2;        // second pose, nullptr for unary priors
  const double* __restrict__ Tobs;   // [n][kStore]
  const double* __restrict__ stiff;  // [dof*dof] shared or [n][dof*dof]
  int stiff_per_block;
  Loss loss;
  const double* __restrict__ poses;  // [K][kStore]
  const int* __restrict__ pose_off;  // reduced offset or -1
  double* __restrict__ S;
  int ldS;
  double* __restrict__ rhs;
  double* __restrict__ scalars;
};

// H(ra.., ca..) += A^T diag(w) B  (A, B: D x D row-major Jacobians).
// lower_only: the block sits on the diagonal, add only c <= r.
template <int D>
BS_D void add_block(double* S, int ldS, int ro, int co, const double* A, const double* B, const double* w,
                    bool lower_only) {
#pragma unroll
  for (int r = 0; r < D; ++r)
#pragma unroll
    for (int c = 0; c < D; ++c) {
      if (lower_only && c > r) continue;
      double v = 0.0;
#pragma unroll
      for (int k = 0; k < D; ++k) v += w[k] * A[k * D + r] * B[k * D + c];
      red_add(S + (size_t)(ro + r) * ldS + co + c, v);
    }
}

template <int G, bool kBinary, bool kCostOnly>
__global__ void __launch_bounds__(128) edge_kernel(const EdgeArgs a, int cost_slot) {
  using Gr = Group<G>;
  constexpr int D = Gr::kDof;
  __shared__ double sred[4];
  double cost = 0.0;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < a.n) {
    const int p1 = a.i1[e];
    const int p2 = kBinary ? a.i2[e] : -1;
    const int o1 = a.pose_off[p1];
    const int o2 = kBinary ? a.pose_off[p2] : -1;
    if (kCostOnly || o1 >= 0 || o2 >= 0) {
      const typename Gr::T T1 = Gr::load(a.poses + (size_t)Gr::kStore * p1);
      const typename Gr::T Toi = Gr::inv(Gr::load(a.Tobs + (size_t)Gr::kStore * e));
      typename Gr::T E, T21;
      if constexpr (kBinary) {
        const typename Gr::T T2 = Gr::load(a.poses + (size_t)Gr::kStore * p2);
        const typename Gr::T T1i = Gr::inv(T1);
        E = Gr::mul(T2, Gr::mul(T1i, Toi));
        T21 = Gr::mul(T2, T1i);
      } else {
        E = Gr::mul(T1, Toi);
      }
      double xi[D], r[D], w[D], wr[D];
      Gr::log(E, xi);
      const double* Sm = a.stiff + (a.stiff_per_block ? (size_t)D * D * e : 0);
#pragma unroll
      for (int i = 0; i < D; ++i) {
        double v = 0.0;
#pragma unroll
        for (int k = 0; k < D; ++k) v += Sm[i * D + k] * xi[k];
        r[i] = v;
        cost += loss_rho(a.loss, v);
      }
      if (!kCostOnly) {
#pragma unroll
        for (int i = 0; i < D; ++i) {
          w[i] = loss_weight(a.loss, r[i]);
          wr[i] = w[i] * r[i];
        }
        double Js[D * D];   // J of the "outer" pose = S
#pragma unroll
        for (int i = 0; i < D * D; ++i) Js[i] = Sm[i];
        if constexpr (kBinary) {
          double Ad[D * D], J1[D * D];
          Gr::adjoint(T21, Ad);
#pragma unroll
          for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j < D; ++j) {
              double v = 0.0;
#pragma unroll
              for (int k = 0; k < D; ++k) v += Sm[i * D + k] * Ad[k * D + j];
              J1[i * D + j] = -v;
            }
          if (o1 >= 0) {
            add_block<D>(a.S, a.ldS, o1, o1, J1, J1, w, true);
#pragma unroll
            for (int c = 0; c < D; ++c) {
              double v = 0.0;
#pragma unroll
              for (int k = 0; k < D; ++k) v += J1[k * D + c] * wr[k];
              red_add(a.rhs + o1 + c, -v);
            }
          }
          if (o2 >= 0) {
            add_block<D>(a.S, a.ldS, o2, o2, Js, Js, w, true);
#pragma unroll
            for (int c = 0; c < D; ++c) {
              double v = 0.0;
#pragma unroll
              for (int k = 0; k < D; ++k) v += Js[k * D + c] * wr[k];
              red_add(a.rhs + o2 + c, -v);
            }
          }
          if (o1 >= 0 && o2 >= 0) {
            if (o2 > o1) add_block<D>(a.S, a.ldS, o2, o1, Js, J1, w, false);
            else if (o1 > o2) add_block<D>(a.S, a.ldS, o1, o2, J1, Js, w, false);
            else {  // same pose on both ends: symmetric cross term on the diagonal block
              add_block<D>(a.S, a.ldS, o1, o1, Js, J1, w, true);
              add_block<D>(a.S, a.ldS, o1, o1, J1, Js, w, true);
            }
          }
        } else if (o1 >= 0) {
          add_block<D>(a.S, a.ldS, o1, o1, Js, Js, w, true);
#pragma unroll
          for (int c = 0; c < D; ++c) {
            double v = 0.0;
#pragma unroll
            for (int k = 0; k < D; ++k) v += Js[k * D + c] * wr[k];
            red_add(a.rhs + o1 + c, -v);
          }
        }
      }
    }
  }
  block_sum_to(cost, a.scalars + cost_slot, sred);
}

// PoseToPoseOrientationResidual.evaluate (pyslam/residuals/pose_to_pose_orientation_residual.py:12-38): binary factor on two
// SE(3) poses from a relative ROTATION measurement C_2_1_obs (SO3):
//     r = S log_SO3( rot(T2 T1^-1) C_obs^-1 ),   J1 = -S [0 | rot(T2 T1^-1)],   J2 = S [0 | I]        (3 x 6 each)
// Tobs holds the 9 entries of C_obs per factor; stiffness 3x3.
template <bool kCostOnly>
__global__ void __launch_bounds__(128) orientation_edge_kernel(const EdgeArgs a, int cost_slot) {
  __shared__ double sred[4];
  double cost = 0.0;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < a.n) {
    const int p1 = a.i1[e], p2 = a.i2[e];
    const int o1 = a.pose_off[p1], o2 = a.pose_off[p2];
    if (kCostOnly || o1 >= 0 || o2 >= 0) {
      const double* R1 = a.poses + 12 * (size_t)p1;
      const double* R2 = a.poses + 12 * (size_t)p2;
      const double* Co = a.Tobs + 9 * (size_t)e;
      double C21[9];
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) C21[3 * i + j] = R2[3 * i] * R1[3 * j] + R2[3 * i + 1] * R1[3 * j + 1] + R2[3 * i + 2] * R1[3 * j + 2];
      SE3 E;
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) E.R[3 * i + j] = C21[3 * i] * Co[3 * j] + C21[3 * i + 1] * Co[3 * j + 1] + C21[3 * i + 2] * Co[3 * j + 2];
      E.t[0] = E.t[1] = E.t[2] = 0.0;
      double xi[6], r[3], w[3], wr[3];
      se3_log(E, xi);                           // zero translation: xi[3..5] = SO3.log(E.R)
      const double* Sm = a.stiff + (a.stiff_per_block ? (size_t)9 * e : 0);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        r[i] = Sm[3 * i] * xi[3] + Sm[3 * i + 1] * xi[4] + Sm[3 * i + 2] * xi[5];
        cost += loss_rho(a.loss, r[i]);
      }
      if (!kCostOnly) {
        double J1[9], J2[9];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          w[i] = loss_weight(a.loss, r[i]);
          wr[i] = w[i] * r[i];
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            J2[3 * i + j] = Sm[3 * i + j];
            J1[3 * i + j] = -(Sm[3 * i] * C21[j] + Sm[3 * i + 1] * C21[3 + j] + Sm[3 * i + 2] * C21[6 + j]);
          }
        }
        auto rhs3 = [&](int o, const double* J) {
#pragma unroll
          for (int c = 0; c < 3; ++c) red_add(a.rhs + o + 3 + c, -(J[c] * wr[0] + J[3 + c] * wr[1] + J[6 + c] * wr[2]));
        };
        if (o1 >= 0) { add_block<3>(a.S, a.ldS, o1 + 3, o1 + 3, J1, J1, w, true); rhs3(o1, J1); }
        if (o2 >= 0) { add_block<3>(a.S, a.ldS, o2 + 3, o2 + 3, J2, J2, w, true); rhs3(o2, J2); }
        if (o1 >= 0 && o2 >= 0) {
          if (o2 > o1) add_block<3>(a.S, a.ldS, o2 + 3, o1 + 3, J2, J1, w, false);
          else if (o1 > o2) add_block<3>(a.S, a.ldS, o1 + 3, o2 + 3, J1, J2, w, false);
          else { add_block<3>(a.S, a.ldS, o1 + 3, o1 + 3, J2, J1, w, true); add_block<3>(a.S, a.ldS, o1 + 3, o1 + 3, J1, J2, w, true); }
        }
      }
    }
  }
  block_sum_to(cost, a.scalars + cost_slot, sred);
}

}  // namespace bs

// lie.cuh -- SE(2)/SE(3) device arithmetic replacing the `liegroups` package on
// the hot path (call sites: pyslam/residuals/pose_residual.py:15-23,
// pose_to_pose_residual.py:16-28, reprojection_residual.py:16-31,
// pyslam/problem.py:400-409).  Formulas and branch points restate upstream
// liegroups (numpy backend) as summarised in SURVEY.md Appendix A:
// tangent order [rho; phi], small-angle branches at |angle| <= 1e-8,
// SO3.log through acos(0.5 tr R - 0.5), left perturbation.
#pragma once
#include "common.cuh"

namespace bs {

// ------------------------------------------------------------------ SE(3)
struct SE3 {
  double R[9];  // row-major
  double t[3];
};

BS_D SE3 se3_load(const double* p) {
  SE3 T;
#pragma unroll
  for (int i = 0; i < 9; ++i) T.R[i] = p[i];
#pragma unroll
  for (int i = 0; i < 3; ++i) T.t[i] = p[9 + i];
  return T;
}
BS_D void se3_store(double* p, const SE3& T) {
#pragma unroll
  for (int i = 0; i < 9; ++i) p[i] = T.R[i];
#pragma unroll
  for (int i = 0; i < 3; ++i) p[9 + i] = T.t[i];
}

BS_D SE3 se3_mul(const SE3& A, const SE3& B) {
  SE3 C;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j)
      C.R[3 * i + j] = A.R[3 * i] * B.R[j] + A.R[3 * i + 1] * B.R[3 + j] + A.R[3 * i + 2] * B.R[6 + j];
    C.t[i] = A.R[3 * i] * B.t[0] + A.R[3 * i + 1] * B.t[1] + A.R[3 * i + 2] * B.t[2] + A.t[i];
  }
  return C;
}

BS_D SE3 se3_inv(const SE3& A) {
  SE3 C;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) C.R[3 * i + j] = A.R[3 * j + i];
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) C.t[i] = -(C.R[3 * i] * A.t[0] + C.R[3 * i + 1] * A.t[1] + C.R[3 * i + 2] * A.t[2]);
  return C;
}

// exp([rho; phi]) = (SO3.exp(phi), J_l(phi) rho)
BS_D SE3 se3_exp(const double* xi) {
  SE3 T;
  const double rx = xi[0], ry = xi[1], rz = xi[2];
  const double px = xi[3], py = xi[4], pz = xi[5];
  const double th = sqrt(px * px + py * py + pz * pz);
  if (th <= kSmallAngle) {
    // R = I + phi^ ;  J_l = I + 0.5 phi^
    T.R[0] = 1.0; T.R[1] = -pz; T.R[2] = py;
    T.R[3] = pz;  T.R[4] = 1.0; T.R[5] = -px;
    T.R[6] = -py; T.R[7] = px;  T.R[8] = 1.0;
    T.t[0] = rx + 0.5 * (py * rz - pz * ry);
    T.t[1] = ry + 0.5 * (pz * rx - px * rz);
    T.t[2] = rz + 0.5 * (px * ry - py * rx);
    return T;
  }
  const double ax = px / th, ay = py / th, az = pz / th;
  double s, c;
  sincos(th, &s, &c);
  const double omc = 1.0 - c;
  T.R[0] = c + omc * ax * ax;      T.R[1] = omc * ax * ay - s * az; T.R[2] = omc * ax * az + s * ay;
  T.R[3] = omc * ay * ax + s * az; T.R[4] = c + omc * ay * ay;      T.R[5] = omc * ay * az - s * ax;
  T.R[6] = omc * az * ax - s * ay; T.R[7] = omc * az * ay + s * ax; T.R[8] = c + omc * az * az;
  const double A = s / th, B = 1.0 - A, C = omc / th;
  const double adr = ax * rx + ay * ry + az * rz;
  T.t[0] = A * rx + B * ax * adr + C * (ay * rz - az * ry);
  T.t[1] = A * ry + B * ay * adr + C * (az * rx - ax * rz);
  T.t[2] = A * rz + B * az * adr + C * (ax * ry - ay * rx);
  return T;
}

// log(T) = [J_l^-1(phi) t ; phi]
BS_D void se3_log(const SE3& T, double* xi) {
  double cosang = 0.5 * (T.R[0] + T.R[4] + T.R[8]) - 0.5;
  cosang = fmin(1.0, fmax(-1.0, cosang));
  const double ang = acos(cosang);
  double px, py, pz;
  if (ang <= kSmallAngle) {
    px = T.R[7]; py = T.R[2]; pz = T.R[3];        // vee(R - I)
  } else {
    const double f = 0.5 * ang / sin(ang);
    px = f * (T.R[7] - T.R[5]);
    py = f * (T.R[2] - T.R[6]);
    pz = f * (T.R[3] - T.R[1]);
  }
  const double th = sqrt(px * px + py * py + pz * pz);
  const double tx = T.t[0], ty = T.t[1], tz = T.t[2];
  if (th <= kSmallAngle) {                        // (I - 0.5 phi^) t
    xi[0] = tx - 0.5 * (py * tz - pz * ty);
    xi[1] = ty - 0.5 * (pz * tx - px * tz);
    xi[2] = tz - 0.5 * (px * ty - py * tx);
  } else {
    const double ax = px / th, ay = py / th, az = pz / th;
    const double h = 0.5 * th;
    const double hc = h / tan(h);
    const double adt = ax * tx + ay * ty + az * tz;
    xi[0] = hc * tx + (1.0 - hc) * ax * adt - h * (ay * tz - az * ty);
    xi[1] = hc * ty + (1.0 - hc) * ay * adt - h * (az * tx - ax * tz);
    xi[2] = hc * tz + (1.0 - hc) * az * adt - h * (ax * ty - ay * tx);
  }
  xi[3] = px; xi[4] = py; xi[5] = pz;
}

// Ad(T) = [[R, t^ R], [0, R]]  (6x6 row-major)
BS_D void se3_adjoint(const SE3& T, double* Ad) {
  const double tx = T.t[0], ty = T.t[1], tz = T.t[2];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const double r0 = T.R[j], r1 = T.R[3 + j], r2 = T.R[6 + j];
    Ad[0 * 6 + j] = r0; Ad[1 * 6 + j] = r1; Ad[2 * 6 + j] = r2;
    Ad[3 * 6 + 3 + j] = r0; Ad[4 * 6 + 3 + j] = r1; Ad[5 * 6 + 3 + j] = r2;
    Ad[3 * 6 + j] = 0.0; Ad[4 * 6 + j] = 0.0; Ad[5 * 6 + j] = 0.0;
    Ad[0 * 6 + 3 + j] = -tz * r1 + ty * r2;
    Ad[1 * 6 + 3 + j] = tz * r0 - tx * r2;
    Ad[2 * 6 + 3 + j] = -ty * r0 + tx * r1;
  }
}

// ------------------------------------------------------------------ SE(2)
struct SE2 {
  double R[4];  // row-major
  double t[2];
};

BS_D SE2 se2_load(const double* p) {
  SE2 T;
  T.R[0] = p[0]; T.R[1] = p[1]; T.R[2] = p[2]; T.R[3] = p[3]; T.t[0] = p[4]; T.t[1] = p[5];
  return T;
}
BS_D void se2_store(double* p, const SE2& T) {
  p[0] = T.R[0]; p[1] = T.R[1]; p[2] = T.R[2]; p[3] = T.R[3]; p[4] = T.t[0]; p[5] = T.t[1];
}
BS_D SE2 se2_mul(const SE2& A, const SE2& B) {
  SE2 C;
  C.R[0] = A.R[0] * B.R[0] + A.R[1] * B.R[2];
  C.R[1] = A.R[0] * B.R[1] + A.R[1] * B.R[3];
  C.R[2] = A.R[2] * B.R[0] + A.R[3] * B.R[2];
  C.R[3] = A.R[2] * B.R[1] + A.R[3] * B.R[3];
  C.t[0] = A.R[0] * B.t[0] + A.R[1] * B.t[1] + A.t[0];
  C.t[1] = A.R[2] * B.t[0] + A.R[3] * B.t[1] + A.t[1];
  return C;
}
BS_D SE2 se2_inv(const SE2& A) {
  SE2 C;
  C.R[0] = A.R[0]; C.R[1] = A.R[2]; C.R[2] = A.R[1]; C.R[3] = A.R[3];
  C.t[0] = -(C.R[0] * A.t[0] + C.R[1] * A.t[1]);
  C.t[1] = -(C.R[2] * A.t[0] + C.R[3] * A.t[1]);
  return C;
}
BS_D SE2 se2_exp(const double* xi) {
  SE2 T;
  const double phi = xi[2];
  double s, c;
  sincos(phi, &s, &c);
  T.R[0] = c; T.R[1] = -s; T.R[2] = s; T.R[3] = c;
  double A, B;
  if (fabs(phi) <= kSmallAngle) { A = 1.0; B = 0.5 * phi; }
  else { A = s / phi; B = (1.0 - c) / phi; }
  T.t[0] = A * xi[0] - B * xi[1];
  T.t[1] = B * xi[0] + A * xi[1];
  return T;
}
BS_D void se2_log(const SE2& T, double* xi) {
  const double phi = atan2(T.R[2], T.R[0]);
  double A, B;  // J_l^-1 = A I - B [[0,-1],[1,0]]
  if (fabs(phi) <= kSmallAngle) { A = 1.0; B = 0.5 * phi; }
  else { const double h = 0.5 * phi; A = h / tan(h); B = h; }
  xi[0] = A * T.t[0] + B * T.t[1];
  xi[1] = -B * T.t[0] + A * T.t[1];
  xi[2] = phi;
}
// Ad(T) = [[R, (t_y, -t_x)^T], [0 0 1]]  (3x3 row-major)
BS_D void se2_adjoint(const SE2& T, double* Ad) {
  Ad[0] = T.R[0]; Ad[1] = T.R[1]; Ad[2] = T.t[1];
  Ad[3] = T.R[2]; Ad[4] = T.R[3]; Ad[5] = -T.t[0];
  Ad[6] = 0.0;    Ad[7] = 0.0;    Ad[8] = 1.0;
}

// Group traits so the pose-graph kernels are written once.
template <int G> struct Group;
template <> struct Group<3> {
  using T = SE3;
  static constexpr int kDof = 6, kStore = 12;
  static BS_D T load(const double* p) { return se3_load(p); }
  static BS_D void store(double* p, const T& x) { se3_store(p, x); }
  static BS_D T mul(const T& a, const T& b) { return se3_mul(a, b); }
  static BS_D T inv(const T& a) { return se3_inv(a); }
  static BS_D T exp(const double* xi) { return se3_exp(xi); }
  static BS_D void log(const T& a, double* xi) { se3_log(a, xi); }
  static BS_D void adjoint(const T& a, double* Ad) { se3_adjoint(a, Ad); }
};
template <> struct Group<2> {
  using T = SE2;
  static constexpr int kDof = 3, kStore = 6;
  static BS_D T load(const double* p) { return se2_load(p); }
  static BS_D void store(double* p, const T& x) { se2_store(p, x); }
  static BS_D T mul(const T& a, const T& b) { return se2_mul(a, b); }
  static BS_D T inv(const T& a) { return se2_inv(a); }
  static BS_D T exp(const double* xi) { return se2_exp(xi); }
  static BS_D void log(const T& a, double* xi) { se2_log(a, xi); }
  static BS_D void adjoint(const T& a, double* Ad) { se2_adjoint(a, Ad); }
};

}  // namespace bs

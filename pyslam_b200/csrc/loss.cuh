// loss.cuh -- element-wise robust losses fused into the linearisation kernels.
// Formulas follow the reference as written: pyslam/losses.py:8-214
// (L2 :8-17, L1 :20-33, Cauchy :51-72, Huber :90-123, Tukey :141-175,
//  t-distribution :193-214).  IRLS is applied per scalar residual component
// (pyslam/problem.py:349-360).
#pragma once
#include "common.cuh"

namespace bs {

enum LossKind : int { kL2 = 0, kL1 = 1, kCauchy = 2, kHuber = 3, kTukey = 4, kTDist = 5 };

struct Loss {
  int kind;
  double k;
};

// rho(x)
BS_D double loss_rho(const Loss& L, double x) {
  switch (L.kind) {
    case kL2: return 0.5 * x * x;
    case kL1: return fabs(x);
    case kCauchy: { const double q = x / L.k; return (0.5 * L.k * L.k) * log(1.0 + q * q); }
    case kHuber: { const double a = fabs(x); return a <= L.k ? 0.5 * x * x : L.k * (a - 0.5 * L.k); }
    case kTukey: {
      const double c = L.k * L.k / 6.0;
      if (fabs(x) <= L.k) { const double q = x / L.k; const double u = 1.0 - q * q; return c * (1.0 - u * u * u); }
      return c;
    }
    default: return 0.5 * (L.k + 1.0) * log(1.0 + x * x / L.k);
  }
}

// IRLS weight w(x)
BS_D double loss_weight(const Loss& L, double x) {
  switch (L.kind) {
    case kL2: return 1.0;
    case kL1: { const double a = fabs(x); return a <= kSmallAngle ? nan("") : 1.0 / a; }
    case kCauchy: { const double q = x / L.k; return fast_rcp(1.0 + q * q); }
    case kHuber: { const double a = fabs(x); return a <= L.k ? 1.0 : L.k * fast_rcp(a); }
    case kTukey: { if (fabs(x) <= L.k) { const double q = x / L.k; return 1.0 - q * q; } return 0.0; }
    default: return (L.k + 1.0) * fast_rcp(L.k + x * x);
  }
}

// Compile-time loss kind (kKind >= 0) or the run-time switch above (kKind < 0): the block kernels are
// instantiated per kind so that the common single-group case carries no switch and no dead branches.
template <int kKind>
BS_D double loss_rho_t(const Loss& L, double x) {
  if constexpr (kKind < 0) return loss_rho(L, x);
  else { Loss c; c.kind = kKind; c.k = L.k; return loss_rho(c, x); }
}
template <int kKind>
BS_D double loss_weight_t(const Loss& L, double x) {
  if constexpr (kKind < 0) return loss_weight(L, x);
  else { Loss c; c.kind = kKind; c.k = L.k; return loss_weight(c, x); }
}

}  // namespace bs

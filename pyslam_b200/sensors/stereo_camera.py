"""Pinhole stereo camera, (u, v, disparity) measurements, origin in the left
camera -- same constructor and methods as reference
pyslam/sensors/stereo_camera.py:7-174.

These numpy methods serve problem set-up (simulating observations,
triangulating initial landmarks) and direct calls from user code.  Inside
`Problem.solve()` projection and its Jacobian are fused into the reprojection
linearisation kernel (csrc/reproj.cuh); only `intrinsics()` crosses the C ABI.
"""
import numpy as np


class StereoCamera:
    def __init__(self, cu, cv, fu, fv, b, w, h):
        self.cu, self.cv = float(cu), float(cv)
        self.fu, self.fv = float(fu), float(fv)
        self.b = float(b)
        self.w, self.h = int(w), int(h)

    def intrinsics(self):
        return (self.cu, self.cv, self.fu, self.fv, self.b)

    def clone(self):
        return type(self)(self.cu, self.cv, self.fu, self.fv, self.b, self.w, self.h)

    def compute_pixel_grid(self):
        self.u_grid, self.v_grid = np.meshgrid(np.arange(self.w, dtype=float),
                                               np.arange(self.h, dtype=float), indexing='xy')

    @staticmethod
    def _rows(a, what):
        a = np.atleast_2d(np.asarray(a, dtype=float))
        if a.shape[1] != 3:
            raise ValueError('{} must have shape (3,) or (N,3)'.format(what))
        return a

    def is_valid_measurement(self, uvd):
        """0 < d < w, 0 < v < h, 0 < u < w (disparity is compared with the image
        width, as the reference does: stereo_camera.py:93-97)."""
        m = self._rows(uvd, 'uvd')
        u, v, d = m[:, 0], m[:, 1], m[:, 2]
        ok = (d > 0.) & (d < self.w) & (v > 0.) & (v < self.h) & (u > 0.) & (u < self.w)
        return ok if ok.size > 1 else bool(ok[0])

    def project(self, pt_c, compute_jacobians=None):
        p = self._rows(pt_c, 'pt_c')
        x, y, inv_z = p[:, 0], p[:, 1], 1. / p[:, 2]
        uvd = np.empty_like(p)
        uvd[:, 0] = self.fu * x * inv_z + self.cu
        uvd[:, 1] = self.fv * y * inv_z + self.cv
        uvd[:, 2] = self.fu * self.b * inv_z
        if not compute_jacobians:
            return np.squeeze(uvd)
        inv_z2 = inv_z * inv_z
        jac = np.zeros((p.shape[0], 3, 3))
        jac[:, 0, 0] = self.fu * inv_z
        jac[:, 1, 1] = self.fv * inv_z
        jac[:, 0, 2] = -self.fu * x * inv_z2
        jac[:, 1, 2] = -self.fv * y * inv_z2
        jac[:, 2, 2] = -self.fu * self.b * inv_z2
        return np.squeeze(uvd), np.squeeze(jac)

    def triangulate(self, uvd, compute_jacobians=None):
        m = self._rows(uvd, 'uvd')
        u, v, d = m[:, 0], m[:, 1], m[:, 2]
        b_d = self.b / d
        aspect = self.fu / self.fv
        pt = np.empty_like(m)
        pt[:, 0] = (u - self.cu) * b_d
        pt[:, 1] = (v - self.cv) * b_d * aspect
        pt[:, 2] = self.fu * b_d
        if not compute_jacobians:
            return np.squeeze(pt)
        b_d2 = b_d / d
        jac = np.zeros((m.shape[0], 3, 3))
        jac[:, 0, 0] = b_d
        jac[:, 1, 1] = b_d * aspect
        jac[:, 0, 2] = (self.cu - u) * b_d2
        jac[:, 1, 2] = (self.cv - v) * b_d2 * aspect
        jac[:, 2, 2] = -self.fu * b_d2
        return np.squeeze(pt), np.squeeze(jac)

    def __repr__(self):
        return ('{}:\n cu: {:f}\n cv: {:f}\n fu: {:f}\n fv: {:f}\n  b: {:f}\n  w: {:d}\n  h: {:d}\n'
                .format(type(self).__name__, self.cu, self.cv, self.fu, self.fv, self.b, self.w, self.h))

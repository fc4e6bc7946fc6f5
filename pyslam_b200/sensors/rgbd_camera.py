"""Pinhole RGB-D camera, (u, v, depth z) measurements -- same constructor and methods as reference
pyslam/sensors/rgbd_camera.py:7-168.

Host-side set-up code like `StereoCamera`; inside `Problem.solve()` the model is selected in the CUDA kernels
by `intrinsics()[4] == 0` (no baseline: the third measurement is the depth itself, csrc/reproj.cuh,
csrc/photometric.cuh, csrc/motion_only.cuh, csrc/ransac.cuh).
"""
import numpy as np


class RGBDCamera:
    def __init__(self, cu, cv, fu, fv, w, h):
        self.cu, self.cv = float(cu), float(cv)
        self.fu, self.fv = float(fu), float(fv)
        self.w, self.h = int(w), int(h)

    def intrinsics(self):
        """(cu, cv, fu, fv, 0): a zero baseline tells the kernels this is the RGB-D model."""
        return (self.cu, self.cv, self.fu, self.fv, 0.)

    def clone(self):
        return type(self)(self.cu, self.cv, self.fu, self.fv, self.w, self.h)

    def compute_pixel_grid(self):
        self.u_grid, self.v_grid = np.meshgrid(np.arange(self.w, dtype=float),
                                               np.arange(self.h, dtype=float), indexing='xy')

    @staticmethod
    def _rows(a, what):
        a = np.atleast_2d(np.asarray(a, dtype=float))
        if a.shape[1] != 3:
            raise ValueError('{} must have shape (3,) or (N,3)'.format(what))
        return a

    def is_valid_measurement(self, uvz):
        """z > 0, 0 < v < h, 0 < u < w (rgbd_camera.py:103-110)."""
        m = self._rows(uvz, 'uvz')
        u, v, z = m[:, 0], m[:, 1], m[:, 2]
        ok = (z > 0.) & (v > 0.) & (v < self.h) & (u > 0.) & (u < self.w)
        return ok if ok.size > 1 else bool(ok[0])

    def project(self, pt_c, compute_jacobians=None):
        p = self._rows(pt_c, 'pt_c')
        x, y, inv_z = p[:, 0], p[:, 1], 1. / p[:, 2]
        uvz = np.empty_like(p)
        uvz[:, 0] = self.fu * x * inv_z + self.cu
        uvz[:, 1] = self.fv * y * inv_z + self.cv
        uvz[:, 2] = p[:, 2]
        if not compute_jacobians:
            return np.squeeze(uvz)
        inv_z2 = inv_z * inv_z
        jac = np.zeros((p.shape[0], 3, 3))
        jac[:, 0, 0] = self.fu * inv_z
        jac[:, 1, 1] = self.fv * inv_z
        jac[:, 0, 2] = -self.fu * x * inv_z2
        jac[:, 1, 2] = -self.fv * y * inv_z2
        jac[:, 2, 2] = 1.
        return np.squeeze(uvz), np.squeeze(jac)

    def triangulate(self, uvz, compute_jacobians=None):
        m = self._rows(uvz, 'uvz')
        u, v, z = m[:, 0], m[:, 1], m[:, 2]
        pt = np.empty_like(m)
        pt[:, 0] = (u - self.cu) * z / self.fu
        pt[:, 1] = (v - self.cv) * z / self.fv
        pt[:, 2] = z
        if not compute_jacobians:
            return np.squeeze(pt)
        jac = np.zeros((m.shape[0], 3, 3))
        jac[:, 0, 0] = z / self.fu
        jac[:, 1, 1] = z / self.fv
        jac[:, 0, 2] = (u - self.cu) / self.fu
        jac[:, 1, 2] = (v - self.cv) / self.fv
        jac[:, 2, 2] = 1.
        return np.squeeze(pt), np.squeeze(jac)

    def __repr__(self):
        # the reference's format string has a `b` placeholder without an argument (rgbd_camera.py:80-85) and raises
        return ('{}:\n cu: {:f}\n cv: {:f}\n fu: {:f}\n fv: {:f}\n  w: {:d}\n  h: {:d}\n'
                .format(type(self).__name__, self.cu, self.cv, self.fu, self.fv, self.w, self.h))

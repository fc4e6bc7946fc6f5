"""Dense direct (photometric) alignment residual -- same constructor and plug-in
protocol as the reference's PhotometricResidualSE3
(pyslam/residuals/photometric_residual.py:38-161).

With a single SE3 parameter and a built-in loss, `Problem.solve()` lowers the
block to the CUDA kernel in csrc/photometric.cuh (one thread per reference pixel,
48 B/pixel); the numpy `evaluate` below exists for direct calls and for the
two-parameter (SO3, t) form, which goes through the generic plug-in path.
"""
import numpy as np

from ..lie import SE3
from ..utils import bilinear_interpolate

BLOCK_PHOTOMETRIC = 4


class PhotometricResidualSE3:
    BLOCK_KIND = BLOCK_PHOTOMETRIC

    def __init__(self, camera, im_ref, depth_ref, im_track, im_jac, intensity_stiffness, depth_stiffness, min_grad=0.):
        """`depth_ref` is the disparity image for a StereoCamera (NaN / out-of-range
        entries are dropped, as are pixels whose gradient norm is below min_grad)."""
        self.camera = camera
        if not hasattr(camera, 'u_grid'):
            camera.compute_pixel_grid()
        uvd = np.stack([camera.u_grid.ravel(), camera.v_grid.ravel(), np.asarray(depth_ref, dtype=float).ravel()], axis=1)
        jac = np.stack([np.asarray(im_jac[0], dtype=float).ravel(), np.asarray(im_jac[1], dtype=float).ravel()], axis=1)
        ref = np.asarray(im_ref, dtype=float).ravel()
        with np.errstate(invalid='ignore'):
            keep = np.asarray(camera.is_valid_measurement(uvd)) & (np.linalg.norm(jac, axis=1) >= min_grad)
        self.uvd_ref = np.ascontiguousarray(uvd[keep])
        self.im_ref = np.ascontiguousarray(ref[keep])
        self.im_jac = np.ascontiguousarray(jac[keep])
        self.im_track = np.ascontiguousarray(im_track, dtype=float)
        self.intensity_stiffness = intensity_stiffness
        self.depth_stiffness = depth_stiffness
        self.intensity_covar = intensity_stiffness ** -2
        self.depth_covar = depth_stiffness ** -2
        self.min_grad = min_grad
        self.pt_ref, self.triang_jac = camera.triangulate(self.uvd_ref, compute_jacobians=True)
        self.pt_ref = np.atleast_2d(self.pt_ref)
        self.triang_jac = self.triang_jac.reshape(-1, 3, 3)

    def evaluate(self, params, compute_jacobians=None):
        if len(params) == 1:
            T = params[0]
        elif len(params) == 2:
            T = SE3(params[0], params[1])
        else:
            raise ValueError('In PhotometricResidual.evaluate() params must have length 1 or 2')
        R = T.rot.as_matrix()
        pt = self.pt_ref @ R.T + T.trans
        uvd, pj = self.camera.project(pt, compute_jacobians=True)
        uvd, pj = np.atleast_2d(uvd), pj.reshape(-1, 3, 3)
        valid = np.atleast_1d(self.camera.is_valid_measurement(uvd))
        est = np.atleast_1d(bilinear_interpolate(self.im_track, uvd[:, 0], uvd[:, 1]))
        p = np.einsum('ni,nij->nj', self.im_jac, pj[:, 0:2, :])                 # image gradient through the projection
        jd = np.einsum('nj,nj->n', p @ R, self.triang_jac[:, :, 2])              # d residual / d disparity
        stiff = 1. / np.sqrt(self.intensity_covar + self.depth_covar * jd ** 2)
        residual = (stiff * (est - self.im_ref))[valid]
        if not compute_jacobians:
            return residual
        jac = None
        if any(compute_jacobians):
            jac = (stiff[:, None] * np.einsum('nj,njk->nk', p, np.atleast_3d(SE3.odot(pt)).reshape(-1, 3, 6)))[valid]
        if len(params) == 1:
            return residual, [jac if compute_jacobians[0] else None]
        return residual, [jac[:, 3:6] if compute_jacobians[0] else None, jac[:, 0:3] if compute_jacobians[1] else None]

"""Built-in residual blocks.

Every class keeps the reference's constructor and the duck-typed plug-in
protocol `evaluate(params, compute_jacobians=None) -> residual |
(residual, jacobians)` (contract: reference examples/Fitting a cubic.ipynb
cells 3-4), so user code and tests can call a block directly.  `Problem.solve()`
does NOT call these `evaluate` methods for the built-in types: it recognises the
class (`BLOCK_KIND`), packs the measurement into SoA batches
(pyslam_b200/problem.py: Problem._lower) and linearises them on the GPU (csrc/panel.cuh, csrc/reproj.cuh,
csrc/posegraph.cuh).  Only residual types the library does not know are
evaluated through this Python protocol and uploaded as dense blocks.

Jacobian conventions reproduced from the reference: left
perturbation, tangent order [rho; phi], and the *approximate* pose Jacobians
J = S (PoseResidual) and J1 = -S Ad(T2 T1^-1), J2 = S (PoseToPoseResidual).
"""
import numpy as np

from ..lie import SE3, group_of

BLOCK_REPROJECTION, BLOCK_POSE, BLOCK_POSE_TO_POSE = 1, 2, 3


def _wants(compute_jacobians, i):
    return bool(compute_jacobians[i])


class ReprojectionResidual:
    """Stiffness-weighted reprojection error of one landmark in one camera;
    params = [T_cam_w (SE3), pt_w (3,)].  Reference:
    pyslam/residuals/reprojection_residual.py:5-37."""
    BLOCK_KIND = BLOCK_REPROJECTION

    def __init__(self, camera, obs, stiffness):
        self.camera = camera
        self.obs = obs
        self.stiffness = stiffness

    def evaluate(self, params, compute_jacobians=None):
        T_cam_w, pt_w = params
        S = np.asarray(self.stiffness, dtype=float)
        pt_cam = T_cam_w.dot(pt_w)
        if not compute_jacobians:
            return S @ (self.camera.project(pt_cam) - np.asarray(self.obs, dtype=float))
        uvd, J_cam = self.camera.project(pt_cam, compute_jacobians=True)
        SJ = S @ J_cam
        out = [None, None]
        if _wants(compute_jacobians, 0):
            out[0] = SJ @ SE3.odot(pt_cam)
        if _wants(compute_jacobians, 1):
            out[1] = SJ @ T_cam_w.rot.as_matrix()
        return S @ (uvd - np.asarray(self.obs, dtype=float)), out


class PoseResidual:
    """Unary prior r = S log(T T_obs^-1) on an SE2/SE3 pose.  Reference:
    pyslam/residuals/pose_residual.py:4-27."""
    BLOCK_KIND = BLOCK_POSE

    def __init__(self, T_obs, stiffness):
        self.T_obs = T_obs
        self.stiffness = stiffness
        self.obstype = type(T_obs)

    def evaluate(self, params, compute_jacobians=None):
        S = np.asarray(self.stiffness, dtype=float)
        r = S @ params[0].dot(self.T_obs.inv()).log()
        if not compute_jacobians:
            return r
        return r, [S.copy() if _wants(compute_jacobians, 0) else None]


class PoseToPoseResidual:
    """Relative-pose factor r = S log(T2 T1^-1 T21_obs^-1); params = [T1, T2].
    Reference: pyslam/residuals/pose_to_pose_residual.py:4-32."""
    BLOCK_KIND = BLOCK_POSE_TO_POSE

    def __init__(self, T_2_1_obs, stiffness):
        self.T_2_1_obs = T_2_1_obs
        self.stiffness = stiffness
        self.obstype = type(T_2_1_obs)

    def evaluate(self, params, compute_jacobians=None):
        T1, T2 = params
        S = np.asarray(self.stiffness, dtype=float)
        T1_inv = T1.inv()
        r = S @ T2.dot(T1_inv.dot(self.T_2_1_obs.inv())).log()
        if not compute_jacobians:
            return r
        out = [None, None]
        if _wants(compute_jacobians, 0):
            out[0] = -(S @ T2.dot(T1_inv).adjoint())
        if _wants(compute_jacobians, 1):
            out[1] = S.copy()
        return r, out


class QuadraticResidual:
    """r = s (a x^2 + b x + c - y); params = [a, b, c].  Reference:
    pyslam/residuals/quadratic_residual.py:4-32.  (No GPU kernel: goes through
    the generic dense-block path like any user-defined residual.)"""

    def __init__(self, x, y, stiffness):
        self.x = np.array([x], dtype=float)
        self.y = np.array([y], dtype=float)
        self.stiffness = np.array([stiffness], dtype=float)

    def evaluate(self, params, compute_jacobians=None):
        basis = (self.x * self.x, self.x, np.ones(1))
        model = sum(np.asarray(p, dtype=float).reshape(-1)[0] * phi for p, phi in zip(params, basis))
        r = self.stiffness * (model - self.y)
        if not compute_jacobians:
            return r
        # unlike the reference (np.squeeze over a ragged list, which numpy>=1.24
        # rejects for mixed True/False), non-requested entries are simply None
        return r, [float((self.stiffness * phi)[0]) if want else None
                   for phi, want in zip(basis, compute_jacobians)]

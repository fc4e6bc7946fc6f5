"""Frame-to-frame (pose-only) reprojection residuals and the SO(3)-only relative factor -- same constructors and
plug-in protocol as the reference's

    ReprojectionMotionOnlyResidual        pyslam/residuals/reprojection_motion_only_residual.py:36-67
    ReprojectionMotionOnlyBatchResidual   pyslam/residuals/reprojection_motion_only_residual.py:70-113
    PoseToPoseOrientationResidual         pyslam/residuals/pose_to_pose_orientation_residual.py:4-38

With a built-in camera and loss `Problem.solve()` lowers them to CUDA kernels (csrc/motion_only.cuh: one thread per
point, 6x6 block reduction; csrc/posegraph.cuh: orientation_edge_kernel); the numpy `evaluate` methods below serve
direct calls and the generic plug-in path (user-defined cameras / losses).
"""
import numpy as np

from ..lie import SE3

BLOCK_MOTION_ONLY, BLOCK_MOTION_ONLY_BATCH, BLOCK_ORIENTATION = 5, 6, 7


def se3_odot_stack(pts):
    """[I | -p^] for every row of pts: (N, 3, 6) -- pyslam's fast_se3_odot."""
    pts = np.atleast_2d(pts)
    out = np.zeros((len(pts), 3, 6))
    out[:, 0, 0] = out[:, 1, 1] = out[:, 2, 2] = 1.
    out[:, 0, 4] = pts[:, 2]; out[:, 0, 5] = -pts[:, 1]
    out[:, 1, 3] = -pts[:, 2]; out[:, 1, 5] = pts[:, 0]
    out[:, 2, 3] = pts[:, 1]; out[:, 2, 4] = -pts[:, 0]
    return out


class ReprojectionMotionOnlyResidual:
    """One point: params = [T_2_1]."""
    BLOCK_KIND = BLOCK_MOTION_ONLY

    def __init__(self, camera, obs_1, obs_2, stiffness):
        self.camera, self.obs_1, self.obs_2, self.stiffness = camera, obs_1, obs_2, stiffness
        self.pt_1 = self.camera.triangulate(self.obs_1)

    def evaluate(self, params, compute_jacobians=None):
        S = np.asarray(self.stiffness, dtype=float)
        pt_2 = params[0].dot(self.pt_1)
        if not compute_jacobians:
            return S @ (self.camera.project(pt_2) - np.asarray(self.obs_2, dtype=float))
        pred, cam_jac = self.camera.project(pt_2, compute_jacobians=True)
        jac = S @ cam_jac @ SE3.odot(pt_2) if compute_jacobians[0] else None
        return S @ (pred - np.asarray(self.obs_2, dtype=float)), [jac]


class ReprojectionMotionOnlyBatchResidual:
    """N points at once: residual (3N,), Jacobian (3N, 6); params = [T_2_1]."""
    BLOCK_KIND = BLOCK_MOTION_ONLY_BATCH

    def __init__(self, camera, obs_1, obs_2, stiffness):
        self.camera, self.obs_1, self.obs_2, self.stiffness = camera, obs_1, obs_2, stiffness
        self.pts_1 = np.atleast_2d(self.camera.triangulate(self.obs_1))
        self.num_pts = self.pts_1.shape[0]

    def evaluate(self, params, compute_jacobians=None):
        S = np.asarray(self.stiffness, dtype=float)
        T = params[0]
        pts_2 = self.pts_1 @ T.rot.as_matrix().T + T.trans
        obs_2 = np.atleast_2d(np.asarray(self.obs_2, dtype=float))
        if not compute_jacobians:
            pred = np.atleast_2d(self.camera.project(pts_2))
            return ((pred - obs_2) @ S.T).reshape(3 * self.num_pts)
        pred, cam_jac = self.camera.project(pts_2, compute_jacobians=True)
        pred, cam_jac = np.atleast_2d(pred), cam_jac.reshape(-1, 3, 3)
        residual = ((pred - obs_2) @ S.T).reshape(3 * self.num_pts)
        jac = None
        if compute_jacobians[0]:
            jac = np.einsum('ij,njk,nkl->nil', S, cam_jac, se3_odot_stack(pts_2)).reshape(3 * self.num_pts, 6)
        return residual, [jac]


class PoseToPoseOrientationResidual:
    """r = S log_SO3(rot(T2 T1^-1) C_2_1_obs^-1); params = [T_1_0, T_2_0] (SE3)."""
    BLOCK_KIND = BLOCK_ORIENTATION

    def __init__(self, C_2_1_obs, stiffness):
        self.C_2_1_obs = C_2_1_obs
        self.stiffness = stiffness
        self.obstype = type(C_2_1_obs)

    def evaluate(self, params, compute_jacobians=None):
        T1, T2 = params
        S = np.asarray(self.stiffness, dtype=float)
        C21 = T2.dot(T1.inv()).rot
        r = S @ C21.dot(self.C_2_1_obs.inv()).log()
        if not compute_jacobians:
            return r
        out = [None, None]
        if compute_jacobians[0]:
            P1 = np.zeros((3, 6))
            P1[:, 3:] = C21.as_matrix()
            out[0] = S @ -P1
        if compute_jacobians[1]:
            P2 = np.zeros((3, 6))
            P2[:, 3:] = np.eye(3)
            out[1] = S @ P2
        return r, out

"""TEST INFRASTRUCTURE -- CPU restatement (numpy) of the reference code on the rows SURVEY 8 marks "next":
f2 (motion-only reprojection, RANSAC), f3 (pyramids), f4 (RGB-D camera, SO(3)-only factor).  Pinned by the
fixtures oracle/make_golden_r2.py produced with the unmodified reference (tests/test_oracle_golden_r2.py).
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
import numpy as np

from . import liegroups as OL


class RGBDCamera:  # rgbd_camera.py:7-168
    def __init__(self, cu, cv, fu, fv, w, h):
        self.cu, self.cv, self.fu, self.fv, self.w, self.h = float(cu), float(cv), float(fu), float(fv), int(w), int(h)

    def is_valid_measurement(self, uvz):  # rgbd_camera.py:103-110
        m = np.atleast_2d(uvz)
        return (m[:, 2] > 0.) & (m[:, 1] > 0.) & (m[:, 1] < self.h) & (m[:, 0] > 0.) & (m[:, 0] < self.w)

    def project(self, pt_c, compute_jacobians=None):  # rgbd_camera.py:113-146
        p = np.atleast_2d(pt_c)
        iz = 1. / p[:, 2]
        uvz = np.stack([self.fu * p[:, 0] * iz + self.cu, self.fv * p[:, 1] * iz + self.cv, p[:, 2]], axis=1)
        if not compute_jacobians:
            return np.squeeze(uvz)
        J = np.zeros((len(p), 3, 3))
        J[:, 0, 0], J[:, 0, 2] = self.fu * iz, -self.fu * p[:, 0] * iz * iz
        J[:, 1, 1], J[:, 1, 2] = self.fv * iz, -self.fv * p[:, 1] * iz * iz
        J[:, 2, 2] = 1.
        return np.squeeze(uvz), np.squeeze(J)

    def triangulate(self, uvz, compute_jacobians=None):  # rgbd_camera.py:149-180
        m = np.atleast_2d(uvz)
        pt = np.stack([(m[:, 0] - self.cu) * m[:, 2] / self.fu, (m[:, 1] - self.cv) * m[:, 2] / self.fv, m[:, 2]], axis=1)
        if not compute_jacobians:
            return np.squeeze(pt)
        J = np.zeros((len(m), 3, 3))
        J[:, 0, 0], J[:, 0, 2] = m[:, 2] / self.fu, (m[:, 0] - self.cu) / self.fu
        J[:, 1, 1], J[:, 1, 2] = m[:, 2] / self.fv, (m[:, 1] - self.cv) / self.fv
        J[:, 2, 2] = 1.
        return np.squeeze(pt), np.squeeze(J)


def odot_stack(p):  # reprojection_motion_only_residual.py:12-32
    out = np.zeros((len(p), 3, 6))
    out[:, [0, 1, 2], [0, 1, 2]] = 1.
    out[:, 0, 4], out[:, 0, 5] = p[:, 2], -p[:, 1]
    out[:, 1, 3], out[:, 1, 5] = -p[:, 2], p[:, 0]
    out[:, 2, 3], out[:, 2, 4] = p[:, 1], -p[:, 0]
    return out


class ReprojectionMotionOnlyBatchResidual:  # reprojection_motion_only_residual.py:70-113
    def __init__(self, camera, obs_1, obs_2, stiffness):
        self.camera, self.obs_2, self.stiffness = camera, np.atleast_2d(obs_2), np.asarray(stiffness, dtype=float)
        self.pts_1 = np.atleast_2d(camera.triangulate(obs_1))
        self.num_pts = len(self.pts_1)

    def evaluate(self, params, compute_jacobians=None):
        T = params[0]
        pts_2 = self.pts_1 @ T.rot.mat.T + T.trans
        if not compute_jacobians:
            return ((np.atleast_2d(self.camera.project(pts_2)) - self.obs_2) @ self.stiffness.T).reshape(-1)
        pred, cj = self.camera.project(pts_2, True)
        r = ((np.atleast_2d(pred) - self.obs_2) @ self.stiffness.T).reshape(-1)
        J = np.einsum('ij,njk,nkl->nil', self.stiffness, cj.reshape(-1, 3, 3), odot_stack(pts_2)).reshape(-1, 6)
        return r, [J if compute_jacobians[0] else None]


def compute_transforms(pts_1_sets, pts_2_sets):  # ransac.py:12-56 (SVD method), stacked [n_hyp, n_min, 3]
    c1, c2 = pts_1_sets.mean(axis=1, keepdims=True), pts_2_sets.mean(axis=1, keepdims=True)
    W = np.einsum('hni,hnj->hij', pts_2_sets - c2, pts_1_sets - c1) / pts_1_sets.shape[1]
    U, _, V = np.linalg.svd(W)
    Sg = np.tile(np.eye(3), (len(W), 1, 1))
    Sg[:, 2, 2] = np.linalg.det(U) * np.linalg.det(V)
    C = U @ Sg @ V
    T = np.tile(np.eye(4), (len(W), 1, 1))
    T[:, :3, :3] = C
    T[:, :3, 3] = c2[:, 0] - np.einsum('hij,hj->hi', C, c1[:, 0])
    return T


def ransac_masks(T_stack, pts_1, obs_2, camera, thresh):  # ransac.py:155-165
    out = []
    for T in T_stack:
        pred = np.atleast_2d(camera.project(pts_1 @ T[:3, :3].T + T[:3, 3]))
        out.append(((pred - obs_2) ** 2).sum(axis=1) < thresh)
    return np.array(out)


class PoseToPoseOrientationResidual:  # pose_to_pose_orientation_residual.py:4-38
    def __init__(self, C_2_1_obs, stiffness):
        self.C_2_1_obs, self.stiffness = C_2_1_obs, np.asarray(stiffness, dtype=float)

    def evaluate(self, params, compute_jacobians=None):
        T1, T2 = params
        C21 = T2.dot(T1.inv()).rot
        r = self.stiffness @ C21.dot(self.C_2_1_obs.inv()).log()
        if not compute_jacobians:
            return r
        P1, P2 = np.zeros((3, 6)), np.zeros((3, 6))
        P1[:, 3:], P2[:, 3:] = C21.mat, np.eye(3)
        return r, [self.stiffness @ -P1 if compute_jacobians[0] else None, self.stiffness @ P2 if compute_jacobians[1] else None]


# ---- pyramids: the published OpenCV algorithms the reference calls (cv2 4.x; keyframes.py:30-46,92-114) ----
def _reflect101(idx, n):
    idx = np.abs(idx)
    return np.where(idx >= n, 2 * n - 2 - idx, idx)


def pyr_down_u8(im):
    """cv2.pyrDown on uint8: separable [1 4 6 4 1], BORDER_REFLECT_101, (sum + 128) >> 8, size ((w+1)/2, (h+1)/2)."""
    h, w = im.shape
    ho, wo = (h + 1) // 2, (w + 1) // 2
    k = np.array([1, 4, 6, 4, 1], dtype=np.int64)
    a = im.astype(np.int64)
    rows = _reflect101(2 * np.arange(ho)[:, None] + np.arange(-2, 3)[None, :], h)          # [ho, 5]
    cols = _reflect101(2 * np.arange(wo)[:, None] + np.arange(-2, 3)[None, :], w)          # [wo, 5]
    tmp = np.einsum('hkw,k->hw', a[rows], k)                                                 # vertical pass, all columns
    out = np.einsum('hwk,k->hw', tmp[:, cols], k)
    return ((out + 128) >> 8).astype(np.uint8)


def sobel_half(im):
    """0.5 * cv2.Sobel(im, -1, 1, 0) and 0.5 * cv2.Sobel(im, -1, 0, 1), ksize 3, BORDER_REFLECT_101."""
    p = np.pad(im, 1, mode='reflect')
    gx = (p[:-2, 2:] - p[:-2, :-2]) + 2 * (p[1:-1, 2:] - p[1:-1, :-2]) + (p[2:, 2:] - p[2:, :-2])
    gy = (p[2:, :-2] - p[:-2, :-2]) + 2 * (p[2:, 1:-1] - p[:-2, 1:-1]) + (p[2:, 2:] - p[:-2, 2:])
    return 0.5 * gx, 0.5 * gy


def image_pyramid(im_u8, levels):  # keyframes.py:30-46
    pyr = [im_u8]
    for _ in range(1, levels):
        pyr.append(pyr_down_u8(pyr[-1]))
    ims = [p.astype(float) / 255. for p in pyr]
    return ims, [np.array(sobel_half(i)) for i in ims]


def subsample_pyramid(m, levels, scale_per_level):  # keyframes.py:59-72,100-112
    out, cur = [m], m
    for l in range(1, levels):
        cur = cur[0::2, 0::2]
        out.append(cur * scale_per_level ** l)
    return out

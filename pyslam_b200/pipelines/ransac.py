"""Frame-to-frame RANSAC -- same class and attributes as reference pyslam/pipelines/ransac.py:98-165.

All `ransac_iters` hypotheses are evaluated in ONE C-ABI call (`bslam_ransac`, csrc/ransac.cuh): rigid transforms of
the minimal sets by the SVD method (one thread each), inlier counts of |project(T p_1) - obs_2|^2 < thresh (one CTA
per hypothesis), first arg-max and the winner's mask.  The random minimal sets are drawn on the host with
`np.random.randint` exactly as the reference does (same generator state -> same sets).
"""
import numpy as np

from .. import engine as _engine
from ..lie import SE3


class FrameToFrameRANSAC:
    def __init__(self, camera, device=0):
        self.camera = camera
        self.ransac_iters = 400
        self.ransac_thresh = 5  # (1**2 + 1**2 + 1**2)
        self.num_min_set_pts = 3
        self.device = device

    def set_obs(self, obs_1, obs_2):
        self.obs_1 = np.atleast_2d(obs_1)
        self.obs_2 = np.atleast_2d(obs_2)
        self.pts_1 = np.atleast_2d(self.camera.triangulate(self.obs_1))
        self.pts_2 = np.atleast_2d(self.camera.triangulate(self.obs_2))
        self.num_pts = self.pts_1.shape[0]

    def perform_ransac(self):
        """(T_21_best, obs_1_inliers, obs_2_inliers, inlier_indices_best)"""
        rand_idx = np.random.randint(self.num_pts, size=(self.ransac_iters, self.num_min_set_pts))
        T_all, counts, best, mask = _engine.ransac(self.pts_1, self.obs_2, self.camera.intrinsics(), self.ransac_thresh,
                                                   sample_idx=rand_idx, pts_2=self.pts_2, device=self.device)
        self.T_21_stacked, self.inlier_counts = T_all, counts
        if counts[best] < 5:
            raise ValueError(' RANSAC failed to find more than 5 inliers. Try adjusting the thresholds.')
        inlier_indices_best = np.where(mask)[0]
        T_21_best = SE3.from_matrix(T_all[best])
        return T_21_best, self.obs_1[inlier_indices_best], self.obs_2[inlier_indices_best], inlier_indices_best

    def compute_transform(self, pts_1_stacked, pts_2_stacked):
        """Rigid transforms of stacked minimal sets [n_hyp, n_min, 3] (reference compute_transform_fast)."""
        p1 = np.asarray(pts_1_stacked, dtype=float)
        p2 = np.asarray(pts_2_stacked, dtype=float)
        n_hyp, n_min = p1.shape[:2]
        idx = np.arange(n_hyp * n_min, dtype=np.int32).reshape(n_hyp, n_min)
        flat1, flat2 = p1.reshape(-1, 3), p2.reshape(-1, 3)
        obs = np.atleast_2d(self.camera.project(flat2))
        return _engine.ransac(flat1, obs, self.camera.intrinsics(), self.ransac_thresh, sample_idx=idx, pts_2=flat2,
                              device=self.device)[0]

    def compute_ransac_cost(self, T_21_stacked, pts_1, obs_2, camera, inlier_thresh):
        """Boolean inlier masks [n_hyp, n_pts] of given hypotheses (reference signature).  The full masks are only
        needed by callers of this method; perform_ransac keeps counts + the winner's mask on the device."""
        T = np.asarray(T_21_stacked, dtype=float).reshape(-1, 4, 4)
        out = np.zeros((len(T), len(np.atleast_2d(pts_1))), dtype=bool)
        for k in range(len(T)):     # one mask per call: the device routine returns the mask of its best hypothesis
            out[k] = _engine.ransac(pts_1, obs_2, camera.intrinsics(), inlier_thresh, T_21=T[k:k + 1], device=self.device)[3]
        return out

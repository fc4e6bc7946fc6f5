"""Sparse (feature-based) visual odometry -- the classes and attributes of reference pyslam/pipelines/sparse.py.

The hot part of `_compute_frame_to_frame_motion` (sparse.py:130-163, 186-216) runs on the GPU: all 400 RANSAC hypotheses in
one call (csrc/ransac.cuh), then the pose-only Gauss-Newton refinement on the inliers (csrc/motion_only.cuh) on ONE engine
handle for the whole sequence.  Feature matching itself is outside the hot path: the reference hard-wires libviso2, which is
not installable here, so the matcher is a plug-in -- any object with

    match(ref_frame, track_frame) -> (obs_0, obs_1)        # (N, 3) arrays: (u, v, disparity) or (u, v, depth)

(`Viso2Matcher` below wraps libviso2 when it can be imported).  `estimate_motion(obs_0, obs_1)` is the matcher-free entry:
prune, RANSAC, refine.
"""
import numpy as np

from .. import engine as _engine
from ..lie import SE3
from ..losses import L2Loss
from ..problem import Options, Problem
from ..residuals import ReprojectionMotionOnlyBatchResidual
from .keyframes import SparseRGBDKeyframe, SparseStereoKeyframe
from .ransac import FrameToFrameRANSAC


class Viso2Matcher:
    """libviso2 quad / flow matching as the reference drives it (sparse.py:46-51, 131-141, 187-199)."""

    def __init__(self, camera, mode):
        import viso2                                   # raises ImportError where libviso2 is not installed
        self.mode = mode
        self.matcher = viso2.Matcher(viso2.Matcher_parameters())
        if mode == 2:
            self.matcher.setIntrinsics(camera.fu, camera.cu, camera.cv, camera.b)

    def match(self, ref_frame, track_frame):
        if self.mode == 2:
            self.matcher.pushBack(ref_frame.im_left, ref_frame.im_right)
            self.matcher.pushBack(track_frame.im_left, track_frame.im_right)
        else:
            self.matcher.pushBack(ref_frame.image)
            self.matcher.pushBack(track_frame.image)
        self.matcher.matchFeatures(self.mode)
        ms = self.matcher.getMatches()
        if self.mode == 2:
            return (np.array([[m.u1p, m.v1p, m.u1p - m.u2p] for m in ms]), np.array([[m.u1c, m.v1c, m.u1c - m.u2c] for m in ms]))
        return (np.array([[m.u1p, m.v1p, ref_frame.depth[int(m.v1p), int(m.u1p)]] for m in ms]),
                np.array([[m.u1c, m.v1c, track_frame.depth[int(m.v1c), int(m.u1c)]] for m in ms]))


class SparseVOPipeline:
    """Base class for sparse VO pipelines"""

    def __init__(self, camera, first_pose=None, matcher=None, device=0):
        self.camera = camera
        self.first_pose = SE3.identity() if first_pose is None else first_pose
        self.keyframes = []
        self.T_c_w = [self.first_pose]
        o = Options()                                   # sparse.py:30-39
        o.allow_nondecreasing_steps = True
        o.max_nondecreasing_steps = 5
        o.min_cost_decrease = 0.99
        o.max_iters = 30
        o.num_threads = 1
        o.linesearch_max_iters = 0
        o.device = device
        self.motion_options = o
        self.keyframe_trans_thresh = 3.0                # meters
        self.keyframe_rot_thresh = 0.3                  # rad
        self.matcher = matcher
        self.matcher_mode = 0
        self.ransac = FrameToFrameRANSAC(self.camera, device)
        self.reprojection_stiffness = np.diag([1., 1., 1.])
        self.mode = 'map'
        self.loss = L2Loss()
        self.device = device
        self._engine = None
        self.last_cost_history = None

    def set_mode(self, mode):
        self.mode = mode
        if self.mode == 'track':
            self.active_keyframe_idx = 0
            self.T_c_w = []

    def estimate_motion(self, obs_0, obs_1):
        """T_1_0 from matched observations: prune non-positive third coordinates, RANSAC, motion-only refinement on the
        inliers (the tail of the reference's `_compute_frame_to_frame_motion`)."""
        obs_0, obs_1 = np.atleast_2d(np.asarray(obs_0, dtype=float)), np.atleast_2d(np.asarray(obs_1, dtype=float))
        keep = (obs_0[:, 2] > 0) & (obs_1[:, 2] > 0)
        self.obs_0, self.obs_1 = obs_0[keep], obs_1[keep]
        self.ransac.set_obs(self.obs_0, self.obs_1)
        T_guess, in_0, in_1, _ = self.ransac.perform_ransac()
        residual = ReprojectionMotionOnlyBatchResidual(self.camera, in_0, in_1, self.reprojection_stiffness)
        if self._engine is None:
            self._engine = _engine.Engine(self.device)
        problem = Problem(self.motion_options, engine=self._engine)
        problem.add_residual_block(residual, ['T_1_0'], loss=self.loss)
        problem.initialize_params({'T_1_0': T_guess})
        params = problem.solve()
        self.last_cost_history = list(problem._cost_history)
        return params['T_1_0']

    def _compute_frame_to_frame_motion(self, ref_frame, track_frame):
        if self.matcher is None:
            raise RuntimeError('no feature matcher: pass matcher= (an object with match(ref_frame, track_frame) -> (obs_0, obs_1)); '
                               'the reference hard-wires libviso2, which is not installed here')
        return self.estimate_motion(*self.matcher.match(ref_frame, track_frame))

    def track(self, trackframe):
        """Track a frame against the active keyframe (sparse.py:71-110)."""
        if len(self.keyframes) == 0:
            self.keyframes.append(trackframe)
            self.active_keyframe_idx = 0
            return
        active = self.keyframes[self.active_keyframe_idx]
        T_track_ref = self._compute_frame_to_frame_motion(active, trackframe)
        T_track_ref.normalize()
        self.T_c_w.append(T_track_ref.dot(active.T_c_w))
        xi = T_track_ref.log()
        if np.linalg.norm(xi[0:3]) > self.keyframe_trans_thresh or np.linalg.norm(xi[3:6]) > self.keyframe_rot_thresh:
            if self.mode == 'map':
                trackframe.T_c_w = self.T_c_w[-1]
                self.keyframes.append(trackframe)
            self.active_keyframe_idx += 1


class SparseStereoPipeline(SparseVOPipeline):
    """Sparse stereo VO pipeline"""

    def __init__(self, camera, first_pose=None, matcher=None, device=0):
        super().__init__(camera, first_pose, matcher, device)
        self.matcher_mode = 2                           # stereo quad matching
        if self.matcher is None:
            try:
                self.matcher = Viso2Matcher(camera, 2)
            except ImportError:
                pass

    def track(self, im_left, im_right):
        T = self.T_c_w[0] if len(self.keyframes) == 0 else None
        super().track(SparseStereoKeyframe(im_left, im_right, T))


class SparseRGBDPipeline(SparseVOPipeline):
    """Sparse RGBD VO pipeline"""

    def __init__(self, camera, first_pose=None, matcher=None, device=0):
        super().__init__(camera, first_pose, matcher, device)
        self.matcher_mode = 0                           # mono-to-mono
        if self.matcher is None:
            try:
                self.matcher = Viso2Matcher(camera, 0)
            except ImportError:
                pass

    def track(self, image, depth):
        T = self.T_c_w[0] if len(self.keyframes) == 0 else None
        super().track(SparseRGBDKeyframe(image, depth, T))

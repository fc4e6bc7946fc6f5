"""Dense (direct) visual odometry -- same classes, attributes and defaults as reference pyslam/pipelines/dense.py.

`_compute_frame_to_frame_motion` is the caller of the photometric hot path: a coarse-to-fine loop
over the pyramid levels, one `Problem` per level on the (SO3, t) parameter form with the translation held constant
on the coarsest levels (dense.py:152-205).  Here every level runs on the GPU through the SAME engine handle
(`Problem(options, engine=...)`: device context, stream and buffers are created once per pipeline), the (SO3, t) form is
lowered to csrc/photometric.cuh, and the pyramids come from csrc/image.cuh.
"""
import numpy as np

from .. import engine as _engine
from ..lie import SE3
from ..losses import HuberLoss
from ..problem import Options, Problem
from ..residuals import PhotometricResidualSE3
from .keyframes import DenseRGBDKeyframe, DenseStereoKeyframe


class DenseVOPipeline:
    """Base class for dense VO pipelines"""

    def __init__(self, camera, first_pose=None, device=0):
        self.camera = camera
        self.first_pose = SE3.identity() if first_pose is None else first_pose
        self.keyframes = []
        self.T_c_w = [self.first_pose]
        self.device = device

        o = Options()                         # defaults of dense.py:27-36
        o.allow_nondecreasing_steps = True
        o.max_nondecreasing_steps = 5
        o.min_cost_decrease = 0.99
        o.max_iters = 30
        o.num_threads = 1
        o.linesearch_max_iters = 0
        o.device = device
        self.motion_options = o

        self.pyrlevels = 4
        self.pyrlevel_sequence = list(range(self.pyrlevels))[::-1]
        self.keyframe_trans_thresh = 3.0      # meters
        self.keyframe_rot_thresh = 0.3        # rad
        self.intensity_stiffness = 1. / 0.01
        self.depth_stiffness = 1. / 0.01
        self.min_grad = 0.1
        self.depth_map_type = 'depth'         # 'depth' | 'disparity'
        self.mode = 'map'                     # 'map' | 'track'
        self.use_motion_model_guess = True
        self.loss = HuberLoss(10.0)
        self._engine = None                   # one engine handle for all levels and frames
        self.level_summaries = []             # (pyrlevel, iterations, initial cost, final cost) of the last motion estimate
        self._make_pyramid_cameras()

    def _make_pyramid_cameras(self):
        self.pyr_cameras = []
        for pyrlevel in self.pyrlevel_sequence:
            f = 2. ** -pyrlevel
            cam = self.camera.clone()
            cam.fu *= f; cam.fv *= f; cam.cu *= f; cam.cv *= f
            cam.h = int(np.ceil(cam.h * f))
            cam.w = int(np.ceil(cam.w * f))
            cam.compute_pixel_grid()
            self.pyr_cameras.append(cam)

    def set_mode(self, mode):
        """Set the localization mode to ['map'|'track']"""
        self.mode = mode
        if self.mode == 'track':
            self.active_keyframe_idx = 0
            self.T_c_w = []

    def track(self, trackframe, guess=None):
        """Track an image against the active keyframe (dense.py:88-150)."""
        if len(self.keyframes) == 0:
            trackframe.compute_pyramids()
            self.keyframes.append(trackframe)
            self.active_keyframe_idx = 0
            return
        active = self.keyframes[self.active_keyframe_idx]
        if guess is None:
            if len(self.T_c_w) == 0:
                guess = SE3.identity()
            else:
                guess = self.T_c_w[-1].dot(active.T_c_w.inv())
            if self.use_motion_model_guess and len(self.T_c_w) > 1:
                guess = self.T_c_w[-1].dot(self.T_c_w[-2].inv().dot(guess))
        else:
            guess = guess.dot(active.T_c_w.inv())
        T_track_ref = self._compute_frame_to_frame_motion(active, trackframe, guess)
        T_track_ref.normalize()
        self.T_c_w.append(T_track_ref.dot(active.T_c_w))
        xi = T_track_ref.log()
        trans_dist, rot_dist = np.linalg.norm(xi[0:3]), np.linalg.norm(xi[3:6])
        if trans_dist > self.keyframe_trans_thresh or rot_dist > self.keyframe_rot_thresh:
            if self.mode == 'map':
                trackframe.T_c_w = self.T_c_w[-1]
                trackframe.compute_pyramids()
                self.keyframes.append(trackframe)
            self.active_keyframe_idx += 1

    def _compute_frame_to_frame_motion(self, ref_frame, track_frame, guess=None):
        guess = SE3.identity() if guess is None else guess
        params = {'R_1_0': guess.rot, 't_1_0_1': guess.trans}
        if self._engine is None:
            self._engine = _engine.Engine(self.device)
        self.level_summaries = []
        for pyrlevel, pyr_camera in zip(self.pyrlevel_sequence, self.pyr_cameras):
            pyrfactor = 2. ** -pyrlevel
            if self.depth_map_type == 'disparity':
                depth_ref = ref_frame.disparity[pyrlevel]
                depth_stiffness = self.depth_stiffness / pyrfactor     # disparities are in pixels of the level
            else:
                depth_ref = ref_frame.depth[pyrlevel]
                depth_stiffness = self.depth_stiffness
            residual = PhotometricResidualSE3(pyr_camera, ref_frame.im_pyr[pyrlevel], depth_ref, track_frame.im_pyr[pyrlevel],
                                              ref_frame.jacobian[pyrlevel], self.intensity_stiffness, depth_stiffness,
                                              self.min_grad)
            problem = Problem(self.motion_options, engine=self._engine)
            problem.add_residual_block(residual, ['R_1_0', 't_1_0_1'], loss=self.loss)
            problem.initialize_params(params)
            if pyrlevel > 2:
                problem.set_parameters_constant('t_1_0_1')
            params = problem.solve()
            h = problem._cost_history
            self.level_summaries.append((pyrlevel, len(h) - 1, h[0], h[-1]))
        return SE3(params['R_1_0'], params['t_1_0_1'])


class DenseStereoPipeline(DenseVOPipeline):
    """Dense stereo VO pipeline"""

    def __init__(self, camera, first_pose=None, device=0):
        super().__init__(camera, first_pose, device)
        self.depth_map_type = 'disparity'
        self.depth_stiffness = 1 / 0.5

    def track(self, im_left, im_right, guess=None, disparity=None):
        T = self.T_c_w[0] if len(self.keyframes) == 0 else None
        super().track(DenseStereoKeyframe(im_left, im_right, self.pyrlevels, T, self.device, disparity=disparity), guess)


class DenseRGBDPipeline(DenseVOPipeline):
    """Dense RGBD VO pipeline"""

    def __init__(self, camera, first_pose=None, device=0):
        super().__init__(camera, first_pose, device)
        self.depth_map_type = 'depth'
        self.depth_stiffness = 1 / 0.01

    def track(self, image, depth, guess=None):
        T = self.T_c_w[0] if len(self.keyframes) == 0 else None
        super().track(DenseRGBDKeyframe(image, depth, self.pyrlevels, T, self.device), guess)

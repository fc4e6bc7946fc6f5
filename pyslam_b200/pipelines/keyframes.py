"""Keyframes of the VO pipelines -- same classes and attributes as reference pyslam/pipelines/keyframes.py.

The image pyramid (cv2.pyrDown chain on the 8-bit image, then / 255), the gradient pyramid (0.5 * Sobel) and the
disparity / depth pyramids (sub-sampling, disparities scaled by 2^-level) are computed on the GPU
(`bslam_image_pyramid`, `bslam_subsample_pyramid`, csrc/image.cuh).  Stereo matching itself (cv2.StereoBM) is
outside the hot path and stays on the CPU, as SURVEY 8 f3 says; a precomputed disparity map can be passed instead.
"""
import numpy as np

from .. import engine as _engine
from ..lie import SE3


class Keyframe:
    """Keyframe base class"""

    def __init__(self, data, T_c_w=None):
        self.data = data
        self.T_c_w = SE3.identity() if T_c_w is None else T_c_w


class DenseKeyframe(Keyframe):
    """Dense keyframe base class"""

    def __init__(self, data, pyrimage, pyrlevels, T_c_w=None, device=0):
        super().__init__(data, T_c_w)
        self.pyrlevels = pyrlevels
        self.device = device
        self.compute_image_pyramid(pyrimage)

    def compute_image_pyramid(self, pyrimage):
        """Image AND gradient pyramids in one device call (the gradients are cheap once the levels are resident)."""
        self.im_pyr, self._jacobian = _engine.image_pyramid(pyrimage, max(self.pyrlevels, 1), gradients=True, device=self.device)
        if self.pyrlevels == 0:        # range(0) in the reference: no levels at all
            self.im_pyr, self._jacobian = [], []

    def compute_jacobian_pyramid(self):
        self.jacobian = self._jacobian


class DenseRGBDKeyframe(DenseKeyframe):
    """Dense RGBD keyframe"""

    def __init__(self, image, depth, pyrlevels=0, T_c_w=None, device=0):
        super().__init__((image, depth), image, pyrlevels, T_c_w, device)

    def compute_depth_pyramid(self):
        self.depth = _engine.subsample_pyramid(self.data[1], self.pyrlevels, 1., self.device) if self.pyrlevels else []

    def compute_pyramids(self):
        self.compute_jacobian_pyramid()
        self.compute_depth_pyramid()


class DenseStereoKeyframe(DenseKeyframe):
    """Dense Stereo keyframe.  `disparity`: optional precomputed full-resolution disparity map (pixels); without it
    cv2.StereoBM is run on the CPU as in the reference."""

    def __init__(self, im_left, im_right, pyrlevels=0, T_c_w=None, device=0, disparity=None):
        super().__init__((im_left, im_right), im_left, pyrlevels, T_c_w, device)
        self._disp0 = disparity

    @property
    def im_left(self):
        return self.data[0]

    @property
    def im_right(self):
        return self.data[1]

    def compute_disparity_pyramid(self):
        disp = self._disp0
        if disp is None:
            import cv2
            disp = cv2.StereoBM_create().compute(self.im_left, self.im_right).astype(float) / 16.
        self.disparity = _engine.subsample_pyramid(disp, self.pyrlevels, 0.5, self.device) if self.pyrlevels else []

    def compute_pyramids(self):
        self.compute_jacobian_pyramid()
        self.compute_disparity_pyramid()


class SparseStereoKeyframe(Keyframe):
    """Sparse Stereo keyframe"""

    def __init__(self, im_left, im_right, T_c_w=None):
        super().__init__((im_left, im_right), T_c_w)

    @property
    def im_left(self):
        return self.data[0]

    @property
    def im_right(self):
        return self.data[1]


class SparseRGBDKeyframe(Keyframe):
    """Sparse RGB-D keyframe"""

    def __init__(self, image, depth, T_c_w=None):
        super().__init__((image, depth), T_c_w)

    @property
    def image(self):
        return self.data[0]

    @property
    def depth(self):
        return self.data[1]

"""Camera models (reference pyslam/sensors/)."""
from .stereo_camera import StereoCamera

__all__ = ['StereoCamera']

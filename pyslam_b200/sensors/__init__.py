"""Camera models (reference pyslam/sensors/)."""
from .rgbd_camera import RGBDCamera
from .stereo_camera import StereoCamera

__all__ = ['StereoCamera', 'RGBDCamera']

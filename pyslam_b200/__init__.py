"""pyslam_b200 -- B200-native Gauss-Newton/LM solver behind pyslam's API.

Drop-in for `pyslam.problem` (Options, Problem), `pyslam.residuals`,
`pyslam.losses`, `pyslam.sensors`, `pyslam.utils`; the solve loop runs in
hand-written sm_100a CUDA behind the C ABI declared in include/bslam.h.
"""
from .problem import Options, Problem          # noqa: F401

__version__ = '0.1.0'

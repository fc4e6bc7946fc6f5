"""Residual plug-ins (reference pyslam/residuals/).

`from pyslam_b200.residuals import ReprojectionResidual, PoseResidual, ...`
resolves the same names as `from pyslam.residuals import ...`.
"""
from .blocks import (ReprojectionResidual, PoseResidual, PoseToPoseResidual,
                     QuadraticResidual)
from .motion_only import (PoseToPoseOrientationResidual, ReprojectionMotionOnlyBatchResidual,
                          ReprojectionMotionOnlyResidual)
from .photometric import PhotometricResidualSE3

__all__ = ['ReprojectionResidual', 'PoseResidual', 'PoseToPoseResidual',
           'QuadraticResidual', 'PhotometricResidualSE3', 'ReprojectionMotionOnlyResidual',
           'ReprojectionMotionOnlyBatchResidual', 'PoseToPoseOrientationResidual']

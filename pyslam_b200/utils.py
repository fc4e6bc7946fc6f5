"""Host-side helpers with the reference's names (pyslam/utils.py:8-89)."""
import numpy as np
import scipy.linalg


def invsqrt(x):
    """Inverse square root of a scalar or of a square matrix (stiffness =
    covariance^{-1/2}); reference pyslam/utils.py:8-13."""
    if hasattr(x, 'shape'):
        return np.linalg.inv(scipy.linalg.sqrtm(x))
    return 1. / np.sqrt(x)


def bilinear_interpolate(im, x, y):
    """Bilinear image sampling with the reference's border semantics
    (pyslam/utils.py:16-77): indices truncate toward zero, the four weights are
    formed before the indices are clamped, clamping repeats the border."""
    im = np.atleast_3d(np.asarray(im, dtype=float))
    x = np.asarray(x, dtype=float)
    y = np.asarray(y, dtype=float)
    xi, yi = np.trunc(x).astype(np.int64), np.trunc(y).astype(np.int64)
    fx, fy = x - xi, y - yi
    xa, xb = np.clip(xi, 0, im.shape[1] - 1), np.clip(xi + 1, 0, im.shape[1] - 1)
    ya, yb = np.clip(yi, 0, im.shape[0] - 1), np.clip(yi + 1, 0, im.shape[0] - 1)
    out = ((1. - fx) * (1. - fy))[:, None] * im[ya, xa] + ((1. - fx) * fy)[:, None] * im[yb, xa] \
        + (fx * (1. - fy))[:, None] * im[ya, xb] + (fx * fy)[:, None] * im[yb, xb]
    return np.squeeze(out)


def stackmul(A, B):
    """Batched small matrix product (pyslam/utils.py:80-89)."""
    return np.matmul(A, B)

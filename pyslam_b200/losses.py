"""Robust-loss plug-ins (same names / constructor arguments / methods as
reference pyslam/losses.py:8-214).

Each class exposes element-wise `loss`, `influence`, `weight` on numpy arrays,
which is the duck-typed protocol `Problem` accepts for *any* loss object.  The
classes below additionally carry `LOSS_KIND`, which lets `Problem.solve()` fuse
the re-weighting into the CUDA linearisation kernels (csrc/loss.cuh) instead
of calling back into Python.  Formulas are reproduced as the reference writes
them (e.g. Tukey influence/weight are not squared, losses.py:141-175).
"""
import numpy as np

# numeric ids shared with include/bslam.h (BSLAM_LOSS_*)
LOSS_L2, LOSS_L1, LOSS_CAUCHY, LOSS_HUBER, LOSS_TUKEY, LOSS_TDIST = range(6)


class _Loss:
    LOSS_KIND = None
    k = 0.

    def __repr__(self):
        return '{}(k={})'.format(type(self).__name__, self.k)


class _ScaledLoss(_Loss):
    def __init__(self, k):
        self.k = k


class L2Loss(_Loss):
    LOSS_KIND = LOSS_L2

    def loss(self, x):
        return 0.5 * x * x

    def influence(self, x):
        return x

    def weight(self, x):
        return np.ones(np.size(x))


class L1Loss(_Loss):
    LOSS_KIND = LOSS_L1

    def loss(self, x):
        return np.abs(x)

    def influence(self, x):
        x = np.asarray(x, dtype=float)
        out = np.sign(x)
        out[np.abs(x) <= 1e-8] = np.nan
        return out

    def weight(self, x):
        x = np.asarray(x, dtype=float)
        with np.errstate(divide='ignore'):
            out = 1. / np.abs(x)
        out[np.abs(x) <= 1e-8] = np.nan
        return out


class CauchyLoss(_ScaledLoss):
    LOSS_KIND = LOSS_CAUCHY

    def loss(self, x):
        q = np.asarray(x, dtype=float) / self.k
        return (0.5 * self.k ** 2) * np.log(1. + q * q)

    def influence(self, x):
        q = np.asarray(x, dtype=float) / self.k
        return x / (1. + q * q)

    def weight(self, x):
        q = np.asarray(x, dtype=float) / self.k
        return 1. / (1. + q * q)


class HuberLoss(_ScaledLoss):
    LOSS_KIND = LOSS_HUBER

    def loss(self, x):
        x = np.asarray(x, dtype=float)
        a = np.abs(x)
        return np.where(a <= self.k, 0.5 * x * x, self.k * (a - 0.5 * self.k))

    def influence(self, x):
        # the reference returns the ufunc object here (losses.py:83-84, a bug;
        # `Problem` never calls influence) -- this returns the values instead.
        x = np.asarray(x, dtype=float)
        return np.where(np.abs(x) <= self.k, x, self.k * np.sign(x))

    def weight(self, x):
        a = np.abs(np.asarray(x, dtype=float))
        with np.errstate(divide='ignore', invalid='ignore'):
            return np.where(a <= self.k, 1., self.k / a)


class TukeyLoss(_ScaledLoss):
    LOSS_KIND = LOSS_TUKEY

    def loss(self, x):
        x = np.asarray(x, dtype=float)
        c = self.k ** 2 / 6.
        q = x / self.k
        return np.where(np.abs(x) <= self.k, c * (1. - (1. - q * q) ** 3), c)

    def influence(self, x):
        x = np.asarray(x, dtype=float)
        q = x / self.k
        return np.where(np.abs(x) <= self.k, x * (1. - q * q), 0.)

    def weight(self, x):
        x = np.asarray(x, dtype=float)
        q = x / self.k
        return np.where(np.abs(x) <= self.k, 1. - q * q, 0.)


class TDistributionLoss(_ScaledLoss):
    LOSS_KIND = LOSS_TDIST

    def loss(self, x):
        x = np.asarray(x, dtype=float)
        return 0.5 * (self.k + 1.) * np.log(1. + x * x / self.k)

    def influence(self, x):
        x = np.asarray(x, dtype=float)
        return (self.k + 1.) * x / (self.k + x * x)

    def weight(self, x):
        x = np.asarray(x, dtype=float)
        return (self.k + 1.) / (self.k + x * x)


def loss_descriptor(loss):
    """(kind id, k) for built-in losses, None for user-defined plug-ins."""
    kind = getattr(type(loss), 'LOSS_KIND', None)
    if kind is None or type(loss).__module__ != __name__:
        return None
    return int(kind), float(getattr(loss, 'k', 0.))

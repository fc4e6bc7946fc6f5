"""Host-side SO(2)/SE(2)/SO(3)/SE(3) parameter containers.

pyslam takes its parameter type from the external `liegroups` package
(reference: setup.py:12; call sites pyslam/problem.py:258-260,406,
residuals/pose_residual.py:10-23, pose_to_pose_residual.py:10-28,
reprojection_residual.py:16-31).  That package is not part of this image, so the
drop-in ships its own containers with the same attribute/method surface
(`rot`, `trans`, `mat`, `dof`, `dim`, `exp`, `log`, `dot`, `inv`, `adjoint`,
`odot`, `perturb`, `as_matrix`, `from_matrix`, `identity`, `normalize`).  They
are used to *describe* a problem (initial values, measurements) and to hand
results back; during `Problem.solve()` the poses live on the GPU as packed
[R|t] rows and every exp/log/adjoint/retract on the hot path runs in
csrc/lie.cuh.  Objects from a real `liegroups` install are accepted as well
(duck-typed on `.rot.mat` / `.trans`, see pyslam_b200/problem.py: Problem._lower).

Conventions: tangent order [rho; phi]; left
perturbation T <- exp(xi) T; small-angle branch when |angle| <= 1e-8.
"""
import numpy as np

_SMALL = 1e-8          # np.isclose(x, 0.) with default tolerances


def _skew3(v):
    x, y, z = v
    return np.array([[0., -z, y], [z, 0., -x], [-y, x, 0.]])


_J2 = np.array([[0., -1.], [1., 0.]])      # so(2) generator


class _Rotation:
    """Shared behaviour of SO2/SO3: a wrapped orthonormal matrix `mat`."""
    __slots__ = ('mat',)

    def __init__(self, mat):
        self.mat = np.array(mat, dtype=float)

    @classmethod
    def identity(cls):
        return cls(np.eye(cls.dim))

    @classmethod
    def from_matrix(cls, mat, normalize=False):
        obj = cls(mat)
        if normalize:
            obj.normalize()
        return obj

    def as_matrix(self):
        return self.mat

    def inv(self):
        return type(self)(self.mat.T)

    def dot(self, other):
        if isinstance(other, _Rotation):
            return type(self)(self.mat @ other.mat)
        v = np.asarray(other, dtype=float)
        if v.shape[-1] != self.dim:
            raise ValueError('expected vectors of length {}'.format(self.dim))
        return v @ self.mat.T

    def perturb(self, phi):
        self.mat = type(self).exp(phi).mat @ self.mat

    def normalize(self):
        u, _, vt = np.linalg.svd(self.mat)
        d = np.ones(self.dim)
        d[-1] = np.linalg.det(u) * np.linalg.det(vt)
        self.mat = (u * d) @ vt

    def __repr__(self):
        return '{}(\n{})'.format(type(self).__name__, self.mat)


class SO2(_Rotation):
    dim, dof = 2, 1

    @classmethod
    def from_angle(cls, angle):
        c, s = np.cos(angle), np.sin(angle)
        return cls([[c, -s], [s, c]])

    @classmethod
    def exp(cls, phi):
        return cls.from_angle(float(np.squeeze(phi)))

    @staticmethod
    def wedge(phi):
        return float(np.squeeze(phi)) * _J2

    @staticmethod
    def left_jacobian(phi):
        phi = float(np.squeeze(phi))
        if abs(phi) <= _SMALL:
            return np.eye(2) + 0.5 * phi * _J2
        return (np.sin(phi) / phi) * np.eye(2) + ((1. - np.cos(phi)) / phi) * _J2

    @staticmethod
    def inv_left_jacobian(phi):
        phi = float(np.squeeze(phi))
        if abs(phi) <= _SMALL:
            return np.eye(2) - 0.5 * phi * _J2
        h = 0.5 * phi
        return (h / np.tan(h)) * np.eye(2) - h * _J2

    def log(self):
        return np.arctan2(self.mat[1, 0], self.mat[0, 0])

    to_angle = log

    def adjoint(self):
        return 1.


class SO3(_Rotation):
    dim, dof = 3, 3

    @classmethod
    def _axis(cls, axis, angle):
        c, s = np.cos(angle), np.sin(angle)
        m = np.eye(3)
        i, j = [(1, 2), (2, 0), (0, 1)][axis]
        m[i, i] = m[j, j] = c
        m[i, j], m[j, i] = -s, s
        return cls(m)

    @classmethod
    def rotx(cls, a):
        return cls._axis(0, a)

    @classmethod
    def roty(cls, a):
        return cls._axis(1, a)

    @classmethod
    def rotz(cls, a):
        return cls._axis(2, a)

    @staticmethod
    def wedge(phi):
        phi = np.asarray(phi, dtype=float)
        if phi.ndim == 1:
            return _skew3(phi)
        return np.stack([_skew3(p) for p in phi])

    @staticmethod
    def vee(m):
        return np.array([m[2, 1], m[0, 2], m[1, 0]])

    @classmethod
    def exp(cls, phi):
        phi = np.asarray(phi, dtype=float)
        th = np.linalg.norm(phi)
        if th <= _SMALL:
            return cls(np.eye(3) + _skew3(phi))
        a = phi / th
        c, s = np.cos(th), np.sin(th)
        return cls(c * np.eye(3) + (1. - c) * np.outer(a, a) + s * _skew3(a))

    @staticmethod
    def left_jacobian(phi):
        phi = np.asarray(phi, dtype=float)
        th = np.linalg.norm(phi)
        if th <= _SMALL:
            return np.eye(3) + 0.5 * _skew3(phi)
        a = phi / th
        sth = np.sin(th) / th
        return sth * np.eye(3) + (1. - sth) * np.outer(a, a) + ((1. - np.cos(th)) / th) * _skew3(a)

    @staticmethod
    def inv_left_jacobian(phi):
        phi = np.asarray(phi, dtype=float)
        th = np.linalg.norm(phi)
        if th <= _SMALL:
            return np.eye(3) - 0.5 * _skew3(phi)
        a = phi / th
        h = 0.5 * th
        hc = h / np.tan(h)
        return hc * np.eye(3) + (1. - hc) * np.outer(a, a) - h * _skew3(a)

    def log(self):
        c = min(1., max(-1., 0.5 * np.trace(self.mat) - 0.5))
        th = np.arccos(c)
        if th <= _SMALL:
            return self.vee(self.mat - np.eye(3))
        return self.vee((0.5 * th / np.sin(th)) * (self.mat - self.mat.T))

    def adjoint(self):
        return self.mat


class _RigidTransform:
    """Shared behaviour of SE2/SE3: (`rot`, `trans`)."""
    __slots__ = ('rot', 'trans')

    def __init__(self, rot, trans):
        self.rot = rot
        self.trans = np.array(trans, dtype=float)

    @classmethod
    def identity(cls):
        return cls(cls.RotationType.identity(), np.zeros(cls.dim - 1))

    @classmethod
    def from_matrix(cls, mat, normalize=False):
        n = cls.dim - 1
        mat = np.asarray(mat, dtype=float)
        return cls(cls.RotationType.from_matrix(mat[:n, :n], normalize), mat[:n, n])

    @classmethod
    def exp(cls, xi):
        xi = np.asarray(xi, dtype=float)
        n = cls.dim - 1
        rho, phi = xi[:n], xi[n:]
        return cls(cls.RotationType.exp(phi), cls.RotationType.left_jacobian(phi) @ rho)

    def log(self):
        phi = self.rot.log()
        return np.hstack([self.RotationType.inv_left_jacobian(phi) @ self.trans, phi])

    def as_matrix(self):
        n = self.dim - 1
        m = np.eye(self.dim)
        m[:n, :n] = self.rot.mat
        m[:n, n] = self.trans
        return m

    def inv(self):
        rt = self.rot.mat.T
        return type(self)(self.RotationType(rt), -(rt @ self.trans))

    def dot(self, other):
        if isinstance(other, _RigidTransform):
            return type(self)(self.RotationType(self.rot.mat @ other.rot.mat),
                              self.rot.mat @ other.trans + self.trans)
        v = np.asarray(other, dtype=float)
        if v.shape[-1] == self.dim - 1:
            return v @ self.rot.mat.T + self.trans
        if v.shape[-1] == self.dim:
            return v @ self.as_matrix().T
        raise ValueError('expected vectors of length {} or {}'.format(self.dim - 1, self.dim))

    __mul__ = dot

    def perturb(self, xi):
        e = type(self).exp(xi)
        self.trans = e.rot.mat @ self.trans + e.trans
        self.rot = self.RotationType(e.rot.mat @ self.rot.mat)

    def normalize(self):
        self.rot.normalize()

    def __repr__(self):
        return '{}(\n{})'.format(type(self).__name__, self.as_matrix())


class SE2(_RigidTransform):
    dim, dof = 3, 3
    RotationType = SO2

    def adjoint(self):
        ad = np.eye(3)
        ad[:2, :2] = self.rot.mat
        ad[0, 2], ad[1, 2] = self.trans[1], -self.trans[0]
        return ad


class SE3(_RigidTransform):
    dim, dof = 4, 6
    RotationType = SO3

    def adjoint(self):
        r = self.rot.mat
        ad = np.zeros((6, 6))
        ad[:3, :3] = ad[3:, 3:] = r
        ad[:3, 3:] = _skew3(self.trans) @ r
        return ad

    @staticmethod
    def odot(p, directional=False):
        """[I | -p^] per point: (3,) -> (3,6), (N,3) -> (N,3,6); homogeneous
        4-vectors scale the identity block by their last entry."""
        p = np.atleast_2d(np.asarray(p, dtype=float))
        out = np.zeros((p.shape[0], 3, 6))
        scale = p[:, 3] if p.shape[1] == 4 else np.full(p.shape[0], 0. if directional else 1.)
        x, y, z = p[:, 0], p[:, 1], p[:, 2]
        for i in range(3):
            out[:, i, i] = scale
        out[:, 0, 4], out[:, 0, 5] = z, -y
        out[:, 1, 3], out[:, 1, 5] = -z, x
        out[:, 2, 3], out[:, 2, 4] = y, -x
        return np.squeeze(out)


def group_of(obj):
    """'se3' / 'se2' / 'so3' / 'so2' for liegroups-like objects (ours or the real
    package's), else None.  Duck-typed so a real `liegroups` install works."""
    rot = getattr(obj, 'rot', None)
    if rot is not None and hasattr(rot, 'mat') and hasattr(obj, 'trans'):
        n = np.shape(rot.mat)[0]
        return {2: 'se2', 3: 'se3'}.get(n)
    if hasattr(obj, 'mat') and hasattr(obj, 'dof') and hasattr(obj, 'perturb'):
        return {2: 'so2', 3: 'so3'}.get(np.shape(obj.mat)[0])
    return None

"""SO(3) -- oracle restatement (test infrastructure; see package docstring)."""
import numpy as np


class SO3:
    dim = 3
    dof = 3

    def __init__(self, mat):
        self.mat = np.array(mat, dtype=float)

    @classmethod
    def identity(cls):
        return cls(np.identity(cls.dim))

    @classmethod
    def from_matrix(cls, mat, normalize=False):
        out = cls(mat)
        if normalize:
            out.normalize()
        return out

    @classmethod
    def rotx(cls, a):
        c, s = np.cos(a), np.sin(a)
        return cls(np.array([[1., 0., 0.], [0., c, -s], [0., s, c]]))

    @classmethod
    def roty(cls, a):
        c, s = np.cos(a), np.sin(a)
        return cls(np.array([[c, 0., s], [0., 1., 0.], [-s, 0., c]]))

    @classmethod
    def rotz(cls, a):
        c, s = np.cos(a), np.sin(a)
        return cls(np.array([[c, -s, 0.], [s, c, 0.], [0., 0., 1.]]))

    @classmethod
    def wedge(cls, phi):
        phi = np.atleast_2d(phi)
        if phi.shape[1] != cls.dof:
            raise ValueError('phi must have shape (3,) or (N,3)')
        Phi = np.zeros([phi.shape[0], cls.dim, cls.dim])
        Phi[:, 0, 1] = -phi[:, 2]
        Phi[:, 1, 0] = phi[:, 2]
        Phi[:, 0, 2] = phi[:, 1]
        Phi[:, 2, 0] = -phi[:, 1]
        Phi[:, 1, 2] = -phi[:, 0]
        Phi[:, 2, 1] = phi[:, 0]
        return np.squeeze(Phi)

    @classmethod
    def vee(cls, Phi):
        return np.array([Phi[2, 1], Phi[0, 2], Phi[1, 0]])

    @classmethod
    def exp(cls, phi):
        phi = np.asarray(phi, dtype=float)
        angle = np.linalg.norm(phi)
        if np.isclose(angle, 0.):
            return cls(np.identity(cls.dim) + cls.wedge(phi))
        axis = phi / angle
        s, c = np.sin(angle), np.cos(angle)
        return cls(c * np.identity(cls.dim) + (1. - c) * np.outer(axis, axis) + s * cls.wedge(axis))

    @classmethod
    def left_jacobian(cls, phi):
        phi = np.asarray(phi, dtype=float)
        angle = np.linalg.norm(phi)
        if np.isclose(angle, 0.):
            return np.identity(cls.dof) + 0.5 * cls.wedge(phi)
        axis = phi / angle
        s, c = np.sin(angle), np.cos(angle)
        return (s / angle) * np.identity(cls.dof) + \
            (1. - s / angle) * np.outer(axis, axis) + \
            ((1. - c) / angle) * cls.wedge(axis)

    @classmethod
    def inv_left_jacobian(cls, phi):
        phi = np.asarray(phi, dtype=float)
        angle = np.linalg.norm(phi)
        if np.isclose(angle, 0.):
            return np.identity(cls.dof) - 0.5 * cls.wedge(phi)
        axis = phi / angle
        half = 0.5 * angle
        cot_half = 1. / np.tan(half)
        return half * cot_half * np.identity(cls.dof) + \
            (1. - half * cot_half) * np.outer(axis, axis) - \
            half * cls.wedge(axis)

    def log(self):
        cos_angle = np.clip(0.5 * np.trace(self.mat) - 0.5, -1., 1.)
        angle = np.arccos(cos_angle)
        if np.isclose(angle, 0.):
            return self.vee(self.mat - np.identity(3))
        return self.vee((0.5 * angle / np.sin(angle)) * (self.mat - self.mat.T))

    def as_matrix(self):
        return self.mat

    def inv(self):
        return self.__class__(self.mat.T)

    def adjoint(self):
        return self.mat

    def dot(self, other):
        if isinstance(other, self.__class__):
            return self.__class__(self.mat.dot(other.mat))
        other = np.atleast_2d(other)
        if other.shape[1] != self.dim:
            raise ValueError('vector must have shape (3,) or (N,3)')
        return np.squeeze(self.mat.dot(other.T).T)

    def perturb(self, phi):
        self.mat = self.__class__.exp(phi).dot(self).mat

    def normalize(self):
        U, _, Vt = np.linalg.svd(self.mat, full_matrices=False)
        mid = np.identity(self.dim)
        mid[self.dim - 1, self.dim - 1] = np.linalg.det(U) * np.linalg.det(Vt)
        self.mat = U.dot(mid).dot(Vt)

    def __repr__(self):
        return '<{}.{}>\n{}'.format(self.__class__.__module__, self.__class__.__name__, self.mat)

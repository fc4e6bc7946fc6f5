"""SO(2) -- oracle restatement (test infrastructure; see package docstring)."""
import numpy as np


class SO2:
    dim = 2
    dof = 1

    def __init__(self, mat):
        self.mat = np.array(mat, dtype=float)

    # -- constructors ------------------------------------------------------
    @classmethod
    def identity(cls):
        return cls(np.identity(cls.dim))

    @classmethod
    def from_angle(cls, angle):
        c, s = np.cos(angle), np.sin(angle)
        return cls(np.array([[c, -s], [s, c]]))

    @classmethod
    def from_matrix(cls, mat, normalize=False):
        out = cls(mat)
        if normalize:
            out.normalize()
        return out

    @classmethod
    def exp(cls, phi):
        return cls.from_angle(float(np.squeeze(phi)))

    # -- algebra -----------------------------------------------------------
    @classmethod
    def wedge(cls, phi):
        phi = float(np.squeeze(phi))
        return np.array([[0., -phi], [phi, 0.]])

    @classmethod
    def vee(cls, Phi):
        return Phi[1, 0]

    @classmethod
    def left_jacobian(cls, phi):
        phi = float(np.squeeze(phi))
        if np.isclose(phi, 0.):
            return np.identity(cls.dim) + 0.5 * cls.wedge(phi)
        s, c = np.sin(phi), np.cos(phi)
        return (s / phi) * np.identity(cls.dim) + ((1. - c) / phi) * cls.wedge(1.)

    @classmethod
    def inv_left_jacobian(cls, phi):
        phi = float(np.squeeze(phi))
        if np.isclose(phi, 0.):
            return np.identity(cls.dim) - 0.5 * cls.wedge(phi)
        half = 0.5 * phi
        cot_half = 1. / np.tan(half)
        return half * cot_half * np.identity(cls.dim) - half * cls.wedge(1.)

    # -- group -------------------------------------------------------------
    def log(self):
        return np.arctan2(self.mat[1, 0], self.mat[0, 0])

    def to_angle(self):
        return self.log()

    def as_matrix(self):
        return self.mat

    def inv(self):
        return self.__class__(self.mat.T)

    def adjoint(self):
        return 1.

    def dot(self, other):
        if isinstance(other, self.__class__):
            return self.__class__(self.mat.dot(other.mat))
        other = np.atleast_2d(other)
        if other.shape[1] != self.dim:
            raise ValueError('vector must have shape ({},) or (N,{})'.format(self.dim, self.dim))
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

"""SE(2) -- oracle restatement (test infrastructure; see package docstring)."""
import numpy as np
from .so2 import SO2


class SE2:
    dim = 3
    dof = 3
    RotationType = SO2

    def __init__(self, rot, trans):
        self.rot = rot
        self.trans = np.array(trans, dtype=float)

    @classmethod
    def identity(cls):
        return cls(SO2.identity(), np.zeros(2))

    @classmethod
    def from_matrix(cls, mat, normalize=False):
        return cls(SO2.from_matrix(mat[0:2, 0:2], normalize), mat[0:2, 2])

    @classmethod
    def exp(cls, xi):
        xi = np.asarray(xi, dtype=float)
        rho, phi = xi[0:2], xi[2]
        return cls(SO2.exp(phi), SO2.left_jacobian(phi).dot(rho))

    @classmethod
    def wedge(cls, xi):
        xi = np.asarray(xi, dtype=float)
        Xi = np.zeros((3, 3))
        Xi[0:2, 0:2] = SO2.wedge(xi[2])
        Xi[0:2, 2] = xi[0:2]
        return Xi

    def log(self):
        phi = SO2.log(self.rot)
        rho = SO2.inv_left_jacobian(phi).dot(self.trans)
        return np.hstack([rho, phi])

    def as_matrix(self):
        T = np.identity(3)
        T[0:2, 0:2] = self.rot.as_matrix()
        T[0:2, 2] = self.trans
        return T

    def inv(self):
        inv_rot = self.rot.inv()
        return self.__class__(inv_rot, -(inv_rot.dot(self.trans)))

    def adjoint(self):
        Ad = np.identity(3)
        Ad[0:2, 0:2] = self.rot.as_matrix()
        Ad[0, 2] = self.trans[1]
        Ad[1, 2] = -self.trans[0]
        return Ad

    def dot(self, other):
        if isinstance(other, self.__class__):
            return self.__class__(self.rot.dot(other.rot),
                                  self.rot.dot(other.trans) + self.trans)
        other = np.atleast_2d(other)
        if other.shape[1] == self.dim - 1:
            return np.squeeze(self.rot.dot(other) + self.trans)
        raise ValueError('vector must have shape (2,) or (N,2)')

    def perturb(self, xi):
        p = self.__class__.exp(xi).dot(self)
        self.rot = p.rot
        self.trans = p.trans

    def normalize(self):
        self.rot.normalize()

    def __repr__(self):
        return '<{}.{}>\n{}'.format(self.__class__.__module__, self.__class__.__name__, self.as_matrix())

"""SE(3) -- oracle restatement (test infrastructure; see package docstring)."""
import numpy as np
from .so3 import SO3


class SE3:
    dim = 4
    dof = 6
    RotationType = SO3

    def __init__(self, rot, trans):
        self.rot = rot
        self.trans = np.array(trans, dtype=float)

    @classmethod
    def identity(cls):
        return cls(SO3.identity(), np.zeros(3))

    @classmethod
    def from_matrix(cls, mat, normalize=False):
        return cls(SO3.from_matrix(mat[0:3, 0:3], normalize), mat[0:3, 3])

    @classmethod
    def exp(cls, xi):
        xi = np.asarray(xi, dtype=float)
        rho, phi = xi[0:3], xi[3:6]
        return cls(SO3.exp(phi), SO3.left_jacobian(phi).dot(rho))

    @classmethod
    def wedge(cls, xi):
        xi = np.asarray(xi, dtype=float)
        Xi = np.zeros((4, 4))
        Xi[0:3, 0:3] = SO3.wedge(xi[3:6])
        Xi[0:3, 3] = xi[0:3]
        return Xi

    @classmethod
    def odot(cls, p, directional=False):
        """(N,3)/(3,) -> (N,3,6)/(3,6): [I | -p^]; pinned element-wise by
        /root/reference/pyslam/residuals/photometric_residual.py:14-35."""
        p = np.atleast_2d(p)
        out = np.zeros([p.shape[0], p.shape[1], cls.dof])
        if p.shape[1] == cls.dim - 1:
            if not directional:
                out[:, 0:3, 0:3] = np.identity(3)
            out[:, 0:3, 3:6] = SO3.wedge(-p)
        elif p.shape[1] == cls.dim:
            out[:, 0:3, 0:3] = p[:, 3, None, None] * np.identity(3)
            out[:, 0:3, 3:6] = SO3.wedge(-p[:, 0:3])
        else:
            raise ValueError('p must have shape (3,), (4,), (N,3) or (N,4)')
        return np.squeeze(out)

    def log(self):
        phi = SO3.log(self.rot)
        rho = SO3.inv_left_jacobian(phi).dot(self.trans)
        return np.hstack([rho, phi])

    def as_matrix(self):
        T = np.identity(4)
        T[0:3, 0:3] = self.rot.as_matrix()
        T[0:3, 3] = self.trans
        return T

    def inv(self):
        inv_rot = self.rot.inv()
        return self.__class__(inv_rot, -(inv_rot.dot(self.trans)))

    def adjoint(self):
        R = self.rot.as_matrix()
        Ad = np.zeros((6, 6))
        Ad[0:3, 0:3] = R
        Ad[0:3, 3:6] = SO3.wedge(self.trans).dot(R)
        Ad[3:6, 3:6] = R
        return Ad

    def dot(self, other):
        if isinstance(other, self.__class__):
            return self.__class__(self.rot.dot(other.rot),
                                  self.rot.dot(other.trans) + self.trans)
        other = np.atleast_2d(other)
        if other.shape[1] == self.dim - 1:
            return np.squeeze(self.rot.dot(other) + self.trans)
        if other.shape[1] == self.dim:
            return np.squeeze(self.as_matrix().dot(other.T).T)
        raise ValueError('vector must have shape (3,), (4,), (N,3) or (N,4)')

    def __mul__(self, other):
        return self.dot(other)

    def perturb(self, xi):
        p = self.__class__.exp(xi).dot(self)
        self.rot = p.rot
        self.trans = p.trans

    def normalize(self):
        self.rot.normalize()

    def __repr__(self):
        return '<{}.{}>\n{}'.format(self.__class__.__module__, self.__class__.__name__, self.as_matrix())

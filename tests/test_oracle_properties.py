"""CPU: properties the reference's own tests never check ("gaps the build must fill"):
finite-difference checks of the oracle's manifold Jacobians, consistency of the loss functions,
and the Lie-group identities the restated liegroups backend must satisfy.  The oracle is what
the GPU parity tests compare against, so its Jacobians are verified here independently of it."""
import copy

import numpy as np
import pytest

from oracle import gn_oracle as O
from oracle import liegroups as OL

RNG = np.random.default_rng(7)


def rand_se3(scale=0.5):
    return OL.SE3.exp(scale * RNG.standard_normal(6))


def rand_se2(scale=0.5):
    return OL.SE2.exp(scale * RNG.standard_normal(3))


def fd_jacobian(f, dof, eps=1e-6):
    """Central differences of f(delta) around delta = 0."""
    f0 = np.atleast_1d(f(np.zeros(dof)))
    J = np.zeros((len(f0), dof))
    for k in range(dof):
        d = np.zeros(dof); d[k] = eps
        J[:, k] = (np.atleast_1d(f(d)) - np.atleast_1d(f(-d))) / (2 * eps)
    return J


def test_reprojection_jacobians_match_finite_differences():
    """reprojection_residual.py:13-37: J_T is the derivative w.r.t. the LEFT perturbation T <- exp(d) T in
    [rho; phi] order, J_p w.r.t. the point (exact, unlike the pose-factor Jacobians)."""
    cam = O.StereoCamera(640., 480., 1000., 1000., 0.25, 1280, 960)
    S = O.invsqrt(np.diag([1., 1., 2.]))
    for _ in range(5):
        T = rand_se3(0.3)
        p_c = np.array([RNG.uniform(-3, 3), RNG.uniform(-2, 2), RNG.uniform(6, 20)])
        p_w = T.inv().dot(p_c)
        obs = cam.project(p_c) + RNG.standard_normal(3)
        res = O.ReprojectionResidual(cam, obs, S)
        r, (J_T, J_p) = res.evaluate([T, p_w], [True, True])

        def f_T(d):
            Tp = copy.deepcopy(T); Tp.perturb(d)
            return res.evaluate([Tp, p_w])

        def f_p(d):
            return res.evaluate([T, p_w + d])

        np.testing.assert_allclose(J_T, fd_jacobian(f_T, 6), rtol=2e-6, atol=1e-6)
        np.testing.assert_allclose(J_p, fd_jacobian(f_p, 3), rtol=2e-6, atol=1e-6)


def test_stereo_camera_jacobians_match_finite_differences():
    cam = O.StereoCamera(150., 100., 250., 200., 1., 640, 480)
    p = np.array([1., 2., 10.])
    uvd, J = cam.project(p, True)
    np.testing.assert_allclose(J, fd_jacobian(lambda d: cam.project(p + d), 3), rtol=1e-7, atol=1e-8)
    xyz, Jt = cam.triangulate(uvd, True)
    np.testing.assert_allclose(xyz, p, rtol=1e-12)
    np.testing.assert_allclose(Jt, fd_jacobian(lambda d: cam.triangulate(uvd + d), 3), rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(Jt @ J, np.eye(3), atol=1e-12)          # inverse maps


@pytest.mark.parametrize('group', ['se2', 'se3'])
def test_pose_factor_jacobians_are_the_reference_approximation(group):
    """pose_residual.py:21-23, pose_to_pose_residual.py:23-28: J = S and J1 = -S Ad(T2 T1^-1), J2 = S drop
    the inverse left Jacobian of the group, so they equal the finite-difference Jacobians only where the residual
    vanishes -- checked at the zero-residual point -- and deviate by O(|r|) away from it."""
    G, rand, dof = (OL.SE2, rand_se2, 3) if group == 'se2' else (OL.SE3, rand_se3, 6)
    S = O.invsqrt(np.diag(np.linspace(0.5, 2., dof)))
    T1, T2 = rand(), rand()
    T21 = T2.dot(T1.inv())
    res = O.PoseToPoseResidual(T21, S)
    r, (J1, J2) = res.evaluate([T1, T2], [True, True])
    np.testing.assert_allclose(r, 0., atol=1e-12)

    def f1(d):
        Tp = copy.deepcopy(T1); Tp.perturb(d)
        return res.evaluate([Tp, T2])

    def f2(d):
        Tp = copy.deepcopy(T2); Tp.perturb(d)
        return res.evaluate([T1, Tp])

    np.testing.assert_allclose(J1, fd_jacobian(f1, dof), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(J2, fd_jacobian(f2, dof), rtol=1e-5, atol=1e-6)
    prior = O.PoseResidual(T1, S)
    r, (J,) = prior.evaluate([T1], [True])

    def f(d):
        Tp = copy.deepcopy(T1); Tp.perturb(d)
        return prior.evaluate([Tp])

    np.testing.assert_allclose(J, fd_jacobian(f, dof), rtol=1e-5, atol=1e-6)
    # away from the zero-residual point the approximation error is first order in the residual
    T2n = copy.deepcopy(T2); T2n.perturb(0.2 * np.ones(dof))
    r, (J1n, J2n) = res.evaluate([T1, T2n], [True, True])

    def f2n(d):
        Tp = copy.deepcopy(T2n); Tp.perturb(d)
        return res.evaluate([T1, Tp])

    err = np.abs(J2n - fd_jacobian(f2n, dof)).max()
    assert 1e-4 < err < 0.5 * np.abs(J2n).max()


@pytest.mark.parametrize('loss', [O.L2Loss(), O.CauchyLoss(2.5), O.HuberLoss(1.5), O.TDistributionLoss(4.)])
def test_loss_weight_is_derivative_of_rho_over_x(loss):
    """losses.py: w(x) = psi(x) / x with psi = rho' for L2, Cauchy, Huber and the t-distribution
    (Tukey's influence is not rho' in the reference, losses.py:141-175, and is exempt)."""
    x = np.concatenate([np.linspace(-6, -0.2, 30), np.linspace(0.2, 6, 30)])
    eps = 1e-6
    drho = (loss.loss(x + eps) - loss.loss(x - eps)) / (2 * eps)
    np.testing.assert_allclose(loss.weight(x) * x, drho, rtol=1e-6, atol=1e-8)


def test_tukey_and_l1_follow_the_reference_formulas():
    t = O.TukeyLoss(3.)
    x = np.array([-4., -2., 0.5, 2.9, 3.1])
    inside = np.abs(x) <= 3.
    np.testing.assert_allclose(t.weight(x), np.where(inside, 1. - (x / 3.) ** 2, 0.))
    np.testing.assert_allclose(t.loss(x), np.where(inside, 1.5 * (1. - (1. - (x / 3.) ** 2) ** 3), 1.5))
    l1 = O.L1Loss()
    np.testing.assert_allclose(l1.loss(x), np.abs(x))
    np.testing.assert_allclose(l1.weight(x), 1. / np.abs(x))


@pytest.mark.parametrize('group', ['se2', 'se3'])
def test_lie_group_identities(group):
    G, rand, dof = (OL.SE2, rand_se2, 3) if group == 'se2' else (OL.SE3, rand_se3, 6)
    for _ in range(5):
        xi = 0.8 * RNG.standard_normal(dof)
        T = G.exp(xi)
        np.testing.assert_allclose(T.log(), xi, rtol=1e-9, atol=1e-10)                       # log(exp(xi)) = xi
        np.testing.assert_allclose(T.dot(T.inv()).as_matrix(), np.eye(T.as_matrix().shape[0]), atol=1e-12)
        A = rand()
        # exp(Ad(A) xi) = A exp(xi) A^-1  (pins the adjoint's block layout and the [rho; phi] ordering)
        lhs = G.exp(A.adjoint() @ xi).as_matrix()
        rhs = A.dot(T).dot(A.inv()).as_matrix()
        np.testing.assert_allclose(lhs, rhs, rtol=1e-9, atol=1e-10)
        # left perturbation: perturb(d) == exp(d) . T
        P = copy.deepcopy(A); d = 0.1 * RNG.standard_normal(dof); P.perturb(d)
        np.testing.assert_allclose(P.as_matrix(), G.exp(d).dot(A).as_matrix(), atol=1e-12)


def test_se3_odot_is_derivative_of_transformed_point():
    """SE3.odot(p) = d(exp(d) p)/dd at d = 0 = [I | -p^] (pinned element-wise by pyslam's fast_se3_odot)."""
    p = RNG.standard_normal(3)
    J = fd_jacobian(lambda d: OL.SE3.exp(d).dot(p), 6)
    np.testing.assert_allclose(OL.SE3.odot(p), J, rtol=1e-7, atol=1e-8)


def test_tiny_angle_branches_are_continuous():
    """liegroups switches to first-order formulas when |phi| <= 1e-8 (np.isclose): exp / log must agree across it."""
    for G, dof in ((OL.SE2, 3), (OL.SE3, 6)):
        base = np.ones(dof)
        base[(2 if dof == 3 else 3):] = 0.
        for ang in (0.5e-8, 2e-8):
            xi = base.copy()
            xi[(2 if dof == 3 else 3):] = ang / np.sqrt(dof - (2 if dof == 3 else 3))
            T = G.exp(xi)
            np.testing.assert_allclose(T.log(), xi, rtol=1e-7, atol=1e-15)

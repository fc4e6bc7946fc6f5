"""CPU: the product's host-side classes (pyslam_b200.lie / losses / sensors /
residuals / utils) against fixtures from the unmodified reference and the
known answers in the reference's tests (tests/test_sensors.py,
tests/test_costs.py, tests/test_utils.py)."""
import numpy as np

from conftest import load_golden
import builders as B
from pyslam_b200 import lie, losses, utils
from pyslam_b200.residuals import PoseResidual, PoseToPoseResidual, QuadraticResidual, ReprojectionResidual
from pyslam_b200.sensors import StereoCamera


def test_losses():
    g = load_golden('losses')
    x = g['x']
    for name, k in [('l2', 0.), ('l1', 0.), ('cauchy', 5.0), ('huber', 1.5), ('tukey', 2.0), ('tdist', 3.0)]:
        L = B.product_loss(name, k)
        np.testing.assert_allclose(L.loss(x), g[name + '_loss'], rtol=1e-14, atol=1e-15)
        np.testing.assert_allclose(L.weight(x.copy()), g[name + '_weight'], rtol=1e-14, atol=1e-15, equal_nan=True)
        assert losses.loss_descriptor(L) == (L.LOSS_KIND, float(k))
    assert losses.loss_descriptor(object()) is None


def test_stereo_camera():
    cam = StereoCamera(150., 100., 250., 200., 1., 300, 200)
    _, J = cam.project([1., 2., 10.], compute_jacobians=True)
    np.testing.assert_allclose(J, [[25., 0., -2.5], [0., 20., -4.], [0., 0., -2.5]])
    _, Jt = cam.triangulate([110., 120., 10.], compute_jacobians=True)
    np.testing.assert_allclose(Jt, [[0.1, 0., 0.4], [0., 0.125, -0.25], [0., 0., -2.5]])
    assert cam.is_valid_measurement([110., 120., 10.]) and not cam.is_valid_measurement([0., 0., -5.])
    assert np.array_equal(cam.is_valid_measurement(np.array([[110., 120., 10.], [-10., 100., 10.], [0., -10., 10.],
                                                             [0., 0., -5.]])), [True, False, False, False])
    uvd, J = cam.project(np.array([[1., 2., 10.], [2., 1., -20.]]), True)
    assert uvd.shape == (2, 3) and J.shape == (2, 3, 3)
    g = load_golden('camera_utils')
    cam = StereoCamera(*g['camera'])
    uvd, J = cam.project(g['pts'], True)
    xyz, Jt = cam.triangulate(g['uvd'], True)
    for a, b in ((uvd, g['uvd']), (J, g['J']), (xyz, g['xyz']), (Jt, g['Jt'])):
        np.testing.assert_allclose(a, b, rtol=1e-15)


def test_utils():
    g = load_golden('camera_utils')
    np.testing.assert_allclose(utils.bilinear_interpolate(g['im'], g['xs'], g['ys']), g['interp'], rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(utils.invsqrt(g['invsqrt_in']), g['invsqrt_out'], rtol=1e-12)
    assert utils.invsqrt(4) == 0.5
    im3 = np.dstack((np.eye(2), np.ones((2, 2)), np.zeros((2, 2))))
    out = utils.bilinear_interpolate(im3, [0.5, 1], [0.5, 0])
    np.testing.assert_allclose(out, [[0.5, 1., 0.], [0., 1., 0.]])


def test_lie_groups_and_residuals():
    g = load_golden('residuals')
    cam = StereoCamera(*load_golden('camera_utils')['camera'])
    for k in range(8):
        r, (JT, Jp) = ReprojectionResidual(cam, g['rp_obs'][k], g['rp_S']).evaluate(
            [B.p_se3(g['rp_T'][k]), g['rp_pts'][k]], [True, True])
        np.testing.assert_allclose(r, g['rp_r'][k], rtol=1e-13)
        np.testing.assert_allclose(JT, g['rp_JT'][k], rtol=1e-13, atol=1e-12)
        np.testing.assert_allclose(Jp, g['rp_Jp'][k], rtol=1e-13, atol=1e-12)
    for name, fr, G in (('se3', B.p_se3, lie.SE3), ('se2', B.p_se2, lie.SE2)):
        S = g[name + '_S']
        for k in range(8):
            T1, T2, To = fr(g[name + '_T1'][k]), fr(g[name + '_T2'][k]), fr(g[name + '_Tobs'][k])
            np.testing.assert_allclose(PoseResidual(To, S).evaluate([T1]), g[name + '_r_pose'][k], rtol=1e-12, atol=1e-13)
            r, (J1, J2) = PoseToPoseResidual(To, S).evaluate([T1, T2], [True, True])
            np.testing.assert_allclose(r, g[name + '_r_p2p'][k], rtol=1e-12, atol=1e-13)
            np.testing.assert_allclose(J1, g[name + '_J1'][k], rtol=1e-12, atol=1e-13)
            np.testing.assert_allclose(J2, g[name + '_J2'][k], rtol=1e-12, atol=1e-13)
            _, jj = PoseToPoseResidual(To, S).evaluate([T1, T2], [False, True])
            assert jj[0] is None and jj[1].shape == S.shape
        for xi, row, lg, ad in zip(g[name + '_xi'], g[name + '_exp'], g[name + '_log'], g[name + '_adj']):
            T = G.exp(xi)
            np.testing.assert_allclose(B.rows_of([T])[0], row, rtol=1e-13, atol=1e-15)
            np.testing.assert_allclose(T.log(), lg, rtol=1e-12, atol=1e-15)
            np.testing.assert_allclose(T.adjoint(), ad, rtol=1e-13, atol=1e-15)
            T2 = fr(row)
            T2.perturb(xi)
            np.testing.assert_allclose(T2.as_matrix(), T.as_matrix() @ T.as_matrix(), rtol=1e-12, atol=1e-14)
        assert lie.group_of(G.identity()) == name
    assert lie.group_of(np.zeros(3)) is None and lie.group_of(lie.SO3.identity()) == 'so3'
    np.testing.assert_allclose(lie.SE3.odot([1., 2., 3.]),
                               [[1, 0, 0, 0, 3, -2], [0, 1, 0, -3, 0, 1], [0, 0, 1, 2, -1, 0]])


def test_quadratic_residual():
    r = QuadraticResidual(2., 3., 1.)
    assert r.evaluate([1., -2., 3.]) == 0. and r.evaluate([0., 3., 1.]) != 0.
    _, j1 = r.evaluate([1., -2., 3.], compute_jacobians=[True, True, True])
    _, j2 = r.evaluate([1., -2., 3.], compute_jacobians=[False, False, False])
    assert np.allclose(j1, [4., 2., 1.]) and not any(j2) and len(j1) == len(j2) == 3


def test_photometric_residual_numpy_evaluate():
    g = load_golden('photometric')
    _, res = B.product_photometric_problem(g, min_grad=float(g['min_grad']))
    assert len(res.im_ref) == int(g['n_ref'])
    r, (J,) = res.evaluate([lie.SE3.exp(g['xi'])], [True])
    np.testing.assert_allclose(r, g['r'], rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(J, g['J'], rtol=1e-11, atol=1e-11)
    T = lie.SE3.exp(g['xi'])
    r2, (Jr, Jt) = res.evaluate([T.rot, T.trans], [True, True])         # (SO3, t) form
    np.testing.assert_allclose(r2, r)
    np.testing.assert_allclose(np.hstack([Jt, Jr]), J)

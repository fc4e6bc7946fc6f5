"""CPU: the round-2 oracle restatements (oracle/pipelines_oracle.py) and the product's HOST classes for SURVEY 8
f2-f4 against the fixtures the unmodified reference produced (oracle/make_golden_r2.py)."""
import os
import tempfile

import numpy as np
import pytest

from conftest import load_golden, rel_err
from oracle import gn_oracle as O
from oracle import liegroups as OL
from oracle import pipelines_oracle as P

import builders as B


def _cams(g, oracle):
    if oracle:
        return {'stereo': O.StereoCamera(*g['stereo_camera']), 'rgbd': P.RGBDCamera(*g['rgbd_camera'])}
    from pyslam_b200.sensors import RGBDCamera, StereoCamera
    return {'stereo': StereoCamera(*[float(v) for v in g['stereo_camera']]), 'rgbd': RGBDCamera(*[float(v) for v in g['rgbd_camera']])}


@pytest.mark.parametrize('name', ['stereo', 'rgbd'])
def test_motion_only_residual_oracle_and_host_class(name):
    g = load_golden('motion_ransac')
    for oracle in (True, False):
        cam = _cams(g, oracle)[name]
        if oracle:
            res = P.ReprojectionMotionOnlyBatchResidual(cam, g[name + '_obs_1'], g[name + '_obs_2'], g['stiffness'])
            T0 = B.o_se3(g[name + '_T0'])
        else:
            from pyslam_b200.residuals import ReprojectionMotionOnlyBatchResidual, ReprojectionMotionOnlyResidual
            res = ReprojectionMotionOnlyBatchResidual(cam, g[name + '_obs_1'], g[name + '_obs_2'], g['stiffness'])
            T0 = B.p_se3(g[name + '_T0'])
            single = ReprojectionMotionOnlyResidual(cam, g[name + '_obs_1'][7], g[name + '_obs_2'][7], g['stiffness'])
            r1, J1 = single.evaluate([T0], [True])
            assert rel_err(r1, g[name + '_r_single']) < 1e-12 and rel_err(J1[0], g[name + '_J_single']) < 1e-12
        r, J = res.evaluate([T0], [True])
        assert rel_err(r, g[name + '_r']) < 1e-12
        assert rel_err(J[0], g[name + '_J']) < 1e-12
        assert rel_err(res.evaluate([T0]), g[name + '_r']) < 1e-12


@pytest.mark.parametrize('name', ['stereo', 'rgbd'])
def test_ransac_oracle(name):
    g = load_golden('motion_ransac')
    cam = _cams(g, True)[name]
    p1 = np.atleast_2d(cam.triangulate(g[name + '_obs_1']))
    p2 = np.atleast_2d(cam.triangulate(g[name + '_obs_2']))
    idx = g[name + '_ransac_idx']
    T = P.compute_transforms(p1[idx], p2[idx])
    # minimal sets with a repeated index give a rank-1 W: the SVD basis (and the transform) is arbitrary there
    proper = np.array([len(set(row)) == 3 for row in idx])
    assert proper.sum() > 380
    assert np.abs(T - g[name + '_ransac_T'])[proper].max() < 1e-9
    masks = P.ransac_masks(g[name + '_ransac_T'], p1, g[name + '_obs_2'], cam, 5)
    assert np.array_equal(masks.sum(axis=1), g[name + '_ransac_counts'])
    best = int(np.argmax(masks.sum(axis=1)))
    assert np.array_equal(np.where(masks[best])[0], g[name + '_ransac_best_inliers'])


def test_orientation_residual_oracle_and_host_class():
    g = load_golden('orientation')
    from pyslam_b200 import lie as PL
    from pyslam_b200.residuals import PoseToPoseOrientationResidual
    for mk_se3, mk_so3, cls in ((B.o_se3, lambda m: OL.SO3(m.reshape(3, 3)), P.PoseToPoseOrientationResidual),
                                (B.p_se3, lambda m: PL.SO3(m.reshape(3, 3)), PoseToPoseOrientationResidual)):
        res = cls(mk_so3(g['C_obs'][2]), g['S3'])
        r, J = res.evaluate([mk_se3(g['T_init'][2]), mk_se3(g['T_init'][3])], [True, True])
        assert rel_err(r, g['single_r']) < 1e-12
        assert rel_err(J[0], g['single_J1']) < 1e-12 and rel_err(J[1], g['single_J2']) < 1e-12


def test_rgbd_camera_oracle_and_host_class():
    g = load_golden('rgbd_camera')
    from pyslam_b200.sensors import RGBDCamera
    for cam in (P.RGBDCamera(*g['params']), RGBDCamera(*g['params'])):
        uvz, Jp = cam.project(g['pts'], True)
        tri, Jt = cam.triangulate(g['uvz'], True)
        assert np.array_equal(uvz, g['uvz']) or rel_err(uvz, g['uvz']) < 1e-15
        assert rel_err(Jp, g['project_jac']) < 1e-15 and rel_err(tri, g['tri']) < 1e-15 and rel_err(Jt, g['tri_jac']) < 1e-15
        assert np.array_equal(np.asarray(cam.is_valid_measurement(g['uvz'])), g['valid'])
    assert RGBDCamera(*g['params']).intrinsics()[4] == 0.


def test_pyramid_oracle_against_reference_keyframe():
    """The restated cv2.pyrDown / Sobel / sub-sampling against what the reference's DenseStereoKeyframe produced."""
    g = load_golden('dense_pipeline')
    levels = int(g['levels'])
    ims, jac = P.image_pyramid(g['left0'], levels)
    for l in range(levels):
        assert np.array_equal(ims[l], g['im_pyr_%d' % l]), l
        assert np.abs(jac[l] - g['jac_%d' % l]).max() < 1e-12, l
    disp = P.subsample_pyramid(g['disp_0'], levels, 0.5)
    for l in range(levels):
        assert np.array_equal(disp[l], g['disp_%d' % l]), l


@pytest.mark.parametrize('conv', ['Twv', 'Tvw'])
def test_trajectory_metrics_host_class(conv):
    g = load_golden('metrics')
    from pyslam_b200 import lie as PL
    from pyslam_b200.metrics import TrajectoryMetrics
    gt = [PL.SE3.from_matrix(m) for m in g['gt']]
    est = [PL.SE3.from_matrix(m) for m in g['est']]
    tm = TrajectoryMetrics(gt, est, convention=conv)
    errs, avg = tm.segment_errors([5., 10., 20.])
    np.testing.assert_allclose(errs, g[conv + '_seg_errs'], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(avg, g[conv + '_seg_avg'], rtol=1e-9, atol=1e-12)
    t, r = tm.traj_errors()
    np.testing.assert_allclose(t, g[conv + '_traj_t'], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(r, g[conv + '_traj_r'], rtol=1e-9, atol=1e-12)
    t, r = tm.rel_errors(delta=2)
    np.testing.assert_allclose(t, g[conv + '_rel_t'], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(r, g[conv + '_rel_r'], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(tm.endpoint_error(range(5, 40), 'cm', 'deg'), g[conv + '_endpoint'], rtol=1e-9)
    np.testing.assert_allclose(tm.rms_err(), g[conv + '_rms_traj'], rtol=1e-9)
    np.testing.assert_allclose(tm.rms_err(error_type='rel', delta=3), g[conv + '_rms_rel'], rtol=1e-9)
    np.testing.assert_allclose(tm.mean_err(), g[conv + '_mean'], rtol=1e-9)
    np.testing.assert_allclose(tm.cum_err()[0], g[conv + '_cum_t'], rtol=1e-9)
    np.testing.assert_allclose(tm.cum_dists, g[conv + '_cum_dists'], rtol=1e-12)
    with tempfile.TemporaryDirectory() as d:                      # the .mat format: poses M x M x N (metrics.py:95-110)
        f = os.path.join(d, 'traj.mat')
        tm.savemat(f, extras={'note': np.array([1.0])})
        import scipy.io
        raw = scipy.io.loadmat(f)
        assert raw['poses_gt'].shape == (4, 4, len(gt)) and raw['poses_est'].shape == (4, 4, len(gt))
        tm2 = TrajectoryMetrics.loadmat(f)
        np.testing.assert_allclose(tm2.rms_err(), tm.rms_err(), rtol=1e-9)

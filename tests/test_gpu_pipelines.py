"""GPU: the rows SURVEY 8 marks "next" -- f2 (motion-only reprojection, RANSAC), f3 (pyramids + the dense
coarse-to-fine loop), f4 (RGB-D camera, SO(3)-only factor) and the (SO3, t) parameter form of the photometric
residual -- through Problem -> ctypes -> C ABI, against fixtures the UNMODIFIED reference produced
(oracle/make_golden_r2.py) and against the CPU oracle."""
import numpy as np
import pytest

from conftest import load_golden, rel_err
import builders as B

pytestmark = pytest.mark.gpu


def _cam(g, name):
    from pyslam_b200.sensors import RGBDCamera, StereoCamera
    if name == 'stereo':
        return StereoCamera(*[float(v) for v in g['stereo_camera']])
    return RGBDCamera(*[float(v) for v in g['rgbd_camera']])


# ------------------------------------------------------------------ f2
@pytest.mark.parametrize('name', ['stereo', 'rgbd'])
@pytest.mark.parametrize('lname', ['l2', 'huber'])
def test_motion_only_batch_against_reference(name, lname):
    import pyslam_b200
    from pyslam_b200 import lie as PL
    from pyslam_b200.residuals import ReprojectionMotionOnlyBatchResidual
    g = load_golden('motion_ransac')
    res = ReprojectionMotionOnlyBatchResidual(_cam(g, name), g[name + '_obs_1'], g[name + '_obs_2'], g['stiffness'])
    pr = pyslam_b200.Problem(B.nondecreasing_options(pyslam_b200.Options))
    pr.add_residual_block(res, ['T_2_1'], B.product_loss(lname, 1.5))
    pr.initialize_params({'T_2_1': PL.SE3.identity()})
    low = pr._ensure_lowered()
    assert low.kinds == [('motion',)]
    cost_lin, _, _ = pr._engine.iterate(0., True)
    dx = pr._engine.get_update(low.dim)[low.ref_from_internal]
    hist = g['%s_%s_cost_history' % (name, lname)]
    assert abs(cost_lin - hist[0]) < 1e-10 * hist[0]
    assert rel_err(dx, g['%s_%s_dx' % (name, lname)][0]) < 1e-7
    pr2 = pyslam_b200.Problem(B.nondecreasing_options(pyslam_b200.Options))
    pr2.add_residual_block(res, ['T_2_1'], B.product_loss(lname, 1.5))
    pr2.initialize_params({'T_2_1': PL.SE3.identity()})
    pr2.solve()
    assert len(pr2._cost_history) == len(hist)
    np.testing.assert_allclose(pr2._cost_history, hist, rtol=1e-7)
    assert rel_err(B.rows_of([pr2.param_dict['T_2_1']])[0], g['%s_%s_T_final' % (name, lname)]) < 1e-7


def test_motion_only_single_blocks_are_batched():
    """N single-point ReprojectionMotionOnlyResidual blocks on one pose = one device batch, same answer as the batch class."""
    import pyslam_b200
    from pyslam_b200 import lie as PL
    from pyslam_b200.residuals import ReprojectionMotionOnlyResidual
    g = load_golden('motion_ransac')
    cam = _cam(g, 'stereo')
    pr = pyslam_b200.Problem(B.nondecreasing_options(pyslam_b200.Options))
    for o1, o2 in zip(g['stereo_obs_1'], g['stereo_obs_2']):
        pr.add_residual_block(ReprojectionMotionOnlyResidual(cam, o1, o2, g['stiffness']), ['T_2_1'], B.product_loss('huber', 1.5))
    pr.initialize_params({'T_2_1': PL.SE3.identity()})
    pr.solve()
    assert set(pr._low.kinds) == {('motion',)}
    np.testing.assert_allclose(pr._cost_history, g['stereo_huber_cost_history'], rtol=1e-7)


@pytest.mark.parametrize('name', ['stereo', 'rgbd'])
def test_ransac_against_reference(name):
    from pyslam_b200 import engine as E
    from pyslam_b200.pipelines import FrameToFrameRANSAC
    g = load_golden('motion_ransac')
    cam = _cam(g, name)
    rs = FrameToFrameRANSAC(cam)
    rs.set_obs(g[name + '_obs_1'], g[name + '_obs_2'])
    idx = g[name + '_ransac_idx']
    T, counts, best, mask = E.ransac(rs.pts_1, rs.obs_2, cam.intrinsics(), rs.ransac_thresh, sample_idx=idx, pts_2=rs.pts_2)
    proper = np.array([len(set(r)) == 3 for r in idx])      # repeated indices: rank-1 problem, arbitrary SVD basis
    assert np.abs(T - g[name + '_ransac_T'])[proper].max() < 1e-8
    assert np.array_equal(counts[proper], g[name + '_ransac_counts'][proper])
    # given hypotheses (compute_ransac_cost): every count, and the masks of a few
    T2, counts2, best2, mask2 = E.ransac(rs.pts_1, rs.obs_2, cam.intrinsics(), rs.ransac_thresh, T_21=g[name + '_ransac_T'])
    assert np.array_equal(counts2, g[name + '_ransac_counts'])
    assert best2 == int(np.argmax(g[name + '_ransac_counts']))
    assert np.array_equal(np.where(mask2)[0], g[name + '_ransac_best_inliers'])
    masks = rs.compute_ransac_cost(g[name + '_ransac_T'][:5], rs.pts_1, rs.obs_2, cam, rs.ransac_thresh)
    assert np.array_equal(masks.sum(axis=1), g[name + '_ransac_counts'][:5])
    # the whole routine with the reference's generator state
    np.random.seed(1234)
    T_best, o1, o2, inl = rs.perform_ransac()
    assert np.array_equal(inl, g[name + '_ransac_best_inliers'])
    assert np.abs(T_best.as_matrix() - g[name + '_ransac_best_T']).max() < 1e-8
    assert np.array_equal(o2, g[name + '_obs_2'][inl])


# ------------------------------------------------------------------ f4
@pytest.mark.parametrize('lname', ['l2', 'cauchy'])
def test_orientation_factor_against_reference(lname):
    import pyslam_b200
    from pyslam_b200 import lie as PL
    from pyslam_b200.residuals import PoseResidual, PoseToPoseOrientationResidual, PoseToPoseResidual
    from pyslam_b200.utils import invsqrt
    g = load_golden('orientation')
    n = len(g['T_init'])
    keys = ['T_%d_0' % k for k in range(n)]
    loss = B.product_loss(lname, 1.0)
    pr = pyslam_b200.Problem(B.nondecreasing_options(pyslam_b200.Options))
    pr.add_residual_block(PoseResidual(B.p_se3(g['T_true'][0]), np.real(invsqrt(1e-6 * np.eye(6)))), keys[0])
    for k in range(n - 1):
        pr.add_residual_block(PoseToPoseResidual(B.p_se3(g['T_obs'][k]), g['S6']), [keys[k], keys[k + 1]], loss)
        pr.add_residual_block(PoseToPoseOrientationResidual(PL.SO3(g['C_obs'][k].reshape(3, 3)), g['S3']), [keys[k], keys[k + 1]], loss)
    pr.initialize_params({k: B.p_se3(r) for k, r in zip(keys, g['T_init'])})
    low = pr._ensure_lowered()
    assert ('orient',) in low.kinds and ('dense',) not in low.kinds
    pr._engine.iterate(0., True)
    dx = pr._engine.get_update(low.dim)[low.ref_from_internal]
    assert rel_err(dx, g[lname + '_dx0']) < 1e-7
    pr._upload_params(pr.param_dict)
    pr.solve()
    hist = g[lname + '_cost_history']
    assert len(pr._cost_history) == len(hist)
    np.testing.assert_allclose(pr._cost_history, hist, rtol=1e-6)
    assert rel_err(B.rows_of([pr.param_dict[k] for k in keys]), g[lname + '_T_final']) < 1e-7


def test_rgbd_bundle_adjustment_against_oracle():
    """ReprojectionResidual with an RGBDCamera (third measurement = depth) through the fused BA kernels."""
    from oracle import gn_oracle as O
    from oracle import liegroups as OL
    from oracle import pipelines_oracle as P
    import pyslam_b200
    from pyslam_b200 import lie as PL, synthetic
    from pyslam_b200.residuals import ReprojectionResidual
    from pyslam_b200.sensors import RGBDCamera
    d = synthetic.stereo_ba(6, 120, track=4, seed=11)
    params = (640., 480., 1000., 1000., 1280, 960)
    ocam, pcam = P.RGBDCamera(*params), RGBDCamera(*params)
    rng = np.random.default_rng(2)
    pc = np.einsum('nij,nj->ni', d['R_true'][d['pose_idx']], d['pts_true'][d['pt_idx']]) + d['t_true'][d['pose_idx']]
    obs = ocam.project(pc) + 0.05 * rng.standard_normal(pc.shape)
    pk, qk = B.ba_keys(d)
    op = O.OracleProblem(B.nondecreasing_options(O.Options))
    pp = pyslam_b200.Problem(B.nondecreasing_options(pyslam_b200.Options))
    for ci, qi, o in zip(d['pose_idx'], d['pt_idx'], obs):
        op.add_residual_block(O.ReprojectionResidual(ocam, o, d['stiffness']), [pk[ci], qk[qi]], O.HuberLoss(1.0))
        pp.add_residual_block(ReprojectionResidual(pcam, o, d['stiffness']), [pk[ci], qk[qi]], B.product_loss('huber', 1.0))
    po = {k: OL.SE3(OL.SO3(R), t) for k, R, t in zip(pk, d['R0'], d['t0'])}
    po.update({k: np.array(p) for k, p in zip(qk, d['pts0'])})
    ppar = {k: PL.SE3(PL.SO3(R), t) for k, R, t in zip(pk, d['R0'], d['t0'])}
    ppar.update({k: np.array(p) for k, p in zip(qk, d['pts0'])})
    op.initialize_params(po); pp.initialize_params(ppar)
    op.set_parameters_constant([pk[0]]); pp.set_parameters_constant([pk[0]])
    op.solve(); pp.solve()
    assert set(pp._low.kinds) == {('reproj',)}
    assert len(pp._cost_history) == len(op._cost_history)
    np.testing.assert_allclose(pp._cost_history, op._cost_history, rtol=1e-6)


# ------------------------------------------------------------------ (SO3, t) photometric form, f3
def _pipeline_inputs(g):
    from pyslam_b200.sensors import StereoCamera
    c = g['camera']
    return StereoCamera(float(c[0]), float(c[1]), float(c[2]), float(c[3]), float(c[4]), int(c[5]), int(c[6]))


def test_image_and_disparity_pyramids_against_reference():
    from pyslam_b200 import engine as E
    g = load_golden('dense_pipeline')
    levels = int(g['levels'])
    ims, jac = E.image_pyramid(g['left0'], levels)
    disp = E.subsample_pyramid(g['disp_0'], levels, 0.5)
    for l in range(levels):
        assert np.array_equal(ims[l], g['im_pyr_%d' % l]), l                 # 8-bit pyrDown chain: bit-exact
        assert np.abs(jac[l] - g['jac_%d' % l]).max() < 1e-12, l
        assert np.array_equal(disp[l], g['disp_%d' % l]), l
    odd = (np.arange(37 * 53).reshape(37, 53) * 7 % 251).astype(np.uint8)      # odd sizes: ceil(w / 2) per level
    from oracle import pipelines_oracle as P
    ims, jac = E.image_pyramid(odd, 3)
    ims_o, jac_o = P.image_pyramid(odd, 3)
    for a, b, c, d_ in zip(ims, ims_o, jac, jac_o):
        assert a.shape == b.shape and np.array_equal(a, b) and np.abs(c - d_).max() < 1e-12


@pytest.mark.parametrize('const_t', [False, True])
def test_photometric_split_form_against_reference(const_t):
    """params ['R_1_0', 't_1_0_1'] (the form pipelines/dense.py uses), translation variable or constant."""
    import pyslam_b200
    from pyslam_b200 import configs, lie as PL
    from pyslam_b200.losses import HuberLoss
    from pyslam_b200.pipelines import DenseStereoKeyframe, DenseStereoPipeline
    from pyslam_b200.residuals import PhotometricResidualSE3
    g = load_golden('dense_pipeline')
    cam = _pipeline_inputs(g)
    levels, lvl = int(g['levels']), int(g['split_level'])
    pipe = DenseStereoPipeline(cam)
    pcam = pipe.pyr_cameras[pipe.pyrlevel_sequence.index(lvl)]
    tf = DenseStereoKeyframe(g['left1'], g['right1'], levels)
    res = PhotometricResidualSE3(pcam, g['im_pyr_%d' % lvl], g['disp_%d' % lvl], tf.im_pyr[lvl], g['jac_%d' % lvl],
                                 pipe.intensity_stiffness, pipe.depth_stiffness / 2. ** -lvl, pipe.min_grad)
    tag = 'split_constt' if const_t else 'split'

    def make():
        pr = pyslam_b200.Problem(configs.dense_options())
        pr.add_residual_block(res, ['R_1_0', 't_1_0_1'], loss=HuberLoss(float(g['loss_k'])))
        pr.initialize_params({'R_1_0': PL.SO3.identity(), 't_1_0_1': np.zeros(3)})
        if const_t:
            pr.set_parameters_constant('t_1_0_1')
        return pr
    pr = make()
    low = pr._ensure_lowered()
    assert low.kinds == [('photo', 'split')]
    pr._engine.iterate(0., True)
    dx = pr._engine.get_update(low.dim)[low.ref_from_internal]
    assert rel_err(dx, g[tag + '_dx0']) < 1e-7
    pr = make()
    pr.solve()
    hist = g[tag + '_history']
    assert len(pr._cost_history) == len(hist)
    np.testing.assert_allclose(pr._cost_history, hist, rtol=1e-6)
    assert np.abs(pr.param_dict['R_1_0'].mat - g[tag + '_R_final']).max() < 1e-7
    assert np.abs(np.asarray(pr.param_dict['t_1_0_1']) - g[tag + '_t_final']).max() < 1e-7


def test_dense_stereo_pipeline_against_reference():
    """DenseStereoPipeline.track twice: keyframe pyramids on the device, 4-level coarse-to-fine loop on one engine handle."""
    from pyslam_b200.pipelines import DenseStereoPipeline
    g = load_golden('dense_pipeline')
    pipe = DenseStereoPipeline(_pipeline_inputs(g))
    pipe.track(g['left0'], g['right0'])
    kf = pipe.keyframes[0]
    for l in range(int(g['levels'])):
        assert np.array_equal(kf.im_pyr[l], g['im_pyr_%d' % l])
        assert np.array_equal(kf.disparity[l], g['disp_%d' % l])          # cv2.StereoBM on the CPU, sub-sampled on the device
    pipe.track(g['left1'], g['right1'])
    assert len(pipe.level_summaries) == int(g['n_solves'])
    for k, (lvl, n_it, c0, c1) in enumerate(pipe.level_summaries):
        h = g['history_%d' % k]
        assert n_it == len(h) - 1, (k, n_it, len(h))
        assert abs(c0 - h[0]) < 1e-6 * h[0] and abs(c1 - h[-1]) < 1e-6 * h[-1]
    assert rel_err(B.rows_of([pipe.T_c_w[-1]])[0], g['T_final']) < 1e-6


def test_dense_rgbd_pipeline_against_reference():
    """DenseRGBDPipeline.track twice (RGBDCamera, depth maps): depth pyramid on the device, photometric kernel in its RGB-D
    form ((SO3, t) parameters, translation constant on the coarsest level), against the reference's own run."""
    from pyslam_b200.pipelines import DenseRGBDPipeline
    from pyslam_b200.sensors import RGBDCamera
    g = load_golden('dense_rgbd_pipeline')
    c = g['camera']
    pipe = DenseRGBDPipeline(RGBDCamera(float(c[0]), float(c[1]), float(c[2]), float(c[3]), int(c[4]), int(c[5])))
    pipe.track(g['im0'], g['depth0'])
    kf = pipe.keyframes[0]
    for l in range(int(g['levels'])):
        assert np.array_equal(kf.depth[l], g['depth_%d' % l])
    pipe.track(g['im1'], g['depth1'])
    assert len(pipe.level_summaries) == int(g['n_solves'])
    for k, (lvl, n_it, c0, c1) in enumerate(pipe.level_summaries):
        h = g['history_%d' % k]
        assert n_it == len(h) - 1, (k, n_it, len(h))
        assert abs(c0 - h[0]) < 1e-6 * h[0] and abs(c1 - h[-1]) < 1e-6 * h[-1]
    assert rel_err(B.rows_of([pipe.T_c_w[-1]])[0], g['T_final']) < 1e-6


@pytest.mark.parametrize('name', ['stereo', 'rgbd'])
def test_sparse_pipeline_motion_estimate_against_reference(name):
    """The hot part of Sparse*Pipeline._compute_frame_to_frame_motion (pipelines/sparse.py:143-163): RANSAC guess + inliers ->
    motion-only refinement, with matches handed in through the matcher plug-in (the reference's matcher is libviso2)."""
    from pyslam_b200.pipelines import SparseRGBDPipeline, SparseStereoPipeline
    g = load_golden('motion_ransac')
    cam = _cam(g, name)

    class FixedMatches:
        def match(self, ref_frame, track_frame):
            return g[name + '_obs_1'], g[name + '_obs_2']
    pipe = (SparseStereoPipeline if name == 'stereo' else SparseRGBDPipeline)(cam, matcher=FixedMatches())
    img = np.zeros((4, 4), np.uint8)
    if name == 'stereo':
        pipe.track(img, img)
    else:
        pipe.track(img, np.ones((4, 4)))
    np.random.seed(1234)                      # the generator state the reference's perform_ransac saw
    if name == 'stereo':
        pipe.track(img, img)
    else:
        pipe.track(img, np.ones((4, 4)))
    hist = g[name + '_f2f_history']
    assert len(pipe.last_cost_history) == len(hist)
    np.testing.assert_allclose(pipe.last_cost_history, hist, rtol=1e-7)
    T_ref = B.p_se3(g[name + '_f2f_T'])
    T_ref.normalize()
    assert rel_err(B.rows_of([pipe.T_c_w[-1]])[0], B.rows_of([T_ref])[0]) < 1e-7
    bare = type(pipe)(cam)                      # without a matcher (and without libviso2) the pipeline says so
    if bare.matcher is None:
        with pytest.raises(RuntimeError):
            bare._compute_frame_to_frame_motion(None, None)

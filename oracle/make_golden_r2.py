"""TEST INFRASTRUCTURE -- round-2 fixtures from the UNMODIFIED reference ((SO3, t) form).

    python oracle/make_golden_r2.py        (build container only: needs /root/reference)

Same loader and shims as oracle/make_golden.py, plus an empty `viso2` stub module (pyslam/pipelines/__init__.py
imports every sub-module, and sparse.py imports the un-installable viso2; nothing of it is called).  Writes
tests/golden/{motion_ransac,orientation,rgbd_camera,dense_pipeline,dense_rgbd_pipeline,metrics}.npz: inputs + the reference's outputs.
"""
import os
import sys
import types
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.reference_loader import load_reference  # noqa: E402

ref = load_reference()
warnings.simplefilter('ignore')
sys.modules.setdefault('viso2', types.ModuleType('viso2'))
from liegroups import SE3, SO3  # noqa: E402  (oracle/liegroups)
from pyslam.problem import Options, Problem  # noqa: E402
from pyslam.residuals import (PoseResidual, PoseToPoseResidual, PoseToPoseOrientationResidual,  # noqa: E402
                              ReprojectionMotionOnlyBatchResidual, ReprojectionMotionOnlyResidual)
from pyslam.sensors import RGBDCamera, StereoCamera  # noqa: E402
from pyslam.utils import invsqrt  # noqa: E402
import pyslam.losses as ref_losses  # noqa: E402
import pyslam.pipelines as ref_pipe  # noqa: E402
from pyslam.metrics import TrajectoryMetrics  # noqa: E402

from pyslam_b200 import synthetic  # noqa: E402  (input generators only)

OUT = os.path.join(ROOT, 'tests', 'golden')


def row(T):
    return np.concatenate([np.asarray(T.rot.mat).ravel(), np.asarray(T.trans)])


def nondecreasing():
    o = Options()
    o.allow_nondecreasing_steps = True
    o.max_nondecreasing_steps = 3
    return o


def traced_solve(problem):
    dxs = []
    orig = problem.solve_one_iter

    def wrapped():
        dx, cost = orig()
        dxs.append(dx.copy())
        return dx, cost
    problem.solve_one_iter = wrapped
    problem.solve()
    return dxs


# ------------------------------------------------------------------ f2: motion-only + RANSAC
def motion_ransac():
    rng = np.random.default_rng(21)
    out = {}
    T_true = SE3.exp(np.array([0.3, -0.1, 0.2, 0.02, -0.03, 0.05]))
    n = 300
    p1 = np.stack([rng.uniform(-4, 4, n), rng.uniform(-3, 3, n), rng.uniform(6, 20, n)], axis=1)
    p2 = T_true.dot(p1)
    S = np.real(invsqrt(np.diag([1., 1., 2.])))
    for name, cam in (('stereo', StereoCamera(*synthetic.BA_CAMERA)), ('rgbd', RGBDCamera(640., 480., 1000., 1000., 1280, 960))):
        obs_1 = cam.project(p1) + 0.2 * rng.standard_normal((n, 3))
        obs_2 = cam.project(p2) + 0.2 * rng.standard_normal((n, 3))
        bad = rng.choice(n, 30, replace=False)
        obs_2[bad, :2] += rng.uniform(-80, 80, (30, 2))
        out[name + '_obs_1'], out[name + '_obs_2'] = obs_1, obs_2
        res = ReprojectionMotionOnlyBatchResidual(cam, obs_1, obs_2, S)
        T0 = SE3.exp(np.array([0.1, 0.05, -0.1, 0.01, 0.02, -0.01]))
        r, J = res.evaluate([T0], [True])
        out[name + '_T0'], out[name + '_r'], out[name + '_J'] = row(T0), r, J[0]
        single = ReprojectionMotionOnlyResidual(cam, obs_1[7], obs_2[7], S)
        r1, J1 = single.evaluate([T0], [True])
        out[name + '_r_single'], out[name + '_J_single'] = r1, J1[0]
        for lname, loss in (('l2', ref_losses.L2Loss()), ('huber', ref_losses.HuberLoss(1.5))):
            pr = Problem(nondecreasing())
            pr.add_residual_block(res, ['T_2_1'], loss)
            pr.initialize_params({'T_2_1': SE3.identity()})
            dxs = traced_solve(pr)
            out['%s_%s_cost_history' % (name, lname)] = np.array(pr._cost_history)
            out['%s_%s_dx' % (name, lname)] = np.array(dxs)
            out['%s_%s_T_final' % (name, lname)] = row(pr.param_dict['T_2_1'])
        # RANSAC (ransac.py:107-165): same generator state as perform_ransac sees
        rs = ref_pipe.FrameToFrameRANSAC(cam)
        rs.set_obs(obs_1, obs_2)
        np.random.seed(1234)
        rand_idx = np.random.randint(rs.num_pts, size=(rs.ransac_iters, rs.num_min_set_pts))
        T_stack = ref_pipe.compute_transform_fast(rs.pts_1[rand_idx], rs.pts_2[rand_idx], np.empty(4))
        masks = rs.compute_ransac_cost(T_stack, rs.pts_1, rs.obs_2, cam, rs.ransac_thresh)
        np.random.seed(1234)
        T_best, o1, o2, idx_best = rs.perform_ransac()
        out[name + '_ransac_idx'] = rand_idx
        out[name + '_ransac_T'] = T_stack
        out[name + '_ransac_counts'] = masks.sum(axis=1)
        out[name + '_ransac_best_T'] = T_best.as_matrix()
        out[name + '_ransac_best_inliers'] = idx_best
        # the tail of Sparse*Pipeline._compute_frame_to_frame_motion (pipelines/sparse.py:150-163,203-216) with the reference's
        # classes: RANSAC guess + inliers -> motion-only batch residual -> Problem(motion_options).solve()
        mo = Options()
        mo.allow_nondecreasing_steps = True
        mo.max_nondecreasing_steps = 5
        mo.min_cost_decrease = 0.99
        mo.max_iters = 30
        mo.num_threads = 1
        mo.linesearch_max_iters = 0
        f2f = Problem(mo)
        f2f.add_residual_block(ReprojectionMotionOnlyBatchResidual(cam, o1, o2, np.diag([1., 1., 1.])), ['T_1_0'], loss=ref_losses.L2Loss())
        f2f.initialize_params({'T_1_0': T_best})
        f2f.solve()
        out[name + '_f2f_history'] = np.array(f2f._cost_history)
        out[name + '_f2f_T'] = row(f2f.param_dict['T_1_0'])
    out['stiffness'] = S
    out['stereo_camera'] = np.array(synthetic.BA_CAMERA)
    out['rgbd_camera'] = np.array([640., 480., 1000., 1000., 1280, 960])
    np.savez_compressed(os.path.join(OUT, 'motion_ransac.npz'), **out)
    print('motion_ransac: stereo huber history', out['stereo_huber_cost_history'], 'ransac best count', masks.sum(axis=1).max())


# ------------------------------------------------------------------ f4: orientation factor, RGB-D camera, metrics
def orientation():
    rng = np.random.default_rng(5)
    out = {}
    n = 8
    T_true = [SE3.identity()]
    step = SE3.exp(np.array([0.5, 0.02, 0.05, 0.03, -0.02, 0.2]))
    for _ in range(n - 1):
        T_true.append(step.dot(T_true[-1]))
    T_init = [T_true[0]] + [SE3.exp(0.05 * rng.standard_normal(6)).dot(T) for T in T_true[1:]]
    C_obs = [SO3.exp(0.01 * rng.standard_normal(3)).dot(T_true[k + 1].dot(T_true[k].inv()).rot) for k in range(n - 1)]
    T_obs = [SE3.exp(0.01 * rng.standard_normal(6)).dot(T_true[k + 1].dot(T_true[k].inv())) for k in range(n - 1)]
    S3 = np.real(invsqrt(1e-3 * np.eye(3)))
    S6 = np.real(invsqrt(1e-2 * np.eye(6)))
    res = PoseToPoseOrientationResidual(C_obs[2], S3)
    r, J = res.evaluate([T_init[2], T_init[3]], [True, True])
    out['single_r'], out['single_J1'], out['single_J2'] = r, J[0], J[1]
    for lname, loss in (('l2', ref_losses.L2Loss()), ('cauchy', ref_losses.CauchyLoss(1.0))):
        pr = Problem(nondecreasing())
        keys = ['T_%d_0' % k for k in range(n)]
        pr.add_residual_block(PoseResidual(T_true[0], np.real(invsqrt(1e-6 * np.eye(6)))), keys[0])
        for k in range(n - 1):
            pr.add_residual_block(PoseToPoseResidual(T_obs[k], S6), [keys[k], keys[k + 1]], loss)
            pr.add_residual_block(PoseToPoseOrientationResidual(C_obs[k], S3), [keys[k], keys[k + 1]], loss)
        pr.initialize_params({k: SE3(SO3(T.rot.mat.copy()), T.trans.copy()) for k, T in zip(keys, T_init)})
        dxs = traced_solve(pr)
        out[lname + '_cost_history'] = np.array(pr._cost_history)
        out[lname + '_dx0'] = dxs[0]
        out[lname + '_T_final'] = np.array([row(pr.param_dict[k]) for k in keys])
    out['T_init'] = np.array([row(T) for T in T_init])
    out['T_true'] = np.array([row(T) for T in T_true])
    out['T_obs'] = np.array([row(T) for T in T_obs])
    out['C_obs'] = np.array([C.mat.ravel() for C in C_obs])
    out['S3'], out['S6'] = S3, S6
    np.savez_compressed(os.path.join(OUT, 'orientation.npz'), **out)
    print('orientation: l2 history', out['l2_cost_history'])


def rgbd_camera():
    rng = np.random.default_rng(9)
    cam = RGBDCamera(319.5, 239.5, 525., 520., 640, 480)
    pts = np.stack([rng.uniform(-2, 2, 50), rng.uniform(-1.5, 1.5, 50), rng.uniform(0.5, 8, 50)], axis=1)
    pts[3, 2] = -1.0
    uvz, Jp = cam.project(pts, True)
    back, Jt = cam.triangulate(uvz, True)
    np.savez_compressed(os.path.join(OUT, 'rgbd_camera.npz'), params=np.array([319.5, 239.5, 525., 520., 640, 480]), pts=pts, uvz=uvz,
                        project_jac=Jp, tri=back, tri_jac=Jt, valid=np.asarray(cam.is_valid_measurement(uvz)))


def metrics():
    rng = np.random.default_rng(13)
    n = 60
    gt = [SE3.identity()]
    for _ in range(n - 1):
        gt.append(gt[-1].dot(SE3.exp(np.array([0.8, 0.02, 0.0, 0.0, 0.01, 0.05]) + 0.01 * rng.standard_normal(6))))
    est = [T.dot(SE3.exp(0.02 * rng.standard_normal(6))) for T in gt]
    out = {'gt': np.array([T.as_matrix() for T in gt]), 'est': np.array([T.as_matrix() for T in est])}
    for conv in ('Twv', 'Tvw'):
        tm = TrajectoryMetrics(gt, est, convention=conv)
        errs, avg = tm.segment_errors([5., 10., 20.])
        t, r = tm.traj_errors()
        tr, rr = tm.rel_errors(delta=2)
        out.update({conv + '_seg_errs': errs, conv + '_seg_avg': avg, conv + '_traj_t': t, conv + '_traj_r': r,
                    conv + '_rel_t': tr, conv + '_rel_r': rr, conv + '_endpoint': np.array(tm.endpoint_error(range(5, 40), 'cm', 'deg')),
                    conv + '_rms_traj': np.array(tm.rms_err()), conv + '_rms_rel': np.array(tm.rms_err(error_type='rel', delta=3)),
                    conv + '_mean': np.array(tm.mean_err()), conv + '_cum_t': tm.cum_err()[0], conv + '_cum_dists': tm.cum_dists})
    np.savez_compressed(os.path.join(OUT, 'metrics.npz'), **out)


# ------------------------------------------------------------------ f3 + (SO3, t) form: the dense stereo pipeline
def dense_pipeline():
    import cv2
    from scipy import ndimage
    rng = np.random.default_rng(3)
    w, h, levels = 320, 240, 4
    big = ndimage.gaussian_filter(rng.random((h + 16, w + 64)), 2.0)
    big = (255 * (big - big.min()) / (big.max() - big.min())).astype(np.uint8)
    disp_true = 12
    left0 = np.ascontiguousarray(big[8:8 + h, 32:32 + w])
    right0 = np.ascontiguousarray(big[8:8 + h, 32 + disp_true:32 + disp_true + w])
    left1 = np.ascontiguousarray(big[8:8 + h, 33:33 + w])             # the camera moved: 1 px image shift
    right1 = np.ascontiguousarray(big[8:8 + h, 33 + disp_true:33 + disp_true + w])
    cam = StereoCamera(w / 2., h / 2., 250., 250., 0.5, w, h)
    pipe = ref_pipe.DenseStereoPipeline(cam)
    assert pipe.pyrlevels == levels
    histories = []
    orig_solve = Problem.solve

    def recording_solve(self):
        res = orig_solve(self)
        histories.append(np.array(self._cost_history))
        return res
    Problem.solve = recording_solve
    try:
        pipe.track(left0, right0)
        pipe.track(left1, right1)
    finally:
        Problem.solve = orig_solve
    kf = pipe.keyframes[0]
    out = {'left0': left0, 'right0': right0, 'left1': left1, 'right1': right1,
           'camera': np.array([w / 2., h / 2., 250., 250., 0.5, w, h]), 'levels': levels,
           'T_final': row(pipe.T_c_w[-1]), 'n_solves': len(histories)}
    for l in range(levels):
        out['im_pyr_%d' % l] = kf.im_pyr[l]
        out['jac_%d' % l] = kf.jacobian[l]
        out['disp_%d' % l] = kf.disparity[l]
    for k, hst in enumerate(histories):
        out['history_%d' % k] = hst
    # one level on its own: the (SO3, t) form through Problem, first iteration traced
    lvl = 1
    pcam = pipe.pyr_cameras[pipe.pyrlevel_sequence.index(lvl)]
    tf = ref_pipe.DenseStereoKeyframe(left1, right1, levels)
    from pyslam.residuals import PhotometricResidualSE3
    res = PhotometricResidualSE3(pcam, kf.im_pyr[lvl], kf.disparity[lvl], tf.im_pyr[lvl], kf.jacobian[lvl], pipe.intensity_stiffness,
                                 pipe.depth_stiffness / 2. ** -lvl, pipe.min_grad)
    for const_t in (False, True):
        pr = Problem(pipe.motion_options)
        pr.add_residual_block(res, ['R_1_0', 't_1_0_1'], loss=pipe.loss)
        pr.initialize_params({'R_1_0': SO3.identity(), 't_1_0_1': np.zeros(3)})
        if const_t:
            pr.set_parameters_constant('t_1_0_1')
        dxs = traced_solve(pr)
        tag = 'split_constt' if const_t else 'split'
        out[tag + '_history'] = np.array(pr._cost_history)
        out[tag + '_dx0'] = dxs[0]
        out[tag + '_R_final'] = pr.param_dict['R_1_0'].mat.copy()
        out[tag + '_t_final'] = np.asarray(pr.param_dict['t_1_0_1']).copy()
    out['split_level'] = lvl
    out['loss_k'] = 10.0
    np.savez_compressed(os.path.join(OUT, 'dense_pipeline.npz'), **out)
    print('dense_pipeline: %d level solves, final pose' % len(histories), out['T_final'][9:], 'split history', out['split_history'][:4])


def dense_rgbd_pipeline():
    """The reference's DenseRGBDPipeline on two synthetic RGB-D frames (depth maps, RGBDCamera, depth pyramid)."""
    from scipy import ndimage
    rng = np.random.default_rng(4)
    w, h = 160, 120
    big = ndimage.gaussian_filter(rng.random((h + 16, w + 64)), 2.0)
    big = (255 * (big - big.min()) / (big.max() - big.min())).astype(np.uint8)
    im0 = np.ascontiguousarray(big[8:8 + h, 32:32 + w])
    im1 = np.ascontiguousarray(big[8:8 + h, 33:33 + w])
    depth0 = 4.0 + 0.5 * ndimage.gaussian_filter(rng.random((h, w)), 4.0)
    depth1 = depth0.copy()
    cam = RGBDCamera(w / 2., h / 2., 200., 200., w, h)
    pipe = ref_pipe.DenseRGBDPipeline(cam)
    histories = []
    orig_solve = Problem.solve

    def recording_solve(self):
        res = orig_solve(self)
        histories.append(np.array(self._cost_history))
        return res
    Problem.solve = recording_solve
    try:
        pipe.track(im0, depth0)
        pipe.track(im1, depth1)
    finally:
        Problem.solve = orig_solve
    kf = pipe.keyframes[0]
    out = {'im0': im0, 'im1': im1, 'depth0': depth0, 'depth1': depth1, 'camera': np.array([w / 2., h / 2., 200., 200., w, h]),
           'levels': pipe.pyrlevels, 'T_final': row(pipe.T_c_w[-1]), 'n_solves': len(histories)}
    for l in range(pipe.pyrlevels):
        out['depth_%d' % l] = kf.depth[l]
    for k, hst in enumerate(histories):
        out['history_%d' % k] = hst
    np.savez_compressed(os.path.join(OUT, 'dense_rgbd_pipeline.npz'), **out)
    print('dense_rgbd_pipeline: %d level solves, final pose' % len(histories), out['T_final'][9:], [len(x) for x in histories])


if __name__ == '__main__':
    which = sys.argv[1:] or ['motion_ransac', 'orientation', 'rgbd_camera', 'metrics', 'dense_pipeline', 'dense_rgbd_pipeline']
    for name in which:
        globals()[name]()

"""TEST INFRASTRUCTURE -- generate tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (where /root/reference exists):

    python oracle/make_golden.py

Imports verbatim `pyslam` through oracle/reference_loader.py (three import
shims + the oracle's liegroups restatement as the parameter type -- the real
`liegroups` is not installable here), runs the reference's own code on seeded
synthetic inputs and stores inputs + outputs.  The fixtures pin
  * oracle/gn_oracle.py  (tests/test_oracle_golden.py, CPU) and
  * the CUDA path        (tests/test_gpu_*.py, through the C ABI).
"""
import copy
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.reference_loader import load_reference  # noqa: E402

ref = load_reference()
warnings.simplefilter('ignore')
from liegroups import SE2, SE3, SO2, SO3  # noqa: E402  (oracle/liegroups)
from pyslam.problem import Options, Problem  # noqa: E402  (verbatim reference)
from pyslam.residuals import (PhotometricResidualSE3, PoseResidual, PoseToPoseResidual,  # noqa: E402
                              QuadraticResidual, ReprojectionResidual)
from pyslam.sensors import StereoCamera  # noqa: E402
from pyslam.utils import invsqrt, bilinear_interpolate  # noqa: E402
import pyslam.losses as ref_losses  # noqa: E402

from pyslam_b200 import synthetic  # noqa: E402  (input generators only)

OUT = os.path.join(ROOT, 'tests', 'golden')
os.makedirs(OUT, exist_ok=True)


def se3_from_row(row):
    return SE3(SO3(row[:9].reshape(3, 3)), row[9:])


def se2_from_row(row):
    return SE2(SO2(row[:4].reshape(2, 2)), row[4:])


def pose_rows(Ts, n):
    return np.array([np.concatenate([T.rot.mat.ravel(), T.trans]) for T in Ts]).reshape(len(Ts), n * n + n)


def traced_solve(problem):
    """problem.solve() while recording every (dx, cost) of solve_one_iter."""
    dxs = []
    orig = problem.solve_one_iter

    def wrapped():
        dx, cost = orig()
        dxs.append(np.array(dx))
        return dx, cost
    problem.solve_one_iter = wrapped
    problem.solve()
    problem.solve_one_iter = orig
    return dxs


# ------------------------------------------------------------------ losses
def golden_losses():
    x = np.concatenate([np.linspace(-6, 6, 97), [0., 1.5, -1.5, 1e-9, 2.0, -2.0, 5.0]])
    out = dict(x=x)
    for name, cls, k in [('l2', ref_losses.L2Loss, None), ('l1', ref_losses.L1Loss, None),
                         ('cauchy', ref_losses.CauchyLoss, 5.0), ('huber', ref_losses.HuberLoss, 1.5),
                         ('tukey', ref_losses.TukeyLoss, 2.0), ('tdist', ref_losses.TDistributionLoss, 3.0)]:
        L = cls() if k is None else cls(k)
        out[name + '_k'] = np.float64(0. if k is None else k)
        out[name + '_loss'] = np.asarray(L.loss(x), dtype=float)
        out[name + '_weight'] = np.asarray(L.weight(x.copy()), dtype=float)
    np.savez_compressed(os.path.join(OUT, 'losses.npz'), **out)


# ------------------------------------------------------------------ camera / utils
def golden_camera():
    rng = np.random.default_rng(1)
    cam = StereoCamera(*synthetic.BA_CAMERA)
    pts = np.stack([rng.uniform(-4, 4, 64), rng.uniform(-3, 3, 64), rng.uniform(2, 30, 64)], axis=1)
    uvd, J = cam.project(pts, compute_jacobians=True)
    xyz, Jt = cam.triangulate(uvd, compute_jacobians=True)
    valid = cam.is_valid_measurement(np.vstack([uvd, [[-1., 5., 3.], [10., 2000., 3.], [10., 10., -1.]]]))
    im = rng.random((12, 17))
    xs = rng.uniform(-2, 19, 200)
    ys = rng.uniform(-2, 14, 200)
    interp = bilinear_interpolate(im, xs, ys)
    M = np.array([[7., 2., 1.], [0., 3., -1.], [-3., 4., -2.]])
    np.savez_compressed(os.path.join(OUT, 'camera_utils.npz'), camera=np.array(synthetic.BA_CAMERA), pts=pts,
                        uvd=uvd, J=J, xyz=xyz, Jt=Jt, valid=np.asarray(valid),
                        im=im, xs=xs, ys=ys, interp=interp, invsqrt_in=M, invsqrt_out=np.real(invsqrt(M)),
                        invsqrt_diag=np.real(invsqrt(np.diag([1., 1., 2.]))))


# ------------------------------------------------------------------ single residual blocks
def golden_residuals():
    rng = np.random.default_rng(2)
    out = {}
    cam = StereoCamera(*synthetic.BA_CAMERA)
    S3 = np.real(invsqrt(np.diag([1., 1., 2.]))) + 0.05 * rng.standard_normal((3, 3))
    xi = 0.3 * rng.standard_normal((8, 6))
    pts = np.stack([rng.uniform(-3, 3, 8), rng.uniform(-2, 2, 8), rng.uniform(5, 20, 8)], axis=1)
    obs = rng.uniform(100, 900, (8, 3))
    r_all, JT_all, Jp_all, Trow = [], [], [], []
    for k in range(8):
        T = SE3.exp(xi[k])
        res = ReprojectionResidual(cam, obs[k], S3)
        r, (JT, Jp) = res.evaluate([T, pts[k]], [True, True])
        r_all.append(r); JT_all.append(JT); Jp_all.append(Jp); Trow.append(pose_rows([T], 3)[0])
    out.update(rp_S=S3, rp_T=np.array(Trow), rp_pts=pts, rp_obs=obs, rp_r=np.array(r_all),
               rp_JT=np.array(JT_all), rp_Jp=np.array(Jp_all))
    for name, G, dof, n in (('se3', SE3, 6, 3), ('se2', SE2, 3, 2)):
        S = np.eye(dof) + 0.1 * rng.standard_normal((dof, dof))
        scale = np.array([1.0, 0.5, 1e-3, 1e-6, 1e-9, 2.5, 3.1, 0.0])       # incl. tiny angles
        T1s, T2s, Tos, r1, r2, J1, J2 = [], [], [], [], [], [], []
        for k in range(8):
            T1 = G.exp(0.8 * rng.standard_normal(dof))
            d = rng.standard_normal(dof)
            d = scale[k] * d / np.linalg.norm(d)
            To = G.exp(0.5 * rng.standard_normal(dof))
            T2 = G.exp(d).dot(To.dot(T1))            # error transform = exp(d)
            ra = PoseResidual(To, S).evaluate([T1], [True])
            rb, (Ja, Jb) = PoseToPoseResidual(To, S).evaluate([T1, T2], [True, True])
            T1s.append(T1); T2s.append(T2); Tos.append(To)
            r1.append(ra[0]); r2.append(rb); J1.append(Ja); J2.append(Jb)
        out.update({name + '_S': S, name + '_T1': pose_rows(T1s, n), name + '_T2': pose_rows(T2s, n),
                    name + '_Tobs': pose_rows(Tos, n), name + '_r_pose': np.array(r1),
                    name + '_r_p2p': np.array(r2), name + '_J1': np.array(J1), name + '_J2': np.array(J2)})
        # exp / log / adjoint samples of the liegroups restatement itself
        xis = np.vstack([0.7 * rng.standard_normal((6, dof)), 1e-9 * rng.standard_normal((2, dof)), np.zeros((1, dof))])
        Ts = [G.exp(x) for x in xis]
        out.update({name + '_xi': xis, name + '_exp': pose_rows(Ts, n),
                    name + '_log': np.array([G.log(T) for T in Ts]),
                    name + '_adj': np.array([T.adjoint() for T in Ts])})
    np.savez_compressed(os.path.join(OUT, 'residuals.npz'), **out)


# ------------------------------------------------------------------ cubic notebook (C1)
class CubicResidual:
    """User-defined residual of examples/Fitting a cubic.ipynb cell 4 (4 scalar params)."""

    def __init__(self, x, y, stiffness):
        self.x, self.y, self.stiffness = x, y, stiffness

    def evaluate(self, params, compute_jacobians=None):
        a, b, c, d = params
        r = np.array([self.stiffness * (a * self.x**3 + b * self.x**2 + c * self.x + d - self.y)]).reshape(1)
        if compute_jacobians:
            full = [self.stiffness * self.x**3, self.stiffness * self.x**2, self.stiffness * self.x, self.stiffness]
            return r, [np.array(j) if cj else None for j, cj in zip(full, compute_jacobians)]
        return r


def golden_cubic():
    out = {}
    for n in (10, 20):
        x = np.linspace(-5, 5, n)
        y = 2. * x**3 + 4. * x**2 - 4. * x
        problem = Problem(Options())
        for xi, yi in zip(x, y):
            problem.add_residual_block(CubicResidual(xi, yi, 1.), ['a', 'b', 'c', 'd'])
        problem.initialize_params({'a': -2., 'b': 10., 'c': -6., 'd': -140.})
        dxs = traced_solve(problem)
        problem.compute_covariance()
        out.update({'n%d_x' % n: x, 'n%d_y' % n: y, 'n%d_dx0' % n: dxs[0], 'n%d_n_iters' % n: len(dxs),
                    'n%d_cost_history' % n: np.array(problem._cost_history),
                    'n%d_final' % n: np.array([float(np.squeeze(problem.param_dict[k])) for k in 'abcd']),
                    'n%d_cov' % n: problem._covariance_matrix})
    np.savez_compressed(os.path.join(OUT, 'cubic.npz'), **out)


# ------------------------------------------------------------------ pose graphs (C2 shape)
def golden_pose_graph(name, data, G, from_row, dof, n):
    opts = Options()
    opts.allow_nondecreasing_steps = True
    opts.max_nondecreasing_steps = 3
    problem = Problem(opts)
    keys = ['T_%d_0' % k for k in range(data['n'])]
    problem.add_residual_block(PoseResidual(from_row(data['prior_T']), data['prior_stiffness']), keys[0])
    for i, j, row in zip(data['odo_i'], data['odo_j'], data['odo_T']):
        problem.add_residual_block(PoseToPoseResidual(from_row(row), data['odo_stiffness']), [keys[i], keys[j]])
    for i, j, row in zip(data['loop_i'], data['loop_j'], data['loop_T']):
        problem.add_residual_block(PoseToPoseResidual(from_row(row), data['loop_stiffness']), [keys[i], keys[j]])
    problem.initialize_params({k: from_row(r) for k, r in zip(keys, data['T_init'])})
    problem._update_partition_dict = problem._get_update_partition_dict()
    H, g, cost = problem._get_precision_information_and_cost()
    t0 = time.perf_counter()
    dxs = traced_solve(problem)
    dt = time.perf_counter() - t0
    out = {k: np.asarray(v) for k, v in data.items()}
    out.update(H0=H.toarray(), g0=np.asarray(g).ravel(), cost0=np.float64(cost), dx0=dxs[0], dx1=dxs[1],
               n_iters=len(dxs), cost_history=np.array(problem._cost_history),
               T_final=pose_rows([problem.param_dict[k] for k in keys], n), ref_seconds=np.float64(dt))
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **out)
    print(name, 'iters', len(dxs), 'cost', problem._cost_history[0], '->', problem._cost_history[-1], '%.1fs' % dt)


# ------------------------------------------------------------------ stereo BA (C3 shape)
def golden_ba(name, n_kf, n_lm, loss_name, loss_k):
    d = synthetic.stereo_ba(n_kf, n_lm, seed=3)
    cam = StereoCamera(*d['camera'])
    loss = {'huber': ref_losses.HuberLoss, 'cauchy': ref_losses.CauchyLoss,
            'l2': lambda k: ref_losses.L2Loss()}[loss_name](loss_k)
    opts = Options()
    opts.allow_nondecreasing_steps = True
    opts.max_nondecreasing_steps = 3
    problem = Problem(opts)
    pk = ['T_cam%d_w' % k for k in range(n_kf)]
    qk = ['pt%d_w' % k for k in range(n_lm)]
    for ci, qi, o in zip(d['pose_idx'], d['pt_idx'], d['obs']):
        problem.add_residual_block(ReprojectionResidual(cam, o, d['stiffness']), [pk[ci], qk[qi]], loss)
    params = {k: SE3(SO3(R), t) for k, R, t in zip(pk, d['R0'], d['t0'])}       # poses first, then points
    params.update({k: p.copy() for k, p in zip(qk, d['pts0'])})
    problem.initialize_params(params)
    problem.set_parameters_constant(pk[0])
    problem._update_partition_dict = problem._get_update_partition_dict()
    t0 = time.perf_counter()
    H, g, cost = problem._get_precision_information_and_cost()
    t_lin = time.perf_counter() - t0
    t0 = time.perf_counter()
    dxs = traced_solve(problem)
    dt = time.perf_counter() - t0
    ones = np.ones(H.shape[0])
    out = {k: np.asarray(v) for k, v in d.items() if k not in ('loss', 'camera', 'intr')}
    out.update(camera=np.array(d['camera']), loss_name=loss_name, loss_k=np.float64(loss_k),
               H0_diag=H.diagonal(), H0_ones=H.dot(ones), H0_fro=np.float64(np.sqrt(H.multiply(H).sum())),
               H0_nnz=np.int64(H.nnz), g0=np.asarray(g).ravel(), cost0=np.float64(cost), dx0=dxs[0], dx1=dxs[1],
               n_iters=len(dxs), cost_history=np.array(problem._cost_history),
               R_final=np.array([problem.param_dict[k].rot.mat for k in pk]),
               t_final=np.array([problem.param_dict[k].trans for k in pk]),
               pts_final=np.array([problem.param_dict[k] for k in qk]),
               ref_seconds_linearize=np.float64(t_lin), ref_seconds_solve=np.float64(dt))
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **out)
    print(name, 'iters', len(dxs), 'cost', problem._cost_history[0], '->', problem._cost_history[-1],
          'linearize %.1fs solve %.1fs' % (t_lin, dt))


# ------------------------------------------------------------------ reference's own BA test (tests/test_problem.py:239-282)
def golden_ba_test():
    np.random.seed(42)
    cam = StereoCamera(640, 480, 1000, 1000, 0.25, 1280, 960)
    pts = [np.array([0., -1., 10.]), np.array([1., 1., 5.]), np.array([-1., 1., 15.])]
    poses = [SE3.identity(), SE3.exp(0.1 * np.ones(6)), SE3.exp(0.2 * np.ones(6)), SE3.exp(0.3 * np.ones(6))]
    obs = [[cam.project(T.dot(p)) for p in pts] for T in poses]
    S = invsqrt(np.diagflat([1, 1, 2]))
    opts = Options()
    opts.allow_nondecreasing_steps = True
    opts.max_nondecreasing_steps = 3
    problem = Problem(opts)
    for i, row in enumerate(obs):
        for j, o in enumerate(row):
            problem.add_residual_block(ReprojectionResidual(cam, o, S), ['T_cam%d_w' % i, 'pt%d_w' % j])
    init = {}
    pts_init = []
    for i in range(3):
        pts_init.append(cam.triangulate(obs[0][i] + 10. * np.random.rand(3)))
        init['pt%d_w' % i] = pts_init[-1]
    for i in range(4):
        init['T_cam%d_w' % i] = SE3.identity()
    problem.initialize_params(init)
    problem.set_parameters_constant('T_cam0_w')
    dxs = traced_solve(problem)
    np.savez_compressed(os.path.join(OUT, 'ba_reference_test.npz'), obs=np.array(obs), stiffness=np.real(S),
                        pts_init=np.array(pts_init), pts_true=np.array(pts), T_true=pose_rows(poses, 3),
                        dx0=dxs[0], n_iters=len(dxs), cost_history=np.array(problem._cost_history),
                        pts_final=np.array([problem.param_dict['pt%d_w' % i] for i in range(3)]),
                        T_final=pose_rows([problem.param_dict['T_cam%d_w' % i] for i in range(4)], 3))
    print('ba_reference_test iters', len(dxs), problem._cost_history)


# ------------------------------------------------------------------ covariance test (tests/test_problem.py:294-321)
def golden_covariance():
    opts = Options()
    opts.allow_nondecreasing_steps = True
    opts.max_nondecreasing_steps = 3
    problem = Problem(opts)
    odom = SE3.exp(0.1 * np.ones(6))
    So = invsqrt(1e-3 * np.eye(6))
    S0 = invsqrt(1e-6 * np.eye(6))
    problem.add_residual_block(PoseResidual(SE3.identity(), S0), 'T0')
    problem.add_residual_block(PoseToPoseResidual(odom, So), ['T0', 'T1'])
    problem.initialize_params({'T0': SE3.identity(), 'T1': SE3.identity()})
    problem.solve()
    problem.compute_covariance()
    np.savez_compressed(os.path.join(OUT, 'covariance_se3.npz'), odom=pose_rows([odom], 3)[0],
                        cov=problem._covariance_matrix, cov_T1=problem.get_covariance_block('T1', 'T1'),
                        cost_history=np.array(problem._cost_history),
                        T_final=pose_rows([problem.param_dict['T0'], problem.param_dict['T1']], 3))


# ------------------------------------------------------------------ dense photometric alignment (C5 shape)
def dense_options():
    o = Options()                     # pyslam/pipelines/dense.py:31-36
    o.allow_nondecreasing_steps = True
    o.max_nondecreasing_steps = 5
    o.min_cost_decrease = 0.99
    o.max_iters = 30
    o.num_threads = 1
    o.linesearch_max_iters = 0
    return o


def golden_photometric():
    d = synthetic.photometric_pair(64, 48, seed=1)
    cam = StereoCamera(*d['camera'])
    cam.compute_pixel_grid()
    disp = d['disparity'].copy()
    disp[5:9, 7:12] = np.nan                     # invalid disparities are dropped by the constructor
    res = PhotometricResidualSE3(cam, d['im_ref'], disp, d['im_track'], d['im_jac'], d['intensity_stiffness'],
                                 d['depth_stiffness'], min_grad=0.002)
    xi = np.array([0.01, -0.005, 0.02, 0.001, -0.002, 0.0015])
    r, (J,) = res.evaluate([SE3.exp(xi)], [True])
    problem = Problem(dense_options())
    problem.add_residual_block(res, ['T_1_0'], ref_losses.CauchyLoss(d['loss'][1]))
    problem.initialize_params({'T_1_0': SE3.identity()})
    problem._update_partition_dict = problem._get_update_partition_dict()
    H, g, cost = problem._get_precision_information_and_cost()
    dxs = traced_solve(problem)
    np.savez_compressed(os.path.join(OUT, 'photometric.npz'), camera=np.array(d['camera']), im_ref=d['im_ref'],
                        im_track=d['im_track'], disparity=disp, im_jac=d['im_jac'],
                        intensity_stiffness=np.float64(d['intensity_stiffness']),
                        depth_stiffness=np.float64(d['depth_stiffness']), min_grad=np.float64(0.002),
                        loss_k=np.float64(d['loss'][1]), xi=xi, n_ref=np.int64(len(res.im_ref)), r=r, J=J,
                        H0=H.toarray(), g0=np.asarray(g).ravel(), cost0=np.float64(cost), dx0=dxs[0], n_iters=len(dxs),
                        cost_history=np.array(problem._cost_history), T_final=pose_rows([problem.param_dict['T_1_0']], 3)[0])
    print('photometric n_ref', len(res.im_ref), 'valid', len(r), 'iters', len(dxs), problem._cost_history[:3], '...',
          problem._cost_history[-1])


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'photometric':
        golden_photometric()
        sys.exit(0)
    golden_losses()
    golden_camera()
    golden_residuals()
    golden_cubic()
    golden_pose_graph('posegraph_se2', synthetic.se2_pose_graph(60, 8, seed=4, loop_span=20), SE2, se2_from_row, 3, 2)
    golden_pose_graph('posegraph_se3', synthetic.se3_pose_graph(30, 5, seed=5, loop_span=10), SE3, se3_from_row, 6, 3)
    golden_ba_test()
    golden_covariance()
    golden_ba('ba_huber', 8, 120, 'huber', 1.5)
    golden_ba('ba_cauchy', 6, 60, 'cauchy', 2.0)
    golden_photometric()
    for f in sorted(os.listdir(OUT)):
        print('%8d  %s' % (os.path.getsize(os.path.join(OUT, f)), f))

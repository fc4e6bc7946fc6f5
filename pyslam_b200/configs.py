"""The BASELINE.json workloads (C2..C5) built on the PRODUCT API from the plain arrays of
`pyslam_b200.synthetic` -- no oracle, no reference code: this is what bench.py and the tools time.
(tests/builders.py builds the same problems on the oracle's classes as well.)"""
import numpy as np

from . import lie as L
from . import losses as Loss
from .problem import Options, Problem

LOSS_NAMES = {'l2': 0, 'l1': 1, 'cauchy': 2, 'huber': 3, 'tukey': 4, 'tdist': 5}


def make_loss(name, k=0.):
    if name in ('l2', 'l1'):
        return {'l2': Loss.L2Loss, 'l1': Loss.L1Loss}[name]()
    return {'cauchy': Loss.CauchyLoss, 'huber': Loss.HuberLoss, 'tukey': Loss.TukeyLoss,
            'tdist': Loss.TDistributionLoss}[name](k)


def nondecreasing_options():
    """Options of the reference's pose-graph / BA tests (tests/test_problem.py:156-161)."""
    o = Options()
    o.allow_nondecreasing_steps = True
    o.max_nondecreasing_steps = 3
    return o


def dense_options():
    """pyslam/pipelines/dense.py:31-36"""
    o = Options()
    o.allow_nondecreasing_steps = True
    o.max_nondecreasing_steps = 5
    o.min_cost_decrease = 0.99
    o.max_iters = 30
    o.linesearch_max_iters = 0
    return o


def ba_engine(d, device=0, fused=None):
    """Stereo BA (C3/C4, or one landmark shard of it) straight on the C ABI: (engine, initial pose table)."""
    from . import engine as E
    eng = E.Engine(device)
    Rt = np.concatenate([np.asarray(d['R0']).reshape(-1, 9), d['t0']], axis=1)
    eng.set_poses_se3(Rt, d['pose_const'])
    eng.set_points(d['pts0'])
    eng.add_reprojection_blocks(d['pose_idx'], d['pt_idx'], d['obs'], d['stiffness'], d['intr'],
                                LOSS_NAMES[d['loss'][0]], d['loss'][1])
    if fused is not None:
        eng.set_fused(fused)
    return eng, Rt


def ba_problem(d):
    """Stereo BA through the drop-in `Problem` API (bulk registration of the reprojection blocks)."""
    from .sensors import StereoCamera
    cam = StereoCamera(*[float(v) for v in np.asarray(d['camera'])])
    pk = ['T_cam%d_w' % k for k in range(len(d['R0']))]
    qk = ['pt%d_w' % k for k in range(len(d['pts0']))]
    pr = Problem(nondecreasing_options())
    pr.add_reprojection_batch(cam, [pk[i] for i in d['pose_idx']], [qk[i] for i in d['pt_idx']], d['obs'],
                              d['stiffness'], make_loss(*d['loss']))
    params = {k: L.SE3(L.SO3(R), t) for k, R, t in zip(pk, d['R0'], d['t0'])}
    params.update({k: np.array(p) for k, p in zip(qk, d['pts0'])})
    pr.initialize_params(params)
    pr.set_parameters_constant([k for k, c in zip(pk, d['pose_const']) if c])
    return pr


def pose_graph_problem(d, group='se2'):
    """C2: prior on pose 0 + odometry + loop closures (pyslam examples/posegraph_relax*.py)."""
    from .residuals import PoseResidual, PoseToPoseResidual
    if group == 'se2':
        fr = lambda row: L.SE2(L.SO2(row[:4].reshape(2, 2)), row[4:])
    else:
        fr = lambda row: L.SE3(L.SO3(row[:9].reshape(3, 3)), row[9:])
    keys = ['T_%d_0' % k for k in range(int(d['n']))]
    pr = Problem(nondecreasing_options())
    pr.add_residual_block(PoseResidual(fr(d['prior_T']), d['prior_stiffness']), keys[0])
    for i, j, row in zip(d['odo_i'], d['odo_j'], d['odo_T']):
        pr.add_residual_block(PoseToPoseResidual(fr(row), d['odo_stiffness']), [keys[i], keys[j]])
    for i, j, row in zip(d['loop_i'], d['loop_j'], d['loop_T']):
        pr.add_residual_block(PoseToPoseResidual(fr(row), d['loop_stiffness']), [keys[i], keys[j]])
    pr.initialize_params({k: fr(r) for k, r in zip(keys, d['T_init'])})
    return pr


def photometric_problem(d, min_grad=0., split=False):
    """C5: one PhotometricResidualSE3 block, Cauchy loss, the dense pipeline's options.  `split`: the (SO3, t)
    parameter form the reference's dense pipeline uses (pyslam/pipelines/dense.py:185-190)."""
    from .residuals import PhotometricResidualSE3
    from .sensors import StereoCamera
    cam = StereoCamera(*[float(v) for v in np.asarray(d['camera'])])
    cam.compute_pixel_grid()
    res = PhotometricResidualSE3(cam, d['im_ref'], d['disparity'], d['im_track'], d['im_jac'],
                                 float(d['intensity_stiffness']), float(d['depth_stiffness']), min_grad=min_grad)
    pr = Problem(dense_options())
    loss = make_loss(*d['loss'])
    if split:
        pr.add_residual_block(res, ['R_1_0', 't_1_0_1'], loss)
        pr.initialize_params({'R_1_0': L.SO3.identity(), 't_1_0_1': np.zeros(3)})
    else:
        pr.add_residual_block(res, ['T_1_0'], loss)
        pr.initialize_params({'T_1_0': L.SE3.identity()})
    return pr, res

"""Build the same problem three ways from plain arrays: with the oracle's
classes (OracleProblem), with the product's drop-in API (pyslam_b200.Problem)
and in the oracle's vectorised array form (BAArrays)."""
import numpy as np

from oracle import gn_oracle as O
from oracle import liegroups as OL

import pyslam_b200
from pyslam_b200 import lie as PL
from pyslam_b200 import losses as PLoss
from pyslam_b200.residuals import PoseResidual, PoseToPoseResidual, ReprojectionResidual
from pyslam_b200.sensors import StereoCamera


def nondecreasing_options(cls):
    o = cls()
    o.allow_nondecreasing_steps = True
    o.max_nondecreasing_steps = 3
    return o


def o_se3(row):
    return OL.SE3(OL.SO3(row[:9].reshape(3, 3)), row[9:])


def o_se2(row):
    return OL.SE2(OL.SO2(row[:4].reshape(2, 2)), row[4:])


def p_se3(row):
    return PL.SE3(PL.SO3(row[:9].reshape(3, 3)), row[9:])


def p_se2(row):
    return PL.SE2(PL.SO2(row[:4].reshape(2, 2)), row[4:])


def rows_of(Ts):
    return np.array([np.concatenate([np.asarray(T.rot.mat).ravel(), np.asarray(T.trans)]) for T in Ts])


def oracle_loss(name, k):
    return O.loss_from_kind(name, k)


def product_loss(name, k):
    return {'l2': PLoss.L2Loss, 'l1': PLoss.L1Loss}[name]() if name in ('l2', 'l1') else \
        {'cauchy': PLoss.CauchyLoss, 'huber': PLoss.HuberLoss, 'tukey': PLoss.TukeyLoss,
         'tdist': PLoss.TDistributionLoss}[name](k)


# ------------------------------------------------------------------ pose graphs
def pose_graph_keys(d):
    return ['T_%d_0' % k for k in range(int(d['n']))]


def _pose_graph(problem, d, from_row, Pose, P2P, loss=None):
    keys = pose_graph_keys(d)
    kw = {} if loss is None else {'loss': loss}
    problem.add_residual_block(Pose(from_row(d['prior_T']), d['prior_stiffness']), keys[0], **kw)
    for i, j, row in zip(d['odo_i'], d['odo_j'], d['odo_T']):
        problem.add_residual_block(P2P(from_row(row), d['odo_stiffness']), [keys[i], keys[j]], **kw)
    for i, j, row in zip(d['loop_i'], d['loop_j'], d['loop_T']):
        problem.add_residual_block(P2P(from_row(row), d['loop_stiffness']), [keys[i], keys[j]], **kw)
    problem.initialize_params({k: from_row(r) for k, r in zip(keys, d['T_init'])})
    return problem


def oracle_pose_graph(d, group, loss=None):
    fr = o_se3 if group == 'se3' else o_se2
    return _pose_graph(O.OracleProblem(nondecreasing_options(O.Options)), d, fr, O.PoseResidual, O.PoseToPoseResidual,
                       None if loss is None else oracle_loss(*loss))


def product_pose_graph(d, group, loss=None):
    fr = p_se3 if group == 'se3' else p_se2
    return _pose_graph(pyslam_b200.Problem(nondecreasing_options(pyslam_b200.Options)), d, fr, PoseResidual,
                       PoseToPoseResidual, None if loss is None else product_loss(*loss))


# ------------------------------------------------------------------ stereo BA
def ba_keys(d):
    return ['T_cam%d_w' % k for k in range(len(d['R0']))], ['pt%d_w' % k for k in range(len(d['pts0']))]


def ba_loss(d):
    if 'loss_name' in getattr(d, 'files', d):
        return str(d['loss_name']), float(d['loss_k'])
    return d['loss']


def oracle_ba_problem(d):
    """Block-by-block OracleProblem (small sizes)."""
    cam = O.StereoCamera(*d['camera'])
    loss = oracle_loss(*ba_loss(d))
    pk, qk = ba_keys(d)
    pr = O.OracleProblem(nondecreasing_options(O.Options))
    for ci, qi, o in zip(d['pose_idx'], d['pt_idx'], d['obs']):
        pr.add_residual_block(O.ReprojectionResidual(cam, o, d['stiffness']), [pk[ci], qk[qi]], loss)
    params = {k: OL.SE3(OL.SO3(R), t) for k, R, t in zip(pk, d['R0'], d['t0'])}
    params.update({k: np.array(p) for k, p in zip(qk, d['pts0'])})
    pr.initialize_params(params)
    pr.set_parameters_constant([k for k, c in zip(pk, d['pose_const']) if c])
    return pr


def oracle_ba_arrays(d):
    name, k = ba_loss(d)
    return O.BAArrays(d['R0'], d['t0'], d['pts0'], d['pose_idx'], d['pt_idx'], d['obs'], d['stiffness'],
                      tuple(np.asarray(d['camera'])[:5]), name, k, pose_const=d['pose_const'])


def product_ba_problem(d, bulk=False):
    cam = StereoCamera(*[float(v) for v in np.asarray(d['camera'])])
    loss = product_loss(*ba_loss(d))
    pk, qk = ba_keys(d)
    pr = pyslam_b200.Problem(nondecreasing_options(pyslam_b200.Options))
    if bulk:
        pr.add_reprojection_batch(cam, [pk[i] for i in d['pose_idx']], [qk[i] for i in d['pt_idx']], d['obs'],
                                  d['stiffness'], loss)
    else:
        for ci, qi, o in zip(d['pose_idx'], d['pt_idx'], d['obs']):
            pr.add_residual_block(ReprojectionResidual(cam, o, d['stiffness']), [pk[ci], qk[qi]], loss)
    params = {k: PL.SE3(PL.SO3(R), t) for k, R, t in zip(pk, d['R0'], d['t0'])}
    params.update({k: np.array(p) for k, p in zip(qk, d['pts0'])})
    pr.initialize_params(params)
    pr.set_parameters_constant([k for k, c in zip(pk, d['pose_const']) if c])
    return pr


class CubicResidual:
    """User-defined residual of the reference's `Fitting a cubic` notebook
    (cell 4): goes through the generic plug-in path."""

    def __init__(self, x, y, stiffness):
        self.x, self.y, self.stiffness = x, y, stiffness

    def evaluate(self, params, compute_jacobians=None):
        a, b, c, d = [float(np.squeeze(p)) for p in params]
        r = np.array([self.stiffness * (a * self.x**3 + b * self.x**2 + c * self.x + d - self.y)])
        if compute_jacobians:
            full = [self.stiffness * self.x**3, self.stiffness * self.x**2, self.stiffness * self.x, self.stiffness]
            return r, [np.array(j) if cj else None for j, cj in zip(full, compute_jacobians)]
        return r


# ------------------------------------------------------------------ dense photometric alignment
def dense_options(cls):
    """pyslam/pipelines/dense.py:31-36"""
    o = cls()
    o.allow_nondecreasing_steps = True
    o.max_nondecreasing_steps = 5
    o.min_cost_decrease = 0.99
    o.max_iters = 30
    o.linesearch_max_iters = 0
    return o


def oracle_photometric_problem(d, min_grad=0., options=None):
    cam = O.StereoCamera(*np.asarray(d['camera']))
    cam.compute_pixel_grid()
    res = O.PhotometricResidualSE3(cam, d['im_ref'], d['disparity'], d['im_track'], d['im_jac'],
                                   float(d['intensity_stiffness']), float(d['depth_stiffness']), min_grad=min_grad)
    pr = O.OracleProblem(options or dense_options(O.Options))
    pr.add_residual_block(res, ['T_1_0'], O.CauchyLoss(float(d['loss_k'])))
    pr.initialize_params({'T_1_0': OL.SE3.identity()})
    return pr, res


def product_photometric_problem(d, min_grad=0., options=None):
    from pyslam_b200.residuals import PhotometricResidualSE3
    cam = StereoCamera(*[float(v) for v in np.asarray(d['camera'])])
    cam.compute_pixel_grid()
    res = PhotometricResidualSE3(cam, d['im_ref'], d['disparity'], d['im_track'], d['im_jac'],
                                 float(d['intensity_stiffness']), float(d['depth_stiffness']), min_grad=min_grad)
    pr = pyslam_b200.Problem(options or dense_options(pyslam_b200.Options))
    pr.add_residual_block(res, ['T_1_0'], PLoss.CauchyLoss(float(d['loss_k'])))
    pr.initialize_params({'T_1_0': PL.SE3.identity()})
    return pr, res

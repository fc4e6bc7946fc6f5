"""GPU: the CUDA path (through pyslam_b200.Problem -> ctypes -> libbslam.so)
against (i) fixtures produced by the unmodified reference and (ii) the CPU
oracle on the same seeded inputs.

Tolerances.  north_star asks for 1e-6 relative on the update vector and the
final cost.  The tests below hold the first linearisation (H, g, cost) to
1e-11, update vectors to 1e-7 and cost histories to 1e-7 -- all arithmetic is
fp64; the residual difference is summation order (atomics) and Schur+Cholesky
versus SuperLU.
"""
import numpy as np
import pytest

from conftest import load_golden, rel_err
import builders as B

pytestmark = pytest.mark.gpu

TOL_LIN = 1e-11      # first linearisation: H, g, cost
TOL_DX = 1e-7        # update vector (north_star: 1e-6)
TOL_COST = 1e-7      # cost history / final cost (north_star: 1e-6)


def normal_equations_ref_order(pr):
    """Dense (H, b, cost) of the product at its current parameters, permuted to
    the reference's update ordering."""
    low = pr._ensure_lowered()
    if low.dense_active:
        pr._dense_linearize()
    cost = pr._engine.linearize()
    H, b = pr._engine.get_normal_equations(low.dim)
    idx = low.ref_from_internal
    return H[np.ix_(idx, idx)], b[idx], cost


def reduced_system_ref_order(pr, n):
    """(S, rhs) after `reduce`, restricted/permuted to the first n entries of the
    reference ordering (valid when those are the non-eliminated parameters)."""
    low = pr._low
    Sfull, rfull = pr._engine.get_reduced_system(low.layout['n_reduced'])
    idx = low.ref_from_internal[:n]
    assert idx.max() < low.layout['n_reduced']
    return Sfull[np.ix_(idx, idx)], rfull[idx]


@pytest.mark.parametrize('name', ['ba_huber', 'ba_cauchy'])
@pytest.mark.parametrize('bulk', [False, True])
def test_ba_against_reference_golden(name, bulk):
    g = load_golden(name)
    pr = B.product_ba_problem(g, bulk=bulk)
    H, b, cost = normal_equations_ref_order(pr)
    ones = np.ones(len(b))
    assert np.allclose(H, H.T, rtol=0, atol=0)
    assert rel_err(np.diag(H), g['H0_diag']) < TOL_LIN
    assert rel_err(H @ ones, g['H0_ones']) < TOL_LIN
    assert abs(np.sqrt((H * H).sum()) - g['H0_fro']) < TOL_LIN * g['H0_fro']
    assert int((H != 0).sum()) == int(g['H0_nnz'])
    assert rel_err(b, g['g0']) < TOL_LIN
    assert abs(cost - g['cost0']) < TOL_LIN * g['cost0']
    dx0, c1 = pr.solve_one_iter()
    assert rel_err(dx0, g['dx0']) < TOL_DX
    assert abs(c1 - g['cost_history'][1]) < TOL_COST * g['cost_history'][1]
    # solve_one_iter must not move the parameters
    dx0b, _ = pr.solve_one_iter()
    assert rel_err(dx0b, dx0) < 1e-12
    pr.solve()
    assert len(pr._cost_history) == len(g['cost_history'])
    np.testing.assert_allclose(pr._cost_history, g['cost_history'], rtol=TOL_COST)
    pk, qk = B.ba_keys(g)
    R = np.array([pr.param_dict[k].rot.mat for k in pk])
    t = np.array([pr.param_dict[k].trans for k in pk])
    pts = np.array([pr.param_dict[k] for k in qk])
    assert rel_err(R, g['R_final']) < 1e-7 and rel_err(t, g['t_final']) < 1e-6 and rel_err(pts, g['pts_final']) < 1e-7


@pytest.mark.parametrize('name,group', [('posegraph_se2', 'se2'), ('posegraph_se3', 'se3')])
def test_pose_graph_against_reference_golden(name, group):
    g = load_golden(name)
    pr = B.product_pose_graph(g, group)
    H, b, cost = normal_equations_ref_order(pr)
    assert rel_err(H, g['H0']) < TOL_LIN
    assert rel_err(b, g['g0']) < TOL_LIN
    assert abs(cost - g['cost0']) < TOL_LIN * g['cost0']
    dx0, _ = pr.solve_one_iter()
    assert rel_err(dx0, g['dx0']) < TOL_DX
    pr.solve()
    assert len(pr._cost_history) == len(g['cost_history'])
    np.testing.assert_allclose(pr._cost_history, g['cost_history'], rtol=TOL_COST)
    Tf = B.rows_of([pr.param_dict[k] for k in B.pose_graph_keys(g)])
    assert rel_err(Tf, g['T_final']) < 1e-7


@pytest.mark.parametrize('loss', [('huber', 0.4), ('cauchy', 0.5), ('tukey', 2.0), ('tdist', 3.0)])
def test_pose_graph_robust_losses_against_oracle(loss):
    from pyslam_b200 import synthetic
    d = synthetic.se2_pose_graph(40, 5, seed=11, loop_span=13)
    o = B.oracle_pose_graph(d, 'se2', loss)
    o._update_partition_dict = o._get_update_partition_dict()
    Ho, bo, co = o.get_precision_information_and_cost()
    pr = B.product_pose_graph(d, 'se2', loss)
    H, b, cost = normal_equations_ref_order(pr)
    assert rel_err(H, Ho.toarray()) < TOL_LIN
    assert rel_err(b, bo) < TOL_LIN
    assert abs(cost - co) < TOL_LIN * co
    o.solve()
    pr.solve()
    assert len(pr._cost_history) == len(o._cost_history)
    np.testing.assert_allclose(pr._cost_history, o._cost_history, rtol=TOL_COST)


def test_reference_bundle_adjust_test_case():
    # reference tests/test_problem.py:239-282, trace of examples/stereo_ba.py
    import pyslam_b200
    from pyslam_b200.lie import SE3
    from pyslam_b200.residuals import ReprojectionResidual
    from pyslam_b200.sensors import StereoCamera
    g = load_golden('ba_reference_test')
    cam = StereoCamera(640, 480, 1000, 1000, 0.25, 1280, 960)
    pr = pyslam_b200.Problem(B.nondecreasing_options(pyslam_b200.Options))
    for i in range(4):
        for j in range(3):
            pr.add_residual_block(ReprojectionResidual(cam, g['obs'][i, j], g['stiffness']),
                                  ['T_cam%d_w' % i, 'pt%d_w' % j])
    init = {'pt%d_w' % j: g['pts_init'][j].copy() for j in range(3)}
    init.update({'T_cam%d_w' % i: SE3.identity() for i in range(4)})
    pr.initialize_params(init)
    pr.set_parameters_constant('T_cam0_w')
    dx0, _ = pr.solve_one_iter()
    assert rel_err(dx0, g['dx0']) < TOL_DX
    out = pr.solve()
    assert len(pr._cost_history) == len(g['cost_history'])
    np.testing.assert_allclose(pr._cost_history[:4], g['cost_history'][:4], rtol=1e-6)
    assert ['%.6e' % c for c in pr._cost_history[:3]] == ['4.209652e+05', '1.921859e+04', '3.924267e+01']
    for j in range(3):
        assert np.linalg.norm(out['pt%d_w' % j] - g['pts_true'][j]) < 1e-4
    for i in range(4):
        assert np.linalg.norm(out['T_cam%d_w' % i].inv().dot(B.p_se3(g['T_true'][i])).log()) < 1e-4


@pytest.mark.parametrize('n', [10, 20])
def test_cubic_fit_plugin_path(n):
    """BASELINE config 1: user-defined Python residual (notebook's CubicResidual)
    through the generic plug-in path; assembly/solve/update on the GPU."""
    import pyslam_b200
    g = load_golden('cubic')
    pr = pyslam_b200.Problem()
    for xi, yi in zip(g['n%d_x' % n], g['n%d_y' % n]):
        pr.add_residual_block(B.CubicResidual(xi, yi, 1.), ['a', 'b', 'c', 'd'])
    pr.initialize_params({'a': -2., 'b': 10., 'c': -6., 'd': -140.})
    dx0, _ = pr.solve_one_iter()
    assert rel_err(dx0, g['n%d_dx0' % n]) < 1e-9
    out = pr.solve()
    assert len(pr._cost_history) == len(g['n%d_cost_history' % n])
    assert abs(pr._cost_history[0] - g['n%d_cost_history' % n][0]) < 1e-12 * pr._cost_history[0]
    np.testing.assert_allclose([float(np.squeeze(out[k])) for k in 'abcd'], [2., 4., -4., 0.], atol=1e-8)
    if n == 10:
        assert pr.summary().startswith('Iterations:   2 | Cost: 3.735817e+05 -->')


def test_quadratic_reference_tests():
    # reference tests/test_problem.py:45-79
    import pyslam_b200
    from pyslam_b200.residuals import QuadraticResidual
    pr = pyslam_b200.Problem()
    pr.add_residual_block(QuadraticResidual(1., 4., 0.5), ['a', 'b', 'c'])
    pr.add_residual_block(QuadraticResidual(0., 1., 2.), ['a', 'b', 'c'])
    pr.initialize_params({'a': 1., 'b': 2., 'c': 1.})
    assert pr.eval_cost() == 0.
    assert pr.eval_cost({'a': 1., 'b': 0., 'c': 0.}) == 3.125
    x = np.linspace(-5, 5, 10)
    y = x * x - 2. * x + 3.
    pr = pyslam_b200.Problem()
    for xi, yi in zip(x, y):
        pr.add_residual_block(QuadraticResidual(xi, yi, 1.), ['a', 'b', 'c'])
    pr.initialize_params({'a': -20., 'b': 10., 'c': -30.})
    dx0, _ = pr.solve_one_iter()
    np.testing.assert_allclose(dx0, [21., -12., 33.], rtol=1e-9)
    out = pr.solve()
    for k, v in {'a': 1., 'b': -2., 'c': 3.}.items():
        assert np.allclose(out[k], v)


def test_constant_parameters_and_mixed_blocks():
    """Constant landmarks / constant poses inside reprojection blocks, and a
    prior on a pose next to reprojection blocks (pose block + BA in one problem)."""
    from oracle import gn_oracle as O
    from oracle import liegroups as OL
    import pyslam_b200
    from pyslam_b200 import synthetic
    from pyslam_b200.lie import SE3, SO3
    from pyslam_b200.residuals import PoseResidual
    d = synthetic.stereo_ba(5, 40, track=4, seed=21)
    op = B.oracle_ba_problem(d)
    pp = B.product_ba_problem(d)
    pk, qk = B.ba_keys(d)
    S6 = np.diag([30., 30., 30., 80., 80., 80.])
    op.add_residual_block(O.PoseResidual(OL.SE3(OL.SO3(d['R_true'][2]), d['t_true'][2]), S6), pk[2])
    pp.add_residual_block(PoseResidual(SE3(SO3(d['R_true'][2]), d['t_true'][2]), S6), pk[2])
    for pr in (op, pp):
        pr.set_parameters_constant([qk[3], qk[17], pk[4]])
    op._update_partition_dict = op._get_update_partition_dict()
    Ho, bo, co = op.get_precision_information_and_cost()
    H, b, cost = normal_equations_ref_order(pp)
    assert H.shape == Ho.shape
    assert rel_err(H, Ho.toarray()) < TOL_LIN
    assert rel_err(b, bo) < TOL_LIN
    assert abs(cost - co) < TOL_LIN * co
    assert abs(pp.eval_cost() - op.eval_cost()) < 1e-12 * co
    op.solve()
    pp.solve()
    assert len(pp._cost_history) == len(op._cost_history)
    np.testing.assert_allclose(pp._cost_history, op._cost_history, rtol=TOL_COST)
    assert np.array_equal(pp.param_dict[qk[3]], d['pts0'][3])          # constants untouched


def test_config3_against_oracle():
    """BASELINE config 3: 50 keyframes x 5 000 landmarks x 30 000 reprojections, Huber(1.5)."""
    from oracle import gn_oracle as O
    from pyslam_b200 import synthetic
    d = synthetic.stereo_ba(50, 5000, seed=0)
    ba = B.oracle_ba_arrays(d)
    pr = B.product_ba_problem(d, bulk=True)
    low = pr._ensure_lowered()
    cost0 = pr._engine.eval_cost()
    assert abs(cost0 - O.ba_cost(ba)) < 1e-12 * cost0
    for it in range(2):
        ref = O.ba_iteration(ba)
        cost_lin, cost_new, dx_norm = pr._engine.iterate(0., True)
        dx = pr._engine.get_update(low.dim)[low.ref_from_internal]
        assert abs(cost_lin - ref['cost_lin']) < 1e-10 * ref['cost_lin']
        assert rel_err(dx, ref['dx']) < 1e-6, 'iteration %d' % it
        assert abs(dx_norm - np.linalg.norm(ref['dx'])) < 1e-6 * dx_norm
        assert abs(cost_new - ref['cost_new']) < 1e-6 * ref['cost_new']


def test_config4_against_full_size_golden(golden):
    """BASELINE config 4 -- the configuration every bench number is quoted on: 500 keyframes x 100 000 landmarks x
    600 000 reprojections, Huber(1.5) -- against tests/golden/c4_summary.npz, which oracle/make_c4_golden.py
    produced with the CPU oracle on the FULL 302 994-dimensional system (3 iterations, ~100 s each).  Through
    Problem -> ctypes -> C ABI, the default (fused panel) lowering.  Tolerance 1e-6 relative (north_star)."""
    from pyslam_b200 import synthetic
    g = golden('c4_summary')
    d = synthetic.stereo_ba(int(g['n_kf']), int(g['n_lm']), track=int(g['track']), seed=int(g['seed']))
    pr = B.product_ba_problem(d, bulk=True)
    low = pr._ensure_lowered()
    eng = pr._engine
    assert eng.fused_info()[0] > 0            # the product path of this shape is the fused panel path
    n_pose = g['dx_pose'].shape[1]
    for it in range(int(g['n_iter'])):
        cost_lin, cost_new, dx_norm = eng.iterate(0., True)
        dx = eng.get_update(low.dim)[low.ref_from_internal]
        assert abs(cost_lin - g['cost_lin'][it]) < 1e-9 * g['cost_lin'][it], 'iteration %d' % it
        assert rel_err(dx[:n_pose], g['dx_pose'][it]) < 1e-6, 'iteration %d' % it
        assert rel_err(dx[g['sample_idx']], g['dx_sample'][it]) < 1e-6, 'iteration %d' % it
        assert abs(dx_norm - g['dx_norm'][it]) < 1e-6 * dx_norm
        assert abs(cost_new - g['cost_new'][it]) < 1e-6 * g['cost_new'][it]


def test_long_tracks_use_the_tail_path():
    """Landmarks with more than 128 observations (beyond one landmark block) are
    linearised by the generic atomic kernel; mixed here with regular landmarks."""
    from oracle import gn_oracle as O
    from pyslam_b200 import synthetic
    d = synthetic.stereo_ba(140, 200, track=5, seed=1)
    d2 = synthetic.stereo_ba(140, 6, track=140, seed=2)
    n1 = len(d['pts0'])
    d['pts0'] = np.vstack([d['pts0'], d2['pts0']])
    d['pose_idx'] = np.concatenate([d['pose_idx'], d2['pose_idx']])
    d['pt_idx'] = np.concatenate([d['pt_idx'], d2['pt_idx'] + n1]).astype(np.int32)
    d['obs'] = np.vstack([d['obs'], d2['obs']])
    ba = B.oracle_ba_arrays(d)
    pr = B.product_ba_problem(d, bulk=True)
    Ho, bo, co = O.ba_linearize(ba)
    H, b, cost = normal_equations_ref_order(pr)
    assert rel_err(H, Ho.toarray()) < TOL_LIN
    assert rel_err(b, bo) < TOL_LIN
    assert abs(cost - co) < TOL_LIN * co
    low = pr._low
    for it in range(2):
        ref = O.ba_iteration(ba)
        cost_lin, cost_new, dx_norm = pr._engine.iterate(0., True)
        dx = pr._engine.get_update(low.dim)[low.ref_from_internal]
        assert rel_err(dx, ref['dx']) < 1e-6
        assert abs(cost_new - ref['cost_new']) < 1e-6 * ref['cost_new']


def test_random_visibility_blocks():
    """Landmarks seen by random subsets of poses (3..40 observations, poor camera
    locality): exercises many landmark-block shapes of the fast kernels."""
    from oracle import gn_oracle as O
    from pyslam_b200 import synthetic
    rng = np.random.default_rng(5)
    d = synthetic.stereo_ba(60, 300, track=6, seed=6)
    pose_idx, pt_idx = [], []
    for q in range(300):
        n = int(rng.integers(3, 41))
        pose_idx.append(rng.choice(60, size=n, replace=False))
        pt_idx.append(np.full(n, q))
    d['pose_idx'] = np.concatenate(pose_idx).astype(np.int32)
    d['pt_idx'] = np.concatenate(pt_idx).astype(np.int32)
    cam = O.StereoCamera(*d['camera'])
    pc = np.einsum('nij,nj->ni', d['R_true'][d['pose_idx']], d['pts_true'][d['pt_idx']]) + d['t_true'][d['pose_idx']]
    d['obs'] = cam.project(pc) + 0.3 * rng.standard_normal((len(pc), 3))
    ba = B.oracle_ba_arrays(d)
    pr = B.product_ba_problem(d, bulk=True)
    Ho, bo, co = O.ba_linearize(ba)
    H, b, cost = normal_equations_ref_order(pr)
    assert rel_err(H, Ho.toarray()) < TOL_LIN
    assert rel_err(b, bo) < TOL_LIN
    assert abs(cost - co) < TOL_LIN * co
    low = pr._low
    # reduced system against the Schur complement of the oracle's H (poses come first in the reference order)
    pr._engine.reduce(0.)
    n = 6 * 59
    S, rhs = reduced_system_ref_order(pr, n)
    Hd = Ho.toarray()
    Sref = Hd[:n, :n] - Hd[:n, n:] @ np.linalg.solve(Hd[n:, n:], Hd[n:, :n])
    rref = bo[:n] - Hd[:n, n:] @ np.linalg.solve(Hd[n:, n:], bo[n:])
    assert rel_err(S, Sref) < 1e-10
    assert rel_err(rhs, rref) < 1e-10
    for it in range(2):
        ref = O.ba_iteration(ba)
        cost_lin, cost_new, dx_norm = pr._engine.iterate(0., True)
        dx = pr._engine.get_update(low.dim)[low.ref_from_internal]
        assert rel_err(dx, ref['dx']) < 1e-6
        assert abs(cost_new - ref['cost_new']) < 1e-6 * ref['cost_new']


def _check_ba_against_oracle(d, n_iters=2):
    from oracle import gn_oracle as O
    ba = B.oracle_ba_arrays(d)
    pr = B.product_ba_problem(d, bulk=True)
    Ho, bo, co = O.ba_linearize(ba)
    H, b, cost = normal_equations_ref_order(pr)
    assert rel_err(H, Ho.toarray()) < TOL_LIN
    assert rel_err(b, bo) < TOL_LIN
    assert abs(cost - co) < TOL_LIN * co
    low = pr._low
    for it in range(n_iters):
        ref = O.ba_iteration(ba)
        cost_lin, cost_new, dx_norm = pr._engine.iterate(0., True)
        dx = pr._engine.get_update(low.dim)[low.ref_from_internal]
        assert rel_err(dx, ref['dx']) < 1e-6, 'iteration %d' % it
        assert abs(cost_new - ref['cost_new']) < 1e-6 * ref['cost_new']


def test_duplicate_observations_of_a_landmark_by_one_pose():
    """Two reprojection blocks on the same (pose, landmark) pair: legal in the reference (two rows of the
    Jacobian), outside the block kernels' one-observation-per-(slot, landmark) layout -> generic path."""
    from oracle import gn_oracle as O
    from pyslam_b200 import synthetic
    rng = np.random.default_rng(11)
    d = synthetic.stereo_ba(12, 150, track=5, seed=9)
    dup = rng.choice(len(d['pose_idx']), size=40, replace=False)
    d['pose_idx'] = np.concatenate([d['pose_idx'], d['pose_idx'][dup]]).astype(np.int32)
    d['pt_idx'] = np.concatenate([d['pt_idx'], d['pt_idx'][dup]]).astype(np.int32)
    d['obs'] = np.vstack([d['obs'], d['obs'][dup] + 0.3 * rng.standard_normal((40, 3))])
    _check_ba_against_oracle(d)


def test_short_tracks_many_landmarks_per_block():
    """Tracks of 2 observations: 64 landmarks per landmark block (the per-block landmark tables of the
    Schur / back-substitution kernels are sized by the largest block)."""
    from pyslam_b200 import synthetic
    d = synthetic.stereo_ba(30, 900, track=2, seed=13)
    _check_ba_against_oracle(d)


def test_iterate_host_equals_separate_calls():
    """bslam_iterate_host (upload + iterate + download, one synchronisation) against set / iterate / get on a
    second handle.  The two handles sum their fp64 atomics in different orders (~1e-16 relative per sum), and
    nothing else differs: ONE iteration agrees to 1e-10 whatever the loss; over several Gauss-Newton iterations the
    rounding is amplified by the solves (conditioning of the reduced system), so later iterations are compared at 1e-6,
    and the first iteration is also repeated on the SAME handle to show the run-to-run spread is of that size."""
    from pyslam_b200 import synthetic
    for loss in (('l2', 0.), ('huber', 1.5)):
        d = synthetic.stereo_ba(12, 300, track=5, seed=21)
        d['loss'] = loss
        pr1 = B.product_ba_problem(d, bulk=True); pr1._ensure_lowered()
        pr2 = B.product_ba_problem(d, bulk=True); pr2._ensure_lowered()
        e1, e2 = pr1._engine, pr2._engine
        Rt0 = np.ascontiguousarray(np.concatenate([d['R0'].reshape(-1, 9), d['t0']], axis=1))
        pts0 = np.ascontiguousarray(d['pts0'], dtype=np.float64)
        Rt, pts, Rt2, pts2 = Rt0.copy(), pts0.copy(), Rt0.copy(), pts0.copy()
        for it in range(3):
            e1.set_poses_se3(Rt); e1.set_points(pts)
            r1 = e1.iterate(0., True)
            e1.get_poses_se3(Rt); e1.get_points(pts)
            r2 = e2.iterate_host(Rt2, pts2, 0., True)
            tol = 1e-10 if it == 0 else 1e-6
            np.testing.assert_allclose(r2, r1, rtol=tol)
            np.testing.assert_allclose(Rt2, Rt, rtol=0, atol=tol)
            np.testing.assert_allclose(pts2, pts, rtol=0, atol=tol * 10)
        assert r1[1] < r1[0]
        # run-to-run spread of ONE handle on identical inputs (atomics order only)
        e1.set_poses_se3(Rt0); e1.set_points(pts0)
        a = e1.iterate(0., True)
        e1.set_poses_se3(Rt0); e1.set_points(pts0)
        b = e1.iterate(0., True)
        np.testing.assert_allclose(a, b, rtol=1e-10)


def test_mixed_groups_losses_and_stiffness():
    """Reprojection blocks with different losses and per-block stiffness in one
    problem (several constant groups inside one kernel launch)."""
    import pyslam_b200
    from oracle import gn_oracle as O
    from oracle import liegroups as OL
    from pyslam_b200 import synthetic
    from pyslam_b200.lie import SE3, SO3
    from pyslam_b200.residuals import ReprojectionResidual
    from pyslam_b200.sensors import StereoCamera
    d = synthetic.stereo_ba(6, 60, track=4, seed=9)
    rng = np.random.default_rng(9)
    pk, qk = B.ba_keys(d)
    ocam, pcam = O.StereoCamera(*d['camera']), StereoCamera(*d['camera'])
    op = O.OracleProblem(B.nondecreasing_options(O.Options))
    pp = pyslam_b200.Problem(B.nondecreasing_options(pyslam_b200.Options))
    for k, (ci, qi, o) in enumerate(zip(d['pose_idx'], d['pt_idx'], d['obs'])):
        S = d['stiffness'] * (1.0 + 0.1 * (k % 3)) + (0.01 * rng.standard_normal((3, 3)) if k % 5 == 0 else 0.)
        name, kk = [('huber', 1.5), ('cauchy', 2.0), ('l2', 0.), ('tdist', 4.0)][k % 4]
        op.add_residual_block(O.ReprojectionResidual(ocam, o, S), [pk[ci], qk[qi]], B.oracle_loss(name, kk))
        pp.add_residual_block(ReprojectionResidual(pcam, o, S), [pk[ci], qk[qi]], B.product_loss(name, kk))
    po = {k: OL.SE3(OL.SO3(R), t) for k, R, t in zip(pk, d['R0'], d['t0'])}
    po.update({k: np.array(p) for k, p in zip(qk, d['pts0'])})
    ppar = {k: SE3(SO3(R), t) for k, R, t in zip(pk, d['R0'], d['t0'])}
    ppar.update({k: np.array(p) for k, p in zip(qk, d['pts0'])})
    op.initialize_params(po)
    pp.initialize_params(ppar)
    op.set_parameters_constant(pk[0])
    pp.set_parameters_constant(pk[0])
    op._update_partition_dict = op._get_update_partition_dict()
    Ho, bo, co = op.get_precision_information_and_cost()
    H, b, cost = normal_equations_ref_order(pp)
    assert rel_err(H, Ho.toarray()) < TOL_LIN
    assert rel_err(b, bo) < TOL_LIN
    assert abs(cost - co) < TOL_LIN * co
    op.solve()
    pp.solve()
    assert len(pp._cost_history) == len(op._cost_history)
    np.testing.assert_allclose(pp._cost_history, op._cost_history, rtol=TOL_COST)


def test_config2_se2_pose_graph_full_size():
    """BASELINE config 2: SE(2) pose graph, 1 000 poses, prior + 999 odometry + 100
    loop-closure PoseToPoseResidual blocks, against the block-by-block oracle."""
    from pyslam_b200 import synthetic
    d = synthetic.se2_pose_graph(1000, 100, seed=0)
    o = B.oracle_pose_graph(d, 'se2')
    p = B.product_pose_graph(d, 'se2')
    o._update_partition_dict = o._get_update_partition_dict()
    Ho, bo, co = o.get_precision_information_and_cost()
    H, b, cost = normal_equations_ref_order(p)
    assert H.shape == (3000, 3000)
    assert rel_err(H, Ho.toarray()) < TOL_LIN
    assert rel_err(b, bo) < TOL_LIN
    assert abs(cost - co) < TOL_LIN * co
    dx0, _ = p.solve_one_iter()
    o.solve()
    assert rel_err(dx0, o.dx_history[0]) < 1e-6
    p.solve()
    assert len(p._cost_history) == len(o._cost_history)
    np.testing.assert_allclose(p._cost_history, o._cost_history, rtol=1e-6)
    Tp = B.rows_of([p.param_dict[k] for k in B.pose_graph_keys(d)])
    To = B.rows_of([o.param_dict[k] for k in B.pose_graph_keys(d)])
    assert rel_err(Tp, To) < 1e-6


def test_photometric_against_reference_golden():
    """BASELINE config 5 shape (small image): PhotometricResidualSE3 + CauchyLoss
    through the fused kernel against the unmodified reference."""
    g = load_golden('photometric')
    pr, _ = B.product_photometric_problem(g, min_grad=float(g['min_grad']))
    H, b, cost = normal_equations_ref_order(pr)
    assert rel_err(H, g['H0']) < 1e-10
    assert rel_err(b, g['g0']) < 1e-10
    assert abs(cost - g['cost0']) < 1e-11 * g['cost0']
    dx0, _ = pr.solve_one_iter()
    assert rel_err(dx0, g['dx0']) < TOL_DX
    pr.solve()
    assert len(pr._cost_history) == len(g['cost_history'])
    np.testing.assert_allclose(pr._cost_history, g['cost_history'], rtol=1e-6)
    assert rel_err(B.rows_of([pr.param_dict['T_1_0']])[0], g['T_final']) < 1e-6


def test_config5_photometric_full_size():
    """BASELINE config 5: 640x480 dense stereo photometric alignment, Cauchy(5),
    against the numpy oracle (first linearisation and the first update)."""
    from pyslam_b200 import synthetic
    d = synthetic.photometric_pair(640, 480, seed=0)
    d['loss_k'] = d['loss'][1]
    o, _ = B.oracle_photometric_problem(d)
    p, res = B.product_photometric_problem(d)
    assert len(res.im_ref) > 300000
    o._update_partition_dict = o._get_update_partition_dict()
    Ho, bo, co = o.get_precision_information_and_cost()
    H, b, cost = normal_equations_ref_order(p)
    assert rel_err(H, Ho.toarray()) < 1e-10
    assert rel_err(b, bo) < 1e-10
    assert abs(cost - co) < 1e-11 * co
    dxo, _ = o.solve_one_iter()
    dx, _ = p.solve_one_iter()
    assert rel_err(dx, dxo) < 1e-6
    assert abs(p.eval_cost() - o.eval_cost()) < 1e-11 * co


def test_covariance_reference_cases():
    """Problem.compute_covariance / get_covariance_block."""
    import pyslam_b200
    from pyslam_b200.lie import SE3
    from pyslam_b200.residuals import PoseResidual, PoseToPoseResidual
    from pyslam_b200.utils import invsqrt
    # (1) reference tests/test_problem.py:294-321
    g = load_golden('covariance_se3')
    pr = pyslam_b200.Problem(B.nondecreasing_options(pyslam_b200.Options))
    odom = SE3.exp(0.1 * np.ones(6))
    So, S0 = invsqrt(1e-3 * np.eye(6)), invsqrt(1e-6 * np.eye(6))
    pr.add_residual_block(PoseResidual(SE3.identity(), S0), 'T0')
    pr.add_residual_block(PoseToPoseResidual(odom, So), ['T0', 'T1'])
    pr.initialize_params({'T0': SE3.identity(), 'T1': SE3.identity()})
    pr.solve()
    pr.compute_covariance()
    est = pr.get_covariance_block('T1', 'T1')
    expected = np.linalg.inv(So.dot(So)) + odom.adjoint().dot(np.linalg.inv(S0.dot(S0)).dot(odom.adjoint().T))
    assert np.allclose(est, expected)
    np.testing.assert_allclose(pr._covariance_matrix, g['cov'], rtol=1e-6, atol=1e-12)
    # (2) cubic notebook (plug-in residuals), cells 8-12
    c = load_golden('cubic')
    pr = pyslam_b200.Problem()
    for xi, yi in zip(c['n10_x'], c['n10_y']):
        pr.add_residual_block(B.CubicResidual(xi, yi, 1.), ['a', 'b', 'c', 'd'])
    pr.initialize_params({'a': -2., 'b': 10., 'c': -6., 'd': -140.})
    pr.solve()
    pr.compute_covariance()
    np.testing.assert_allclose(pr._covariance_matrix, c['n10_cov'], rtol=1e-7, atol=1e-14)
    assert abs(pr.get_covariance_block('a', 'a') - 0.00017205419580419603) < 1e-11
    # (3) stereo BA with eliminated landmarks against the oracle's dense inverse
    gb = load_golden('ba_cauchy')
    o = B.oracle_ba_problem(gb)
    p = B.product_ba_problem(gb)
    o._update_partition_dict = o._get_update_partition_dict()
    o.compute_covariance()
    p._ensure_lowered()
    p.compute_covariance()
    assert rel_err(p._covariance_matrix, o._covariance_matrix) < 1e-6
    pk, qk = B.ba_keys(gb)
    np.testing.assert_allclose(p.get_covariance_block(pk[2], qk[5]), o.get_covariance_block(pk[2], qk[5]), rtol=1e-5, atol=1e-9)
    assert p.get_covariance_block(pk[0], pk[0]) is None        # constant parameter

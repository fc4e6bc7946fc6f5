"""CPU: host logic of the drop-in `Problem` (lowering, update ordering,
termination rules, plug-in blocks, write-back) with the engine replaced by an
oracle-backed test double (tests/fake_engine.py), compared with fixtures from
the unmodified reference and with the oracle's `OracleProblem`."""
import copy

import numpy as np
import pytest

from conftest import load_golden, rel_err
import builders as B
from fake_engine import FakeEngine
from oracle import gn_oracle as O

import pyslam_b200
from pyslam_b200 import engine as E
from pyslam_b200.residuals import QuadraticResidual


@pytest.fixture(autouse=True)
def fake_engine(monkeypatch, request):
    if 'real_engine' not in request.keywords:
        monkeypatch.setattr(E, 'Engine', FakeEngine)


# ---- API bookkeeping: reference tests/test_problem.py:12-43
def test_residual_blocks_and_param_dict_and_constants():
    pr = pyslam_b200.Problem()
    keys = ['a', 'b', 'c']
    pr.add_residual_block(QuadraticResidual(2., 4., 1.), keys)
    assert keys == pr.block_param_keys[0]
    pr.add_residual_block(QuadraticResidual(2., 4., 1.), 'a')
    assert pr.block_param_keys[1] == ['a']
    pr = pyslam_b200.Problem()
    params = {'a': 1, 'b': 2, 'c': 3}
    pr.initialize_params(params)
    assert pr.param_dict == params
    params.update({'d': 4})
    pr.initialize_params({'d': 4})
    assert pr.param_dict == params
    pr = pyslam_b200.Problem()
    pr.set_parameters_constant('a')
    assert pr.constant_param_keys == ['a']
    pr.set_parameters_constant(['a', 'b_param'])
    assert pr.constant_param_keys == ['a', 'b_param']
    pr.set_parameters_variable('a')
    assert pr.constant_param_keys == ['b_param']
    pr.set_parameters_variable('c')
    assert pr.constant_param_keys == ['b_param']
    pr.set_parameters_variable(['a', 'b_param', 'c'])
    assert pr.constant_param_keys == []


def test_options_defaults_match_reference():
    o, r = pyslam_b200.Options(), O.Options()
    for k, v in vars(r).items():
        assert getattr(o, k) == v, k
    assert o.lm_lambda == 0.
    # independent default Options per Problem (the reference shares one instance, problem.py:43)
    assert pyslam_b200.Problem().options is not pyslam_b200.Problem().options


def test_eval_cost_and_quadratic_fit():
    # reference tests/test_problem.py:45-79
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
    dx0, c = pr.solve_one_iter()
    np.testing.assert_allclose(dx0, [21., -12., 33.], rtol=1e-10)
    assert pr.param_dict == {'a': -20., 'b': 10., 'c': -30.}       # solve_one_iter does not move parameters
    out = pr.solve()
    assert abs(pr._cost_history[0] - 489552.5102880659) < 1e-6
    for k, v in {'a': 1., 'b': -2., 'c': 3.}.items():
        assert np.allclose(out[k], v)
    with pytest.raises(ValueError):
        pr.summary('nope')
    assert pr.summary().startswith('Iterations:')
    assert 'Rel change' in pr.summary('full')


@pytest.mark.parametrize('n', [10, 20])
def test_cubic_notebook_plugin(n):
    g = load_golden('cubic')
    pr = pyslam_b200.Problem()
    for xi, yi in zip(g['n%d_x' % n], g['n%d_y' % n]):
        pr.add_residual_block(B.CubicResidual(xi, yi, 1.), ['a', 'b', 'c', 'd'])
    pr.initialize_params({'a': -2., 'b': 10., 'c': -6., 'd': -140.})
    pr.solve()
    assert len(pr._cost_history) - 1 == int(g['n%d_n_iters' % n])
    np.testing.assert_allclose(pr._cost_history[0], g['n%d_cost_history' % n][0], rtol=1e-12)
    pr.compute_covariance()
    np.testing.assert_allclose(pr._covariance_matrix, g['n%d_cov' % n], rtol=1e-7, atol=1e-14)
    if n == 10:
        assert abs(pr.get_covariance_block('a', 'a') - 0.00017205419580419603) < 1e-11


@pytest.mark.parametrize('name', ['ba_huber', 'ba_cauchy'])
@pytest.mark.parametrize('bulk', [False, True])
def test_ba_ordering_and_termination(name, bulk):
    g = load_golden(name)
    pr = B.product_ba_problem(g, bulk=bulk)
    dx0, c1 = pr.solve_one_iter()
    assert rel_err(dx0, g['dx0']) < 1e-8           # update vector in the reference's ordering
    assert abs(c1 - g['cost_history'][1]) < 1e-9 * c1
    pr.solve()
    np.testing.assert_allclose(pr._cost_history, g['cost_history'], rtol=1e-8)
    pk, qk = B.ba_keys(g)
    assert rel_err(np.array([pr.param_dict[k] for k in qk]), g['pts_final']) < 1e-8
    assert rel_err(np.array([pr.param_dict[k].trans for k in pk]), g['t_final']) < 1e-8
    assert pr._low.all_fused and pr._engine.reproj


@pytest.mark.parametrize('name,group', [('posegraph_se2', 'se2'), ('posegraph_se3', 'se3')])
def test_pose_graph_ordering_and_termination(name, group):
    g = load_golden(name)
    pr = B.product_pose_graph(g, group)
    dx0, _ = pr.solve_one_iter()
    assert rel_err(dx0, g['dx0']) < 1e-8
    pr.solve()
    np.testing.assert_allclose(pr._cost_history, g['cost_history'], rtol=1e-8)
    assert rel_err(B.rows_of([pr.param_dict[k] for k in B.pose_graph_keys(g)]), g['T_final']) < 1e-8


def _noisy_pose_graph(seed):
    from pyslam_b200 import synthetic
    d = synthetic.se2_pose_graph(30, 4, seed=seed, loop_span=9)
    rng = np.random.default_rng(seed)
    d['T_init'] = d['T_init'].copy()
    d['T_init'][:, 4:] += 0.8 * rng.standard_normal((30, 2))     # bad start: non-monotone cost
    return d


@pytest.mark.parametrize('opts', [
    dict(),                                                     # defaults: stop on first non-decrease
    dict(linesearch_max_iters=0),                               # Appendix B.3: cost history lags, 1 iteration
    dict(max_iters=2),                                          # Appendix B.2: max_iters + 1 iterations
    dict(allow_nondecreasing_steps=True, max_nondecreasing_steps=2, min_cost_decrease=0.999),
    dict(allow_nondecreasing_steps=True, max_nondecreasing_steps=1),
    dict(min_cost=1e3),
    dict(min_update_norm=10.),
])
def test_termination_rules_follow_reference(opts):
    d = _noisy_pose_graph(7)
    o = B.oracle_pose_graph(d, 'se2')
    p = B.product_pose_graph(d, 'se2')
    for pr, cls in ((o, O.Options), (p, pyslam_b200.Options)):
        pr.options = cls()
        for k, v in opts.items():
            setattr(pr.options, k, v)
    o.solve()
    p.solve()
    assert len(p._cost_history) == len(o._cost_history)
    np.testing.assert_allclose(p._cost_history, o._cost_history, rtol=1e-8)
    Tp = B.rows_of([p.param_dict[k] for k in B.pose_graph_keys(d)])
    To = B.rows_of([o.param_dict[k] for k in B.pose_graph_keys(d)])
    assert rel_err(Tp, To) < 1e-8          # incl. the "best params" roll-back rule (problem.py:163-175)


def test_mixed_plugin_and_builtin_blocks():
    """A user-defined Python residual on a pose next to built-in factors, and a
    user-defined loss on a built-in residual (-> plug-in path for that block)."""
    from pyslam_b200.lie import SE2
    d = _noisy_pose_graph(3)

    class AnchorX:                      # pulls t_x of a pose towards a value
        def __init__(self, x):
            self.x = x

        def evaluate(self, params, compute_jacobians=None):
            T = params[0]
            r = np.array([3. * (T.trans[0] - self.x)])
            if compute_jacobians:
                # left perturbation: d t_x / d[rho_x, rho_y, phi] = [1, 0, -t_y]
                return r, [3. * np.array([[1., 0., -T.trans[1]]]) if compute_jacobians[0] else None]
            return r

    class MyHuber:                      # duck-typed loss, unknown to the library
        def loss(self, x):
            return O.HuberLoss(0.7).loss(x)

        def weight(self, x):
            return O.HuberLoss(0.7).weight(x)

    o = B.oracle_pose_graph(d, 'se2')
    p = B.product_pose_graph(d, 'se2')
    keys = B.pose_graph_keys(d)
    from pyslam_b200.residuals import PoseToPoseResidual
    o.add_residual_block(AnchorX(0.3), keys[11])
    p.add_residual_block(AnchorX(0.3), keys[11])
    o.add_residual_block(O.PoseToPoseResidual(B.o_se2(d['loop_T'][0]), np.eye(3)), [keys[2], keys[20]], MyHuber())
    p.add_residual_block(PoseToPoseResidual(B.p_se2(d['loop_T'][0]), np.eye(3)), [keys[2], keys[20]], MyHuber())
    assert abs(p.eval_cost() - o.eval_cost()) < 1e-10 * o.eval_cost()
    o.solve()
    p.solve()
    assert [k[0] for k in p._low.kinds].count('dense') == 2
    np.testing.assert_allclose(p._cost_history, o._cost_history, rtol=1e-8)
    Tp = B.rows_of([p.param_dict[k] for k in keys])
    To = B.rows_of([o.param_dict[k] for k in keys])
    assert rel_err(Tp, To) < 1e-8


@pytest.mark.real_engine
def test_no_gpu_fails_loudly():
    """Without a CUDA device (or without the built library) the product refuses
    to solve instead of falling back to a CPU path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    pr = pyslam_b200.Problem()
    pr.add_residual_block(QuadraticResidual(1., 4., 0.5), ['a', 'b', 'c'])
    pr.initialize_params({'a': 1., 'b': 2., 'c': 1.})
    with pytest.raises(E.EngineError):
        pr.solve()
    with pytest.raises(E.EngineError):
        pr.eval_cost()


def test_photometric_lowering_and_dense_options():
    """Config-5 shape through the host: PhotometricResidualSE3 is lowered to a
    photometric batch, linesearch_max_iters = 0 (cost history lags one
    iteration), nondecreasing-steps bookkeeping of dense.py."""
    g = load_golden('photometric')
    pr, _ = B.product_photometric_problem(g, min_grad=float(g['min_grad']))
    pr.solve()
    assert pr._low.kinds == [('photo', 'se3')] and pr._engine.photo_b
    assert len(pr._cost_history) - 1 == int(g['n_iters'])
    np.testing.assert_allclose(pr._cost_history, g['cost_history'], rtol=1e-7)
    assert pr._cost_history[0] == pr._cost_history[1]
    assert rel_err(B.rows_of([pr.param_dict['T_1_0']])[0], g['T_final']) < 1e-7
    # the (SO3, t) two-parameter form on an engine without bslam_add_photometric_block_split (this test double): plug-in path
    from pyslam_b200.lie import SE3
    pr2, res = B.product_photometric_problem(g, min_grad=float(g['min_grad']))
    pr2.residual_blocks, pr2.block_param_keys, pr2.block_loss_functions = [], [], []
    pr2.param_dict = {}
    pr2.add_residual_block(res, ['R_1_0', 't_1_0_1'], O.CauchyLoss(float(g['loss_k'])))
    pr2.initialize_params({'R_1_0': SE3.identity().rot, 't_1_0_1': np.zeros(3)})
    pr2.options.max_iters = 3
    pr2.solve()
    assert pr2._low.kinds == [('dense',)] and 'R_1_0' in pr2._low.opaque
    assert pr2._cost_history[-1] < 0.5 * pr2._cost_history[0]


def test_covariance_block_mapping():
    gb = load_golden('ba_cauchy')
    o = B.oracle_ba_problem(gb)
    p = B.product_ba_problem(gb)
    o._update_partition_dict = o._get_update_partition_dict()
    o.compute_covariance()
    p.compute_covariance()
    assert rel_err(p._covariance_matrix, o._covariance_matrix) < 1e-7
    pk, qk = B.ba_keys(gb)
    np.testing.assert_allclose(p.get_covariance_block(pk[2], qk[5]), o.get_covariance_block(pk[2], qk[5]), rtol=1e-6, atol=1e-10)


def test_landmark_parameters_are_row_views_updated_in_place():
    """initialize_params copies the float 3-vectors as rows of one block (single copy to / from the device) while
    every parameter stays its own ndarray object, updated in place as the reference's `+=` does (problem.py:405-409);
    replacing a parameter object, other shapes and aliased dicts fall back to the per-object path."""
    rng = np.random.default_rng(0)
    src = {'p%d' % i: rng.standard_normal(3) for i in range(100)}
    src['scalar'] = np.array([1.0])
    src['four'] = np.arange(4.0)
    keep = {k: v.copy() for k, v in src.items()}
    pr = pyslam_b200.Problem()
    pr.initialize_params(src)
    assert list(pr.param_dict) == list(src)                          # insertion order = update-vector order
    for k, v in src.items():
        assert pr.param_dict[k] is not v and np.array_equal(pr.param_dict[k], keep[k])
    src['p3'][:] = 7.0                                               # deep copy: the caller's arrays are not shared
    assert np.array_equal(pr.param_dict['p3'], keep['p3'])
    keys, block, views = pr._point_blocks[0]
    assert len(keys) == 100 and block.shape == (100, 3) and all(pr.param_dict[k] is v for k, v in zip(keys, views))
    held = pr.param_dict['p5']
    block[5] += 1.0                                                  # what a download does
    assert np.array_equal(held, keep['p5'] + 1.0)
    # aliased values: exactly copy.deepcopy (aliasing preserved), no block
    a = np.zeros(3)
    pr2 = pyslam_b200.Problem()
    pr2.initialize_params({'x': a, 'y': a, **{'q%d' % i: np.ones(3) for i in range(80)}})
    assert pr2.param_dict['x'] is pr2.param_dict['y'] and not pr2._point_blocks


@pytest.mark.parametrize('replace', [False, True])
def test_point_block_round_trip_through_the_engine(replace):
    g = load_golden('ba_huber')
    pr = B.product_ba_problem(g, bulk=True)
    pk, qk = B.ba_keys(g)
    if replace:      # a parameter object swapped behind the block's back: per-object path, same numbers
        pr.param_dict[qk[2]] = np.array(pr.param_dict[qk[2]])
    held = [pr.param_dict[k] for k in qk]
    pr.solve()
    assert (pr._point_block(pr.param_dict) is None) == (replace or len(qk) < 64)
    assert all(pr.param_dict[k] is h for k, h in zip(qk, held))      # updated in place
    assert rel_err(np.array(held), g['pts_final']) < 1e-8
    ref = pr._get_update_partition_dict()
    assert pr._update_partition_dict == ref and list(pr._update_partition_dict) == list(ref)

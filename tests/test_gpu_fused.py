"""GPU: the fused panel kernels (pyslam_b200/csrc/panel.cuh -- linearise + eliminate in one kernel, W never
written, back-substitution by re-linearisation) against the CPU oracle, in every lowering mode of
bslam_set_fused: 0 = landmark-block kernels with materialised W, 1 = dense panels fused (default),
2 = every run of landmarks that fits a panel.  Same tolerances as test_gpu_parity.py."""
import numpy as np
import pytest

from conftest import rel_err
import builders as B

pytestmark = pytest.mark.gpu


def _problem(d, mode, lam=0.):
    pr = B.product_ba_problem(d, bulk=True)
    pr.options.fused_mode = mode
    pr.options.lm_lambda = lam
    return pr


def _schur_of_oracle(H, b, n, lam=0.):
    Hd = H.toarray()
    if lam > 0.:
        Hd = Hd + lam * np.diag(np.diag(Hd))
    S = Hd[:n, :n] - Hd[:n, n:] @ np.linalg.solve(Hd[n:, n:], Hd[n:, :n])
    r = b[:n] - Hd[:n, n:] @ np.linalg.solve(Hd[n:, n:], b[n:])
    return S, r


def _check(d, mode, lam=0., n_iters=2, expect_panels=None, tol_dx=1e-6):
    from oracle import gn_oracle as O
    ba = B.oracle_ba_arrays(d)
    pr = _problem(d, mode, lam)
    low = pr._ensure_lowered()
    eng = pr._engine
    n_panels, n_fused = eng.fused_info()
    if expect_panels is not None:
        assert (n_panels > 0) == expect_panels, (n_panels, n_fused)
    if mode == 0:
        assert n_panels == 0
    # the reduced system the iteration factorises against the Schur complement of the oracle's H
    Ho, bo, co = O.ba_linearize(ba)
    n = 6 * int((~np.asarray(d['pose_const'])).sum())
    eng.linearize_reduce(lam)
    Sfull, rfull = eng.get_reduced_system(low.layout['n_reduced'])
    idx = low.ref_from_internal[:n]
    Sref, rref = _schur_of_oracle(Ho, bo, n, lam)
    assert rel_err(Sfull[np.ix_(idx, idx)], Sref) < 1e-10
    assert rel_err(rfull[idx], rref) < 1e-10
    assert abs(eng.scalars()[0] - co) < 1e-11 * co
    for it in range(n_iters):
        ref = O.ba_iteration(ba, lam=lam)
        cost_lin, cost_new, dx_norm = eng.iterate(lam, True)
        dx = eng.get_update(low.dim)[low.ref_from_internal]
        assert abs(cost_lin - ref['cost_lin']) < 1e-10 * ref['cost_lin']
        assert rel_err(dx, ref['dx']) < tol_dx, 'iteration %d' % it
        assert abs(dx_norm - np.linalg.norm(ref['dx'])) < 1e-6 * dx_norm
        assert abs(cost_new - ref['cost_new']) < 1e-6 * ref['cost_new']
    # final parameters
    Rt = eng.get_poses_se3()
    assert np.abs(Rt[:, :9].reshape(-1, 3, 3) - ba.R).max() < 1e-8
    assert np.abs(Rt[:, 9:] - ba.t).max() < 1e-8
    assert np.abs(eng.get_points() - ba.pts).max() < 1e-7
    return pr


@pytest.mark.parametrize('mode', [0, 1, 2])
def test_consecutive_tracks(mode):
    """The shape of BASELINE configs 3/4 (every landmark seen by 6 consecutive keyframes): all panels dense."""
    from pyslam_b200 import synthetic
    d = synthetic.stereo_ba(20, 1500, track=6, seed=3)
    _check(d, mode, expect_panels=mode > 0)


@pytest.mark.parametrize('mode', [1, 2])
def test_small_problem(mode):
    """150 landmarks: panels with few landmarks (mode 2) or none at all (mode 1 needs >= 16 landmarks ...)."""
    from pyslam_b200 import synthetic
    d = synthetic.stereo_ba(6, 150, seed=1)
    _check(d, mode)
    d = synthetic.stereo_ba(5, 12, track=4, seed=2)
    _check(d, mode, expect_panels=mode == 2)


@pytest.mark.parametrize('mode', [1, 2])
@pytest.mark.parametrize('lam', [1e-3, 0.5])
def test_lm_damping(mode, lam):
    """lambda * diag(H) damping (north_star's LM; extension of the reference, oracle leg: H + lam diag(H))."""
    from pyslam_b200 import synthetic
    d = synthetic.stereo_ba(15, 800, track=5, seed=4)
    _check(d, mode, lam=lam)


@pytest.mark.parametrize('loss', [('l2', 0.), ('cauchy', 1.0), ('tukey', 20.0), ('tdist', 4.0), ('huber', 0.7)])
def test_losses_through_panels(loss):
    from pyslam_b200 import synthetic
    d = synthetic.stereo_ba(12, 900, track=4, seed=5)
    d['loss'] = loss
    _check(d, 1, expect_panels=True)


def test_l1_loss():
    """L1Loss: weight 1/|x|, NaN at |x| <= 1e-8 (pyslam/losses.py:30-33); noisy data keep residuals away from 0."""
    from pyslam_b200 import synthetic
    d = synthetic.stereo_ba(10, 600, track=5, seed=6)
    d['loss'] = ('l1', 0.)
    for mode in (0, 1):
        _check(d, mode, n_iters=1, tol_dx=1e-6)


@pytest.mark.parametrize('mode', [1, 2])
def test_mixed_panels_and_blocks(mode):
    """Consecutive tracks (panels) + landmarks seen by random pose subsets (landmark blocks, materialised W) +
    a few landmarks seen by every pose + a second constant pose, one problem."""
    from oracle import gn_oracle as O
    from pyslam_b200 import synthetic
    rng = np.random.default_rng(8)
    d = synthetic.stereo_ba(40, 1200, track=6, seed=7)
    n0 = len(d['pts0'])
    extra_pose, extra_pt = [], []
    d2 = synthetic.stereo_ba(40, 160, track=6, seed=17)
    for q in range(150):
        n = int(rng.integers(3, 20))
        extra_pose.append(rng.choice(40, size=n, replace=False))
        extra_pt.append(np.full(n, n0 + q))
    for q in range(150, 154):            # seen by every pose
        extra_pose.append(np.arange(40))
        extra_pt.append(np.full(40, n0 + q))
    d['pts0'] = np.vstack([d['pts0'], d2['pts0'][:154]])
    d['pts_true'] = np.vstack([d['pts_true'], d2['pts_true'][:154]])
    pi = np.concatenate(extra_pose).astype(np.int32)
    qi = np.concatenate(extra_pt).astype(np.int32)
    cam = O.StereoCamera(*d['camera'])
    pc = np.einsum('nij,nj->ni', d['R_true'][pi], d['pts_true'][qi]) + d['t_true'][pi]
    keep = pc[:, 2] > 1.0
    pi, qi, pc = pi[keep], qi[keep], pc[keep]
    d['pose_idx'] = np.concatenate([d['pose_idx'], pi]).astype(np.int32)
    d['pt_idx'] = np.concatenate([d['pt_idx'], qi]).astype(np.int32)
    d['obs'] = np.vstack([d['obs'], cam.project(pc) + 0.3 * rng.standard_normal((len(pc), 3))])
    d['pose_const'] = d['pose_const'].copy()
    d['pose_const'][7] = True
    pr = _check(d, mode, expect_panels=True)
    n_panels, n_fused = pr._engine.fused_info()
    assert 0 < n_fused < len(d['pts0'])


def test_panel_rows_limit():
    """Tracks of 12 poses: more rows than a panel holds -> no panels, landmark blocks only, same answer."""
    from pyslam_b200 import synthetic
    d = synthetic.stereo_ba(30, 400, track=12, seed=9)
    pr = _check(d, 2)
    assert pr._engine.fused_info()[0] == 0


def test_solve_histories_agree_between_modes():
    """Problem.solve() end to end (termination logic included) in the three lowering modes."""
    from pyslam_b200 import synthetic
    d = synthetic.stereo_ba(16, 1000, track=6, seed=10)
    hist = []
    for mode in (0, 1, 2):
        pr = _problem(d, mode)
        pr.solve()
        hist.append(np.array(pr._cost_history))
    assert len(hist[0]) == len(hist[1]) == len(hist[2])
    np.testing.assert_allclose(hist[1], hist[0], rtol=1e-9)
    np.testing.assert_allclose(hist[2], hist[0], rtol=1e-9)

"""Synthetic workloads of BASELINE.json / SURVEY.md 8(d): seeded generators
that return plain numpy arrays, so the same numbers can be fed to this
package, to the CPU oracle and to bench.py.  Host-side set-up code only."""
import numpy as np

from .lie import SE2, SE3

BA_CAMERA = (640., 480., 1000., 1000., 0.25, 1280, 960)   # tests/test_problem.py:213 of the reference


def stereo_ba(n_kf, n_lm, track=6, seed=0, obs_sigma=0.3, pose_sigma=0.01, point_sigma=0.05):
    """Stereo bundle adjustment, SURVEY 8(d) C3/C4: K keyframes on a gently
    curving path, L landmarks each seen by `track` consecutive keyframes,
    noisy (u,v,d) observations, perturbed initial poses/points, pose 0 exact
    and constant.  Returns a dict of arrays (true and initial values)."""
    rng = np.random.default_rng(seed)
    cu, cv, fu, fv, b = BA_CAMERA[:5]
    T_true = [SE3.exp(np.array([0.05 * k, 0., 0., 0., 0.002 * k, 0.])) for k in range(n_kf)]
    R_true = np.array([T.rot.mat for T in T_true])
    t_true = np.array([T.trans for T in T_true])
    k0 = rng.integers(0, n_kf - track + 1, size=n_lm)
    p_c = np.stack([rng.uniform(-4., 4., n_lm), rng.uniform(-3., 3., n_lm), rng.uniform(6., 20., n_lm)], axis=1)
    # p_w = T_k0^-1 p_c
    pts_true = np.einsum('nji,nj->ni', R_true[k0], p_c - t_true[k0])
    pose_idx = (k0[:, None] + np.arange(track)[None, :]).ravel()
    pt_idx = np.repeat(np.arange(n_lm), track)
    pc = np.einsum('nij,nj->ni', R_true[pose_idx], pts_true[pt_idx]) + t_true[pose_idx]
    iz = 1. / pc[:, 2]
    obs = np.stack([fu * pc[:, 0] * iz + cu, fv * pc[:, 1] * iz + cv, fu * b * iz], axis=1)
    obs = obs + obs_sigma * rng.standard_normal(obs.shape)
    R0, t0 = R_true.copy(), t_true.copy()
    for k in range(1, n_kf):
        d = SE3.exp(pose_sigma * rng.standard_normal(6))
        R0[k] = d.rot.mat @ R_true[k]
        t0[k] = d.rot.mat @ t_true[k] + d.trans
    pts0 = pts_true + point_sigma * rng.standard_normal(pts_true.shape)
    pose_const = np.zeros(n_kf, bool)
    pose_const[0] = True
    from .utils import invsqrt
    return dict(n_kf=n_kf, n_lm=n_lm, intr=(cu, cv, fu, fv, b), camera=BA_CAMERA,
                R_true=R_true, t_true=t_true, pts_true=pts_true,
                R0=R0, t0=t0, pts0=pts0, pose_idx=pose_idx.astype(np.int32), pt_idx=pt_idx.astype(np.int32),
                obs=obs, stiffness=np.real(invsqrt(np.diag([1., 1., 2.]))), pose_const=pose_const,
                loss=('huber', 1.5))


def se2_pose_graph(n=1000, n_loops=100, seed=0, loop_span=200):
    """SE(2) pose-graph relaxation, SURVEY 8(d) C2: a circular trajectory,
    noisy odometry, exact loop closures i -> i+loop_span, stiff prior on pose 0,
    dead-reckoned initial guess.  Poses are returned as (n,6) [R|t] rows."""
    rng = np.random.default_rng(seed)
    step = SE2.exp(np.array([0.1, 0., 2. * np.pi / 200.]))
    T = [SE2.identity()]
    for _ in range(n - 1):
        T.append(step.dot(T[-1]))
    odo = []
    for k in range(1, n):
        noise = SE2.exp(0.01 * rng.standard_normal(3))
        odo.append(noise.dot(T[k].dot(T[k - 1].inv())))
    li = rng.integers(0, n - loop_span, size=n_loops) if n > loop_span else np.zeros(0, int)
    lj = li + loop_span
    loops = [T[j].dot(T[i].inv()) for i, j in zip(li, lj)]
    init = [SE2.identity()]
    for k in range(1, n):
        init.append(odo[k - 1].dot(init[-1]))
    row = lambda X: np.concatenate([X.rot.mat.ravel(), X.trans])
    from .utils import invsqrt
    return dict(n=n, T_true=np.array([row(X) for X in T]), T_init=np.array([row(X) for X in init]),
                odo_i=np.arange(0, n - 1, dtype=np.int32), odo_j=np.arange(1, n, dtype=np.int32),
                odo_T=np.array([row(X) for X in odo]),
                loop_i=li.astype(np.int32), loop_j=lj.astype(np.int32),
                loop_T=np.array([row(X) for X in loops]).reshape(-1, 6),
                prior_T=row(T[0]),
                prior_stiffness=np.real(invsqrt(1e-12 * np.eye(3))),
                odo_stiffness=np.real(invsqrt(1e-3 * np.eye(3))),
                loop_stiffness=np.real(invsqrt(1e-2 * np.eye(3))))


def se3_pose_graph(n=40, n_loops=6, seed=0, loop_span=10):
    """SE(3) analogue of `se2_pose_graph` (helix trajectory); poses as (n,12)
    [R|t] rows.  Shapes follow reference examples/posegraph_relax.py."""
    rng = np.random.default_rng(seed)
    step = SE3.exp(np.array([0.2, 0.01, 0.03, 0.02, -0.01, 2. * np.pi / 25.]))
    T = [SE3.identity()]
    for _ in range(n - 1):
        T.append(step.dot(T[-1]))
    odo = []
    for k in range(1, n):
        noise = SE3.exp(0.01 * rng.standard_normal(6))
        odo.append(noise.dot(T[k].dot(T[k - 1].inv())))
    li = rng.integers(0, n - loop_span, size=n_loops) if n > loop_span else np.zeros(0, int)
    lj = li + loop_span
    loops = [T[j].dot(T[i].inv()) for i, j in zip(li, lj)]
    init = [SE3.identity()]
    for k in range(1, n):
        init.append(odo[k - 1].dot(init[-1]))
    row = lambda X: np.concatenate([X.rot.mat.ravel(), X.trans])
    from .utils import invsqrt
    return dict(n=n, T_true=np.array([row(X) for X in T]), T_init=np.array([row(X) for X in init]),
                odo_i=np.arange(0, n - 1, dtype=np.int32), odo_j=np.arange(1, n, dtype=np.int32),
                odo_T=np.array([row(X) for X in odo]),
                loop_i=li.astype(np.int32), loop_j=lj.astype(np.int32),
                loop_T=np.array([row(X) for X in loops]).reshape(-1, 12),
                prior_T=row(T[0]),
                prior_stiffness=np.real(invsqrt(1e-6 * np.eye(6))),
                odo_stiffness=np.real(invsqrt(1e-3 * np.eye(6))),
                loop_stiffness=np.real(invsqrt(1e-2 * np.eye(6))))


def photometric_pair(w=640, h=480, seed=0, shift=1):
    """Dense stereo photometric alignment, SURVEY 8(d) C5: a smooth random
    reference image (Gaussian-blurred noise, normalised to [0,1]), a tracking
    image = the reference shifted horizontally by `shift` pixels, disparity
    20 + U[0,10], image Jacobian 0.5 * Sobel (pyslam/pipelines/keyframes.py:44-45).
    Returns a dict; camera = (cu, cv, fu, fv, b, w, h)."""
    from scipy import ndimage
    rng = np.random.default_rng(seed)
    big = ndimage.gaussian_filter(rng.random((h, w + 8)), 3.0)
    big = (big - big.min()) / (big.max() - big.min())
    im_ref = np.ascontiguousarray(big[:, 4:4 + w])
    im_track = np.ascontiguousarray(big[:, 4 + shift:4 + shift + w])
    disp = 20. + 10. * rng.random((h, w))
    kx = np.array([[-1., 0., 1.], [-2., 0., 2.], [-1., 0., 1.]])
    gradx = 0.5 * ndimage.correlate(im_ref, kx, mode='reflect')
    grady = 0.5 * ndimage.correlate(im_ref, kx.T, mode='reflect')
    scale = w / 640.
    camera = (320. * scale, 240. * scale, 500. * scale, 500. * scale, 0.5, w, h)
    return dict(camera=camera, im_ref=im_ref, im_track=im_track, disparity=disp, im_jac=np.array([gradx, grady]),
                intensity_stiffness=100., depth_stiffness=2., loss=('cauchy', 5.0))

#!/usr/bin/env python
"""bench.py -- Gauss-Newton/LM iterations per second on the stereo bundle
adjustment of BASELINE.json (config 4: 500 keyframes x 100 000 landmarks x
600 000 reprojections, Huber(1.5), Schur complement), landmarks sharded over
N GPUs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one full iteration of the hot path: linearise (residuals,
Jacobians, robust weights, J^T W J / J^T W r assembly) + Schur complement +
reduced Cholesky + back-substitution + manifold retraction + cost at the new
point -- what `Problem.solve_one_iter()` + the update does in the reference
(pyslam/problem.py:143-156,182-194).

Timing protocol (both `value` and `e2e`): every step is timed on its own
(CUDA events on the library's stream for `value`, host wall clock around the
C-ABI call for `e2e`); BEFORE each step the L2 is flushed (a 256 MB buffer is
overwritten) and the device is idle, OUTSIDE the timed interval -- the fused
iteration's working set (~45 MB) would otherwise stay resident in the 126 MB
L2 from one iteration to the next.  The K per-step times are summed, max over
ranks.  The K-step measurement is repeated `rounds` times and the MEDIAN round
is reported (all rounds are listed); SM clocks / throttle reasons are sampled
by a separate process over all rounds.

Prints ONE JSON line (keys at the bottom of run_ours).  `value`: problem
resident in HBM.  `e2e`: through the C ABI with HOST (pinned) buffers -- every
step uploads the parameters, iterates and downloads the updated parameters
(`bslam_iterate_host`: one call, one synchronisation).  `parity`: the first
iterations of the same problem against the full-size CPU-oracle fixture
tests/golden/c4_summary.npz, at every N.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'LM iterations/sec on stereo BA (500 keyframes x 100000 landmarks x 600000 reprojections)'
N_KF, N_LM, TRACK = 500, 100000, 6
FLUSH_BYTES = 256 << 20


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))), 'measured (MEASURED_PEAKS.json)'
    except Exception:
        return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0}, 'fallback (B200_PROFILING.md)'


def fp64_peak():
    """fp64 pipe peak (DFMA and DMMA share it on sm_100a): measured by tools/fp64_peak.cu on this pool's B200
    (profiles/r2_fp64_peak.json), else the nominal datasheet figure."""
    try:
        j = json.load(open(os.path.join(ROOT, 'profiles', 'r2_fp64_peak.json')))
        return float(j['fp64_tflops']), 'measured (profiles/r2_fp64_peak.json: %s)' % j.get('how', 'tools/fp64_peak.cu')
    except Exception:
        return 37.0, 'nominal B200 fp64 (MEASURED_PEAKS.json has no fp64 entry)'


# ------------------------------------------------------------------ clocks: sampled by ANOTHER process
_SAMPLER = r'''
import sys, time
idx, path = int(sys.argv[1]), sys.argv[2]
import pynvml as nv
nv.nvmlInit()
h = nv.nvmlDeviceGetHandleByIndex(idx)
names = [('hw_slowdown', 'HwSlowdown'), ('hw_thermal_slowdown', 'HwThermalSlowdown'),
         ('sw_thermal_slowdown', 'SwThermalSlowdown'), ('sw_power_cap', 'SwPowerCap')]
bits = []
for name, a in names:
    v = getattr(nv, 'nvmlClocksEventReason' + a, None) or getattr(nv, 'nvmlClocksThrottleReason' + a, None)
    if v is not None:
        bits.append((name, v))
get = getattr(nv, 'nvmlDeviceGetCurrentClocksEventReasons', None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
with open(path, 'w') as f:
    f.write('max %d\n' % mx); f.flush()
    while True:
        r = get(h)
        f.write('%d %d %s\n' % (time.time_ns(), nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM),
                                ','.join(n for n, b in bits if r & b)))
        f.flush()
        time.sleep(0.002)
'''


class ClockSampler:
    """SM clock + throttle reasons during the measurement, polled through NVML every ~2 ms by a SEPARATE process
    (no GIL / scheduling interference with the timed loops); nvidia-smi -lms as the fallback."""

    def __init__(self, gpu):
        vis = os.environ.get('CUDA_VISIBLE_DEVICES')
        self.idx = gpu
        if vis:
            try:
                self.idx = int(vis.split(',')[gpu])
            except Exception:
                pass
        self.proc, self.path, self.mode, self.t0 = None, None, None, None

    def start(self):
        fd, self.path = tempfile.mkstemp(suffix='.clk')
        os.close(fd)
        try:
            import pynvml  # noqa: F401  (the child needs it)
            self.proc = subprocess.Popen([sys.executable, '-c', _SAMPLER, str(self.idx), self.path],
                                         stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            self.mode = 'nvml'
            for _ in range(200):          # wait until the child is sampling
                if os.path.getsize(self.path) > 0:
                    break
                time.sleep(0.01)
        except Exception:
            q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
                 'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
            try:
                self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.idx), '--query-gpu=' + q,
                                              '--format=csv,noheader,nounits', '-lms', '20'],
                                             stdout=open(self.path, 'w'), stderr=subprocess.DEVNULL)
                self.mode = 'smi'
            except Exception:
                self.proc = None
        self.t0 = time.time_ns()

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': []}
        if self.proc is None:
            return out
        t1 = time.time_ns()
        time.sleep(0.02)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')
        for line in open(self.path):
            f = line.split()
            try:
                if self.mode == 'nvml':
                    if f[0] == 'max':
                        mx = float(f[1])
                    elif self.t0 <= int(f[0]) <= t1:
                        sm.append(float(f[1]))
                        if len(f) > 2:
                            reasons.update(x for x in f[2].split(',') if x)
                else:
                    g = [x.strip() for x in line.split(',')]
                    sm.append(float(g[0]))
                    mx = max(mx or 0., float(g[1]))
                    reasons.update(n for n, v in zip(names, g[2:6]) if v.lower().startswith('active'))
            except Exception:
                continue
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_min_mhz=float(min(sm)), sm_max_mhz=mx, reasons=sorted(reasons),
                       samples=len(sm), source='NVML polled by a separate process every ~2 ms over all timed rounds'
                       if self.mode == 'nvml' else 'nvidia-smi -lms 20')
        return out


def build_engine(d, device):
    """Upload one BA problem through the C ABI (kept for the tools/ and tests that import it)."""
    from pyslam_b200 import configs
    eng, Rt = configs.ba_engine(d, device)
    eng.finalize()
    return eng, Rt


class Timer:
    """The timing protocol of the file header."""

    def __init__(self, torch, stream, world):
        self.torch, self.stream, self.world = torch, stream, world
        self.flush_buf = torch.empty(FLUSH_BYTES // 8, dtype=torch.float64, device='cuda')

    def flush(self):
        self.flush_buf.zero_()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, ms):
        if self.world > 1:
            import torch.distributed as dist
            t = self.torch.tensor([ms], dtype=self.torch.float64, device='cuda')
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
            self.torch.cuda.synchronize()

    def device_ms(self, step, steps, align=None):
        """Sum of the per-step device times (CUDA events on the library's stream).  `align`: enqueues a device-side
        rendezvous of the ranks (bslam_peer_barrier) right before the first event, so that every rank starts the
        step together -- each rank's flush / host synchronisation ends at a slightly different time."""
        ev = [(self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        self.barrier()
        for e0, e1 in ev:
            self.flush()
            if align is not None:
                align()
            e0.record(self.stream)
            step()
            e1.record(self.stream)
        self.barrier()
        return self.max_over_ranks(sum(e0.elapsed_time(e1) for e0, e1 in ev))

    def wall_ms(self, step, steps):
        """Sum of the per-step host wall times (each step ends with its own synchronisation)."""
        tot = 0.
        self.barrier()
        for _ in range(steps):
            self.flush()
            t0 = time.perf_counter()
            step()
            tot += time.perf_counter() - t0
        self.barrier()
        return self.max_over_ranks(tot * 1e3)


def c4_parity(solver, eng, d, rank, world, reset):
    """First iterations from the initial guess against tests/golden/c4_summary.npz (full-size CPU oracle)."""
    import torch
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'c4_summary.npz'))
    lay = eng.layout()
    n_pose = g['dx_pose'].shape[1]
    # reference order: the variable poses in table order (6 each), then the points (3 each)
    var = np.flatnonzero(~np.asarray(d['pose_const']))
    pose_src = (lay['se3'][var][:, None] + np.arange(6)[None, :]).ravel()
    s_idx = g['sample_idx'] - n_pose                    # entries of the landmark part, global numbering
    n_glob = int(g['n_lm'])
    local_of = np.full(n_glob, -1, np.int64)            # global landmark -> this rank's local landmark
    local_of[np.asarray(d.get('lm_ids', np.arange(n_glob)))] = np.arange(len(d['pts0']))
    mine = local_of[s_idx // 3] >= 0
    lm_src = lay['pt'][local_of[s_idx[mine] // 3]] + s_idx[mine] % 3
    reset()
    worst_dx, worst_cost = 0., 0.
    n_it = int(g['n_iter'])
    for it in range(n_it):
        cost_lin, cost_new, dx_norm = solver.iterate(0., True)
        dx = eng.get_update(lay['dim'])
        e_lm = float(np.sum((dx[lm_src] - g['dx_sample'][it][mine]) ** 2))
        r_lm = float(np.sum(g['dx_sample'][it][mine] ** 2))
        e_pose = float(np.sum((dx[pose_src] - g['dx_pose'][it]) ** 2)) if rank == 0 else 0.
        r_pose = float(np.sum(g['dx_pose'][it] ** 2)) if rank == 0 else 0.
        acc = np.array([e_lm + e_pose, r_lm + r_pose])
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor(acc, dtype=torch.float64, device='cuda')
            dist.all_reduce(t)
            acc = t.cpu().numpy()
        worst_dx = max(worst_dx, float(np.sqrt(acc[0] / acc[1])))
        worst_dx = max(worst_dx, abs(dx_norm - float(g['dx_norm'][it])) / float(g['dx_norm'][it]))
        worst_cost = max(worst_cost, abs(cost_lin - float(g['cost_lin'][it])) / float(g['cost_lin'][it]),
                         abs(cost_new - float(g['cost_new'][it])) / float(g['cost_new'][it]))
    return {'dx_rel_err': worst_dx, 'cost_rel_err': worst_cost, 'iterations': n_it, 'tolerance': 1e-6,
            'ok': bool(worst_dx < 1e-6 and worst_cost < 1e-6),
            'against': 'tests/golden/c4_summary.npz: CPU oracle (numpy assembly + SuperLU spsolve on the full 302 994-dim '
                       'system) on this exact problem; dx over the 2 994 pose entries + 4 096 sampled landmark entries, '
                       '||dx||, cost at the linearisation point and at x [+] dx; worst over the iterations'}


def series_entry(torch, eng, steps=30, warm=3):
    """iterations/s of an already lowered engine, same timing protocol (single GPU)."""
    eng.snapshot()
    for _ in range(warm):
        eng.iterate(0., True)
    eng.restore()
    tm = Timer(torch, eng.torch_stream(), 1)
    ms = tm.device_ms(lambda: eng.iterate(0., True), steps) / steps
    eng.restore()
    return {'iterations_per_s': round(1e3 / ms, 1), 'us_per_iteration': round(1e3 * ms, 1), 'steps': steps}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from pyslam_b200 import configs, synthetic
    from pyslam_b200.dist import build_sharded_ba

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit('launch with torch.distributed.run --nproc-per-node %d' % args.gpus)
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))

    full = synthetic.stereo_ba(N_KF, N_LM, track=TRACK, seed=0)
    solver, d, Rt0 = build_sharded_ba(full, rank, world, local)
    eng = solver.engine
    stream = eng.torch_stream()
    n_obs_total, n_obs = len(full['obs']), len(d['obs'])
    tm = Timer(torch, stream, world)

    def reset():
        eng.set_poses_se3(Rt0)
        eng.set_points(d['pts0'])

    step = lambda: solver.iterate(0., True)
    align = eng.peer_barrier if solver.mode == 'peer' else None      # ranks start every timed step together

    # ---- device-resident: K iterations of Gauss-Newton from the initial guess, `rounds` times ----
    for _ in range(max(3, args.warmup)):
        step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    rounds_ms, costs = [], []
    launches = 0
    for r in range(args.rounds):
        reset()
        l0 = eng.launch_count()
        costs = []
        rounds_ms.append(tm.device_ms(lambda: costs.append(step()), args.steps, align))
        launches = eng.launch_count() - l0
    ms = float(np.median(rounds_ms))
    ms_per_step = ms / args.steps
    value = 1000.0 / ms_per_step

    # ---- end to end through the C ABI with host buffers (pinned) ----
    pin_Rt = torch.from_numpy(Rt0.copy()).pin_memory()
    pin_pts = torch.from_numpy(np.ascontiguousarray(d['pts0'])).pin_memory()
    h2d = pin_Rt.numel() * 8 + pin_pts.numel() * 8
    d2h = pin_Rt.numel() * 8 + pin_pts.numel() * 8 + 16 * 8
    np_Rt, np_pts = pin_Rt.numpy(), pin_pts.numpy()

    def e2e_step():
        # host parameters -> device, one iteration, updated parameters -> the same (pinned) host buffers,
        # which are the inputs of the next step; one C-ABI call, one synchronisation
        return solver.iterate_host(np_Rt, np_pts, 0., True)

    for _ in range(3):
        e2e_step()
    e2e_rounds = []
    for r in range(max(1, min(3, args.rounds))):
        np_Rt[...] = Rt0
        np_pts[...] = d['pts0']
        e2e_rounds.append(tm.wall_ms(e2e_step, args.steps))
    clocks = sampler.stop() if rank == 0 else None
    e2e_ms = float(np.median(e2e_rounds))
    e2e_value = 1000.0 * args.steps / e2e_ms

    # ---- parity of this very configuration (every N) ----
    parity = c4_parity(solver, eng, d, rank, world, reset)

    # ---- per-phase CUDA events of the library (un-graphed pass; single GPU) ----
    t_phase = {}
    if world == 1:
        reset()
        eng.enable_timing(True)
        for _ in range(args.steps):
            tm.flush()
            step()
            for k, v in eng.timings().items():
                t_phase[k] = t_phase.get(k, 0.) + v / args.steps
        eng.enable_timing(False)

    # ---- the 3 M-observation series (track 30), every N ----
    series = {}
    try:
        full30 = synthetic.stereo_ba(N_KF, N_LM, track=30, seed=0)
        s30, d30, Rt30 = build_sharded_ba(full30, rank, world, local)
        e30 = s30.engine
        for _ in range(3):
            s30.iterate(0., True)
        e30.set_poses_se3(Rt30); e30.set_points(d30['pts0'])
        tm30 = Timer(torch, e30.torch_stream(), world)
        ms30 = tm30.device_ms(lambda: s30.iterate(0., True), 20, e30.peer_barrier if s30.mode == 'peer' else None) / 20
        series['C4 track 30 (500 kf x 100000 lm x 3000000 obs), %d GPU(s)' % world] = {
            'iterations_per_s': round(1e3 / ms30, 1), 'us_per_iteration': round(1e3 * ms30, 1), 'steps': 20,
            'fused_panels': int(e30.fused_info()[0])}
        del s30, e30, tm30
    except Exception as ex:      # reported, never fatal for the headline
        series['C4 track 30'] = {'error': repr(ex)[:300]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    if world == 1:
        try:
            p = configs.pose_graph_problem(synthetic.se2_pose_graph(1000, 100, seed=0), 'se2')
            p._ensure_lowered()
            series['C2 SE(2) pose graph, 1000 poses, 1100 factors'] = series_entry(torch, p._engine)
            p = configs.ba_problem(synthetic.stereo_ba(50, 5000, seed=0))
            p._ensure_lowered()
            series['C3 stereo BA 50 x 5000 x 30000, Huber'] = series_entry(torch, p._engine)
            p, res = configs.photometric_problem(synthetic.photometric_pair(640, 480, seed=0))
            p._ensure_lowered()
            series['C5 dense photometric 640x480 (%d px), Cauchy' % len(res.im_ref)] = series_entry(torch, p._engine)
        except Exception as ex:
            series['C2/C3/C5'] = {'error': repr(ex)[:300]}

    # ---- the drop-in API end to end: Problem.solve() on the same problem (host lowering + iterations + download) ----
    api = None
    if world == 1:
        try:
            t0 = time.perf_counter()
            pr = configs.ba_problem(full)
            t_build = time.perf_counter() - t0
            t0 = time.perf_counter()
            pr.solve()                                   # fresh problem: lowering happens inside
            torch.cuda.synchronize()
            t_total = time.perf_counter() - t0
            pr2 = configs.ba_problem(full)
            t0 = time.perf_counter()
            pr2._ensure_lowered()                        # the lowering on its own, second instance
            torch.cuda.synchronize()
            t_lower = time.perf_counter() - t0
            del pr2
            n_it = len(pr._cost_history) - 1
            api = {'iterations': n_it, 'final_cost': float(pr._cost_history[-1]),
                   'register_blocks_s': t_build, 'solve_s_total': t_total, 'lower_s': t_lower,
                   'iterations_per_s_total': n_it / t_total,
                   'what': 'pyslam_b200.Problem with 500 SE3 + 100 000 point parameters (string keys) and one '
                           'add_reprojection_batch of 600 000 blocks; lower_s = key -> table lowering + bslam_finalize (ordering, '
                           'panels, uploads), once per problem, measured on a second instance; solve_s_total = Problem.solve() on the fresh '
                           'problem: that lowering + eval_cost + iterations with the reference termination logic + download into the '
                           '100 500 parameter objects -- the drop-in API is bound by the one-time lowering (bulk key handling in Python + bslam_finalize), not by the iterations'}
            del pr
        except Exception as ex:
            api = {'error': repr(ex)[:300]}

    pk, pk_src = peaks()
    n_pan, n_fused = eng.fused_info()
    out = {
        'metric': METRIC, 'value': value, 'unit': 'iterations/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': max(3, args.warmup), 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'strong',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': 'stereo BA config 4: %d keyframes x %d landmarks x %d reprojections, track %d, '
                               'Huber(1.5), pose 0 constant, Gauss-Newton (lambda=0) with Schur complement'
                               % (N_KF, N_LM, n_obs_total, TRACK),
                   'parallelism': solver.describe(),
                   'l2_policy': 'L2 flushed before EVERY timed step (a %d MB buffer is overwritten, then the device idles; outside '
                                'the timed interval): the fused iteration never materialises W, so its working set '
                                '(observations 19 MB, points/V/V^-1 14 MB, reduced tiles 4 MB, updates) would fit the 126 MB L2'
                                % (FLUSH_BYTES >> 20),
                   'timing': 'per-step CUDA events on the library stream, summed over the K steps, max over ranks; '
                             'median of %d rounds' % args.rounds + ('; N > 1: a device-side rendezvous of the ranks '
                             '(bslam_peer_barrier) is enqueued before the first event of every step' if align else ''),
                   'reduced_system_dim': 6 * (N_KF - 1), 'fused_panels': int(n_pan), 'fused_landmarks': int(n_fused)},
        'rounds_ms_per_step': [r / args.steps for r in rounds_ms],
        'clocks': clocks,
        'e2e': {'value': e2e_value, 'unit': 'iterations/s', 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                'ms_per_step': e2e_ms / args.steps, 'rounds_ms_per_step': [r / args.steps for r in e2e_rounds],
                'api': 'bslam_iterate_host (C ABI): pinned host pose/point tables up, one iteration, updated tables down; '
                       'host wall clock per step, same flush protocol'},
        'gpu_launches': int(launches),
        'e2e_problem_solve': api,
        'parity': parity,
        'final_cost': costs[-1][1], 'first_cost': costs[0][0],
        'series': series,
    }
    if world == 1:
        K, L = N_KF, len(d['pts0'])
        t_fused = t_phase.get('fused', 0.) * 1e-3
        # a fused variant that never materialises W moves 32 N + 432 K + 120 L bytes
        alg_bytes = 32 * n_obs + 432 * K + 120 * L
        # algorithmic flops: 545 per observation + Schur complement per landmark with t observations:
        # Y = W V^-1 (108 t), S -= Y W^T over t (t + 1) / 2 block pairs (216 each), rhs -= Y b_p (36 t)
        t_obs = TRACK
        alg_flops = 545.0 * n_obs + L * (108.0 * t_obs + 216.0 * t_obs * (t_obs + 1) / 2 + 36.0 * t_obs)
        f64_pk, f64_src = fp64_peak()
        traffic = None
        try:      # DRAM bytes of one launch from the committed `ncu --set full` capture of the same kernel
            traffic = json.load(open(os.path.join(ROOT, 'profiles', 'roofline_traffic.json'))).get('fused_panel_kernel')
        except Exception:
            pass
        if t_fused > 0:
            out['roofline'] = {
                'kernel': 'fused_panel_kernel (linearise + eliminate, W never written)', 'bound': 'tensor',
                'achieved': alg_flops / t_fused / 1e12, 'peak': f64_pk, 'unit': 'TFLOP/s',
                'frac': alg_flops / t_fused / 1e12 / f64_pk, 'traffic': traffic,
                'note': 'fp64 pipe (DFMA + DMMA.8x8x4 share it on sm_100a; tcgen05 has no fp64 kind): the fused kernel is '
                        'bound by it, not by HBM -- see roofline_hbm for the byte view',
                'peak_source': f64_src, 'algorithmic_flops_per_launch': alg_flops,
                'timed_in': 'un-graphed pass of the same K steps with per-kernel CUDA events on the launching stream, L2 flushed',
                'avg_launch_ms': t_fused * 1e3}
            out['roofline_hbm'] = {
                'kernel': 'fused_panel_kernel', 'bound': 'hbm', 'achieved': alg_bytes / t_fused / 1e9, 'peak': pk['hbm_gbs'],
                'unit': 'GB/s', 'frac': alg_bytes / t_fused / 1e9 / pk['hbm_gbs'], 'traffic': traffic,
                'algorithmic_bytes_per_launch': alg_bytes, 'formula': '32 N_obs + 432 N_cam + 120 N_pt (SURVEY 8d, fused variant)',
                'peak_source': pk_src,
                'note': 'the materialised-W formulation this replaces moved 176 N + 432 K + 120 L = 117.8 MB in the assembly '
                        'kernel alone and ~347 MB per iteration; the fused iteration moves ~75 MB'}
        out['phase_ms'] = {'prepare': t_phase.get('linearize', 0.), 'fused_panel': t_phase.get('fused', 0.),
                           'schur_blocks': max(0., t_phase.get('schur', 0.) - t_phase.get('fused', 0.)),
                           'cholesky_solve': t_phase.get('cholesky', 0.), 'retract_poses': t_phase.get('retract', 0.),
                           'panel_finish': t_phase.get('cost', 0.), 'total': t_phase.get('total', 0.)}
        out['phase_ms_note'] = ('un-graphed pass (event + launch gaps included); panel_finish = landmark back-substitution '
                                '(W^T dx_c recomputed) + retraction + cost at the new point, fused')
        out['cpu_baseline'] = cpu_baseline(budget_s=25.0)
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------- CPU legs (oracle as the reported baseline)
SAMPLES = [(50, 10000), (25, 5000), (10, 2000)]          # same landmark density as the full problem (200 / keyframe)


def oracle_problem(n_kf, n_lm):
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import builders as B
    from pyslam_b200 import synthetic
    return B.oracle_ba_arrays(synthetic.stereo_ba(n_kf, n_lm, track=TRACK, seed=0))


def oracle_step_seconds(n_kf, n_lm, steps=1, warmup=0):
    """Seconds per Gauss-Newton iteration of the CPU oracle (numpy assembly +
    scipy.sparse spsolve on the full system, i.e. the reference's algorithm,
    pyslam/problem.py:279-336,186) on an n_kf x n_lm problem."""
    from oracle import gn_oracle as O
    ba = oracle_problem(n_kf, n_lm)
    for _ in range(warmup):
        O.ba_iteration(ba)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.ba_iteration(ba)
    return (time.perf_counter() - t0) / steps


def pick_sample(budget_s, n_steps):
    t_small = oracle_step_seconds(*SAMPLES[-1])
    # measured in the build container: 0.25 s, 2.1 s, 13.2 s per iteration for the three samples
    rel = {SAMPLES[0]: 13.2 / 0.25, SAMPLES[1]: 2.07 / 0.25, SAMPLES[2]: 1.0}
    for s in SAMPLES:
        if t_small * rel[s] * n_steps <= budget_s:
            return s
    return SAMPLES[-1]


def cpu_baseline(budget_s):
    n_kf, n_lm = pick_sample(budget_s, 1)
    t = oracle_step_seconds(n_kf, n_lm)
    scale = N_LM / n_lm
    import scipy
    full_s = None
    try:
        full_s = [float(x) for x in np.load(os.path.join(ROOT, 'tests', 'golden', 'c4_summary.npz'))['seconds']]
    except Exception:
        pass
    return {'value': 1.0 / (t * scale), 'unit': 'iterations/s', 'cores': 1, 'kind': 'port',
            'sample': 'BOUNDED SAMPLE: one Gauss-Newton iteration of the numpy/scipy oracle (vectorised assembly + SuperLU spsolve '
                      'on the full system, the reference algorithm) on a 1/%d-scale problem (%d kf x %d lm x %d obs, same '
                      'density): %.2f s; value = 1/(%d x that), a LINEAR extrapolation that is optimistic for the CPU (spsolve '
                      'scales ~n^2.6).  The full-size figure is what `bench.py --impl reference` measures; in the build container '
                      'the full-size oracle took %s s per iteration (tests/golden/c4_summary.npz).  The unmodified reference '
                      'needs 94 s/iteration at 1/20 scale and cannot run the full size (SURVEY F7).'
                      % (scale, n_kf, n_lm, n_lm * TRACK, t, scale, full_s),
            'host_cpus': os.cpu_count(), 'scipy': scipy.__version__}


def run_reference(args):
    """The reference's algorithm on the host cores at the FULL size of config 4: the numpy/scipy oracle port
    (the unmodified reference cannot hold this problem: its block grid needs 6e10 cells).  One
    iteration takes ~100-250 s, so the number of steps actually run is bounded by a wall-clock budget and
    PRINTED as `steps` (the requested K/W are echoed as `requested_*`)."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from oracle import gn_oracle as O
    budget_s = float(os.environ.get('BSLAM_REF_BUDGET_S', '170'))
    t_build = time.perf_counter()
    ba = oracle_problem(N_KF, N_LM)
    t_build = time.perf_counter() - t_build
    times, costs = [], []
    while True:
        t0 = time.perf_counter()
        r = O.ba_iteration(ba)
        times.append(time.perf_counter() - t0)
        costs.append(float(r['cost_new']))
        if len(times) >= max(1, args.steps) or sum(times) + times[-1] > budget_s:
            break
    t = float(np.mean(times))
    value = 1.0 / t
    sample = ('FULL SIZE, same config as the GPU arm: %d Gauss-Newton iteration(s) of the numpy/scipy oracle port of pyslam '
              '(vectorised assembly + scipy SuperLU spsolve on the full 302 994-dim system = pyslam/problem.py:182-194,279-336) '
              'on %d kf x %d lm x %d obs from the same initial guess; measured %s s per iteration; no warm-up, no extrapolation'
              % (len(times), N_KF, N_LM, N_LM * TRACK, ['%.1f' % x for x in times]))
    out = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'iterations/s', 'n_gpus': args.gpus,
           'steps': len(times), 'warmup': 0, 'requested_steps': args.steps, 'requested_warmup': args.warmup,
           'ms_per_step': t * 1e3, 'higher_is_better': True,
           'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
           'config': {'workload': 'stereo BA config 4: %d keyframes x %d landmarks x %d reprojections, track %d, '
                                  'Huber(1.5), pose 0 constant, Gauss-Newton (lambda=0)' % (N_KF, N_LM, N_LM * TRACK, TRACK),
                      'sample': sample, 'problem_build_s': t_build},
           'cost_after_each_step': costs,
           'cpu_baseline': {'value': value, 'unit': 'iterations/s', 'cores': 1, 'kind': 'port', 'sample': sample,
                            'host_cpus': os.cpu_count()},
           'e2e': {'value': value, 'unit': 'iterations/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(out))


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--rounds', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    a = ap.parse_args()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_ours(a)

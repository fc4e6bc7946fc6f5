"""Iterations/s of BASELINE configs 2, 3 and 5 on one GPU (device-resident, CUDA events on the library's stream).
Config 4 is bench.py's workload; these are the other GPU configs of SURVEY 8(d)."""
import sys, json
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import torch
import builders as B
from pyslam_b200 import synthetic

def time_engine(eng, K=50, warm=5):
    eng.snapshot()
    for _ in range(warm):
        eng.iterate(0., True)
    eng.restore()
    st = eng.torch_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(st)
    for _ in range(K):
        eng.iterate(0., True)
    e1.record(st); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    eng.restore()
    eng.enable_timing(True)
    eng.iterate(0., True); eng.iterate(0., True)
    t = {k: round(1e3 * v, 1) for k, v in eng.timings().items() if v > 0}
    eng.enable_timing(False)
    return {'iterations_per_s': round(1e3 / ms, 1), 'us_per_iteration': round(1e3 * ms, 1), 'phase_us': t}

out = {}
p = B.product_pose_graph(synthetic.se2_pose_graph(1000, 100, seed=0), 'se2'); p._ensure_lowered()
out['C2 SE(2) pose graph, 1000 poses, 1100 factors'] = time_engine(p._engine)
p = B.product_ba_problem(synthetic.stereo_ba(50, 5000, seed=0), bulk=True); p._ensure_lowered()
out['C3 stereo BA 50 x 5000 x 30000, Huber'] = time_engine(p._engine)
d = synthetic.photometric_pair(640, 480, seed=0); d['loss_k'] = d['loss'][1]
p, res = B.product_photometric_problem(d); p._ensure_lowered()
r = time_engine(p._engine)
n_px = len(res.im_ref)
r['pixels'] = n_px
r['linearize_GBps_at_48B_per_pixel'] = round(48.0 * n_px / (r['phase_us'].get('linearize', 1e9) * 1e-6) / 1e9, 1)
out['C5 dense photometric 640x480, Cauchy'] = r
print(json.dumps(out, indent=1))

"""Per-phase CUDA-event timings (un-graphed) and graphed iterations/s of BASELINE config 4."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
import bench
from pyslam_b200 import synthetic
d = synthetic.stereo_ba(500, 100000, track=6, seed=0)
eng, Rt0 = bench.build_engine(d, 0)
for _ in range(3):
    eng.iterate(0., True)
eng.set_poses_se3(Rt0); eng.set_points(d['pts0'])
import torch
st = eng.torch_stream()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record(st)
K = 20
for _ in range(K):
    c = eng.iterate(0., True)
e1.record(st); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
eng.set_poses_se3(Rt0); eng.set_points(d['pts0'])
eng.enable_timing(True)
acc = {}
for _ in range(K):
    eng.iterate(0., True)
    for k, v in eng.timings().items():
        acc[k] = acc.get(k, 0.) + v / K
print('it/s %.1f  us/iter %.1f  final cost %.6e | ' % (1e3 / ms, 1e3 * ms, c[1]) + ' '.join('%s %.1f' % (k, 1e3 * v) for k, v in acc.items() if v > 0))

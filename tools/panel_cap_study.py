"""Fused panel kernel and finish kernel times of ONE landmark shard (rank 0 of N) of BASELINE config 4 with panels of
at most 32 or 64 landmarks (BSLAM_PANEL_CAP): which capacity suits which shard size.  Single GPU."""
import os, sys
sys.path.insert(0, '.')
import numpy as np
from pyslam_b200 import configs, synthetic
from pyslam_b200.dist import shard_stereo_ba
full = synthetic.stereo_ba(500, 100000, track=6, seed=0)
for world in (1, 2, 4, 8):
    d = shard_stereo_ba(full, 0, world) if world > 1 else full
    for cap in (32, 64):
        os.environ['BSLAM_PANEL_CAP'] = str(cap)
        eng, Rt0 = configs.ba_engine(d, 0)
        eng.finalize()
        for _ in range(3):
            eng.iterate(0., True)
        eng.enable_timing(True)
        acc = {}
        for _ in range(10):
            eng.set_poses_se3(Rt0); eng.set_points(d['pts0'])
            eng.iterate(0., True)
            for k, v in eng.timings().items():
                acc[k] = acc.get(k, 0.) + 100. * v
        print('shard 1/%d  cap %d  panels %5d  fused %6.1f us  finish %6.1f us  cholesky %6.1f  total %6.1f' % (world, cap, eng.fused_info()[0], acc['fused'], acc['cost'], acc['cholesky'], acc['total']), flush=True)
        eng.close()

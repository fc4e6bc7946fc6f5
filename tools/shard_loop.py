"""A few un-graphed iterations of ONE landmark shard (rank 0 of N, default 8) of BASELINE config 4: ncu target for the
latency floors of the panel kernels on small grids."""
import os, sys
sys.path.insert(0, '.')
from pyslam_b200 import configs, synthetic
from pyslam_b200.dist import shard_stereo_ba
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
full = synthetic.stereo_ba(500, 100000, track=6, seed=0)
d = shard_stereo_ba(full, 0, world) if world > 1 else full
eng, Rt0 = configs.ba_engine(d, 0)
eng.finalize()
for _ in range(4):
    print(eng.iterate(0., True))

"""Run a few un-graphed bslam_iterate calls of BASELINE config 4 (profiling target for ncu; BSLAM_NO_GRAPH=1)."""
import sys
sys.path.insert(0, '.')
import bench
from pyslam_b200 import synthetic
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
d = synthetic.stereo_ba(500, 100000, track=6, seed=0)
eng, _ = bench.build_engine(d, 0)
print('panels, fused landmarks:', eng.fused_info())
for _ in range(n):
    print(eng.iterate(0., True))
print('done', eng.launch_count())

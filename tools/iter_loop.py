"""Run a few un-graphed iterations of BASELINE config 4 (profiling target for ncu)."""
import sys
sys.path.insert(0, '.')
import bench
from pyslam_b200 import synthetic
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5
d = synthetic.stereo_ba(500, 100000, track=6, seed=0)
eng, _ = bench.build_engine(d, 0)
for _ in range(n):
    eng.linearize(fetch_cost=False); eng.reduce(0.); eng.solve_reduced(); eng.retract(True); eng.scalars()
print('done', eng.launch_count())

"""Lowering time of the drop-in Problem API on C4 with a heap that already holds other large problems (as in bench.py)."""
import sys, time
sys.path.insert(0, '.')
from pyslam_b200 import configs, synthetic
full = synthetic.stereo_ba(500, 100000, track=6, seed=0)
keep = [configs.ba_problem(full) for _ in range(2)]          # 2 x (100 500 parameter objects + 1.2 M keys) alive
keep[0].solve()
for rep in range(3):
    pr = configs.ba_problem(full)
    t0 = time.perf_counter()
    pr._ensure_lowered()
    t1 = time.perf_counter()
    pr.solve()
    t2 = time.perf_counter()
    print('lower %.3f s   solve (already lowered) %.3f s' % (t1 - t0, t2 - t1), flush=True)
    keep.append(pr)

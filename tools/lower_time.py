"""One-time lowering cost of the drop-in Problem API on C4 (Python key handling + bslam_finalize stages).
BSLAM_FINALIZE_TIMING=1 python tools/lower_time.py    (second lowering = warm process: module loaded, allocator primed)"""
import sys, time, cProfile, pstats
sys.path.insert(0, '.')
from pyslam_b200 import configs, synthetic
full = synthetic.stereo_ba(500, 100000, track=6, seed=0)
for rep in range(2):
    pr = configs.ba_problem(full)
    t0 = time.perf_counter()
    sys.stderr.write('--- lowering %d\n' % rep)
    cProfile.run('pr._ensure_lowered()', '/tmp/lower.prof')
    print('lower %d total %.3f s' % (rep, time.perf_counter() - t0), flush=True)
pstats.Stats('/tmp/lower.prof').sort_stats('tottime').print_stats(12)

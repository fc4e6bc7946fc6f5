"""Print the task timeline of one traced run of chol_solve_kernel on BASELINE config 4
(or --dense for a fully dense reduced matrix of the same size)."""
import sys
import numpy as np
sys.path.insert(0, '.')
import bench
from pyslam_b200 import synthetic

d = synthetic.stereo_ba(500, 100000, track=6, seed=0)
eng, _ = bench.build_engine(d, 0)
if '--dense' in sys.argv:
    m = eng.tile_structure()
    eng.merge_tile_structure(np.ones_like(m))
for _ in range(2):
    eng.linearize(fetch_cost=False); eng.reduce(0.); eng.solve_reduced(); eng.scalars()
eng.linearize(fetch_cost=False); eng.reduce(0.)
tr = eng.chol_trace()
t0 = tr[:, 2].min()
tr[:, 2:5] -= t0
print('tasks', len(tr), 'span us %.1f' % (tr[:, 4].max() / 1e3), 'SMs used', len(set(tr[:, 5])))
dur = (tr[:, 4] - tr[:, 2]) / 1e3
wait = (tr[:, 3] - tr[:, 2]) / 1e3
work = (tr[:, 4] - tr[:, 3]) / 1e3
kind = np.where(tr[:, 0] < 0, 'bwd', np.where(tr[:, 0] == tr[:, 1], 'diag', np.where(tr[:, 0] == tr[:, 0].max(), 'rhs', 'off')))
for k in ('diag', 'off', 'rhs', 'bwd'):
    s = kind == k
    if s.any():
        print('%-5s n=%4d  total(us): mean %.1f max %.1f | until deps ready: mean %.1f | after deps: mean %.2f max %.2f'
              % (k, s.sum(), dur[s].mean(), dur[s].max(), wait[s].mean(), work[s].mean(), work[s].max()))
order = np.argsort(tr[:, 4])
print('last 25 tasks to finish: (i, j) start deps end [us]')
for t in order[-25:]:
    print('  (%3d,%3d) %8.1f %8.1f %8.1f  sm %d' % (tr[t, 0], tr[t, 1], tr[t, 2] / 1e3, tr[t, 3] / 1e3, tr[t, 4] / 1e3, tr[t, 5]))
if '--all' in sys.argv:
    for t in np.argsort(tr[:, 2]):
        print('  (%3d,%3d) %8.1f %8.1f %8.1f' % (tr[t, 0], tr[t, 1], tr[t, 2] / 1e3, tr[t, 3] / 1e3, tr[t, 4] / 1e3))

"""Critical chain of one traced chol_solve_kernel run on BASELINE config 4 (or config 2 with the argument c2): the diagonal tasks in the order they
finish, with the time since the previous link, plus the shape of the plan (tiles, levels)."""
import sys
import numpy as np
sys.path.insert(0, '.')
import bench
from pyslam_b200 import synthetic

if 'c2' in sys.argv:      # BASELINE config 2: SE(2) pose graph, 1000 poses, 100 loop closures
    from pyslam_b200 import configs
    pr = configs.pose_graph_problem(synthetic.se2_pose_graph(1000, 100, seed=0), 'se2')
    pr._ensure_lowered()
    eng = pr._engine
else:
    d = synthetic.stereo_ba(500, 100000, track=6, seed=0)
    eng, _ = bench.build_engine(d, 0)
for _ in range(2):
    eng.linearize(fetch_cost=False); eng.reduce(0.); eng.solve_reduced(); eng.scalars()
eng.linearize(fetch_cost=False); eng.reduce(0.)
tr = eng.chol_trace().astype(np.int64)
t0 = tr[:, 2].min()
tr[:, 2:5] -= t0
nt = int(tr[:, 1].max()) + 1
print('tasks', len(tr), 'nt', nt, 'span us %.1f' % (tr[:, 4].max() / 1e3))
end = {(int(r[0]), int(r[1])): r[4] / 1e3 for r in tr if r[0] >= 0}
dep = {(int(r[0]), int(r[1])): r[3] / 1e3 for r in tr if r[0] >= 0}
start = {(int(r[0]), int(r[1])): r[2] / 1e3 for r in tr if r[0] >= 0}
m = eng.fill_structure() if hasattr(eng, 'fill_structure') else None
# walk back from the last diagonal task: its latest-finishing producer among the tiles of its row
tiles = sorted(end)
rows = {}
for (i, j) in tiles:
    rows.setdefault(i, []).append(j)
last_diag = max((k for k in tiles if k[0] == k[1]), key=lambda k: end[k])
chain = [last_diag]
cur = last_diag
while True:
    i = cur[0]
    prods = [(i, j) for j in rows[i] if j < i]
    if not prods:
        break
    crit = max(prods, key=lambda k: end[k])      # the off-diagonal tile (i, j) that arrived last
    j = crit[1]
    chain.append(crit)
    chain.append((j, j))
    cur = (j, j)
chain.reverse()
prev = 0.
print('critical chain (tile: start, deps ready, end, +since previous link) [us]')
for k in chain:
    print('  (%3d,%3d) %7.1f %7.1f %7.1f  +%.1f' % (k[0], k[1], start[k], dep[k], end[k], end[k] - prev))
    prev = end[k]
bw = tr[tr[:, 0] < 0]
print('backward: first start %.1f, last end %.1f, n=%d' % (bw[:, 2].min() / 1e3, bw[:, 4].max() / 1e3, len(bw)))
print('diagonal tasks: n=%d, mean (end - deps ready) %.2f us' % (sum(1 for k in tiles if k[0] == k[1]),
      np.mean([end[k] - dep[k] for k in tiles if k[0] == k[1]])))

#!/bin/bash
# compute-sanitizer over a small stereo BA: the fused iteration (TMA-staged panel kernel, Cholesky with fused retraction,
# finish kernel), the step-wise materialised-W path, and a 2-shard peer iteration (handles of one process) -- memcheck and racecheck
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import bench
from pyslam_b200 import configs, synthetic
from pyslam_b200.dist import connect_local, covisibility_pairs, iterate_local, shard_stereo_ba
d = synthetic.stereo_ba(40, 3000, track=6, seed=3)
eng, _ = bench.build_engine(d, 0)
for _ in range(2):
    print('fused', eng.iterate(0., True))
for _ in range(2):
    eng.linearize(fetch_cost=False); eng.reduce(0.); eng.solve_reduced(); eng.retract(True); print('stepwise', eng.scalars()[:3])
pairs = covisibility_pairs(d['pose_idx'], d['pt_idx'])
engines = []
for r in range(2):
    e, _ = configs.ba_engine(shard_stereo_ba(d, r, 2, by_time=True), 0)
    e.add_coupling(3, pairs[:, 0], pairs[:, 1]); e.finalize(); engines.append(e)
solvers = connect_local(engines)
for _ in range(2):
    print('sharded', iterate_local(solvers, 0., True)[0])
PY
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san.py > gpurun_out/san_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|=========     at" gpurun_out/san_$tool.log | head -12; grep -E "^(fused|stepwise|sharded)" gpurun_out/san_$tool.log | tail -3
done

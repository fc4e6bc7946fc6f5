#!/bin/bash
# compute-sanitizer over a small stereo BA (all block kernels + Schur + Cholesky + finish) -- memcheck and racecheck
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import bench
from pyslam_b200 import synthetic
d = synthetic.stereo_ba(40, 3000, track=6, seed=3)
eng, _ = bench.build_engine(d, 0)
for _ in range(2):
    eng.linearize(fetch_cost=False); eng.reduce(0.); eng.solve_reduced(); eng.retract(True); print(eng.scalars()[:3])
PY
for tool in memcheck racecheck; do
  timeout 250 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san.py > gpurun_out/san_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|=========     at" gpurun_out/san_$tool.log | head -12; grep -E "^\[" gpurun_out/san_$tool.log | tail -2
done

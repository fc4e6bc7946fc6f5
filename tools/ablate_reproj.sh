for f in 0 1 2 4 6 7; do BSLAM_DBG=$f python - <<'PY'
import os, sys, numpy as np
sys.path.insert(0,'.')
import bench
from pyslam_b200 import synthetic
d = synthetic.stereo_ba(500,100000,track=6,seed=0)
eng,_ = bench.build_engine(d,0)
eng.enable_timing(True)
ts=[]
for k in range(6):
    eng.linearize(fetch_cost=False); eng.reduce(0.); eng.solve_reduced(); eng.retract(True); eng.scalars()
    ts.append(eng.timings()['reproj'])
print('DBG', os.environ['BSLAM_DBG'], 'reproj us', np.round(1e3*np.array(ts[2:]),1))
PY
done

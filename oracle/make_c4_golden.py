#!/usr/bin/env python
"""Golden summary of the BENCHED configuration (BASELINE config 4: 500 keyframes x
100 000 landmarks x 600 000 reprojections, Huber(1.5), the exact problem bench.py
builds with `synthetic.stereo_ba(500, 100000, track=6, seed=0)`), produced by the
CPU oracle (oracle/gn_oracle.py: vectorised assembly + scipy SuperLU `spsolve` on
the FULL 302 994-dimensional system, i.e. the reference's algorithm,
pyslam/problem.py:182-194,279-336).  The unmodified reference cannot run this size
(its block grid needs 6e10 cells); the oracle is pinned to the
reference on the smaller fixtures of make_golden.py.

TEST INFRASTRUCTURE ONLY.  Takes ~4 CPU-minutes per iteration and ~20 GB of RAM.

    python oracle/make_c4_golden.py [n_iterations=2]   ->  tests/golden/c4_summary.npz

Stored per iteration: cost at the linearisation point, cost at x [+] dx, ||dx||,
the 2 994 pose entries of dx (reference order) and 4 096 sampled landmark entries.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

N_KF, N_LM, TRACK, SEED = 500, 100000, 6, 0


def main(n_iter=2):
    import builders as B
    from oracle import gn_oracle as O
    from pyslam_b200 import synthetic
    d = synthetic.stereo_ba(N_KF, N_LM, track=TRACK, seed=SEED)
    ba = B.oracle_ba_arrays(d)
    n_pose = 6 * (N_KF - 1)
    rng = np.random.default_rng(12345)
    sample = np.sort(rng.choice(3 * N_LM, size=4096, replace=False)) + n_pose
    out = dict(n_kf=N_KF, n_lm=N_LM, track=TRACK, seed=SEED, n_iter=n_iter, sample_idx=sample,
               cost_lin=[], cost_new=[], dx_norm=[], dx_pose=[], dx_sample=[], seconds=[])
    for it in range(n_iter):
        t0 = time.perf_counter()
        r = O.ba_iteration(ba)
        dt = time.perf_counter() - t0
        out['cost_lin'].append(r['cost_lin']); out['cost_new'].append(r['cost_new'])
        out['dx_norm'].append(float(np.linalg.norm(r['dx'])))
        out['dx_pose'].append(r['dx'][:n_pose].copy()); out['dx_sample'].append(r['dx'][sample].copy())
        out['seconds'].append(dt)
        print('iteration %d: cost %.9e -> %.9e, |dx| = %.6e, %.1f s' % (it, r['cost_lin'], r['cost_new'], out['dx_norm'][-1], dt),
              flush=True)
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'c4_summary.npz'),
                        **{k: np.asarray(v) for k, v in out.items()})


if __name__ == '__main__':
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 2)

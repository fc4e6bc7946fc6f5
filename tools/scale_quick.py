"""Quick strong-scaling probe of BASELINE config 4 (value only): run under torch.distributed.run with N ranks.
Prints one line per run: N, iterations/s (device-resident, L2 flushed per step, max over ranks), mode, parity of dx."""
import os, sys, json
import numpy as np
sys.path.insert(0, '.')
import torch, torch.distributed as dist
import bench
from pyslam_b200 import synthetic
from pyslam_b200.dist import build_sharded_ba

world = int(os.environ.get('WORLD_SIZE', '1')); rank = int(os.environ.get('RANK', '0')); local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
track = int(sys.argv[1]) if len(sys.argv) > 1 else 6
full = synthetic.stereo_ba(500, 100000, track=track, seed=0)
solver, d, Rt0 = build_sharded_ba(full, rank, world, local)
eng = solver.engine
tm = bench.Timer(torch, eng.torch_stream(), world)
for _ in range(5):
    solver.iterate(0., True)
eng.set_poses_se3(Rt0); eng.set_points(d['pts0'])
K = 50
align = eng.peer_barrier if solver.mode == 'peer' else None
ms = tm.device_ms(lambda: solver.iterate(0., True), K, align) / K
ms_unaligned = tm.device_ms(lambda: solver.iterate(0., True), K) / K
eng.enable_timing(True)
acc = {}
for _ in range(10):
    solver.iterate(0., True)
    for k, v in eng.timings().items():
        acc[k] = acc.get(k, 0.) + v / 10
eng.enable_timing(False)
par = None
if track == 6:
    par = bench.c4_parity(solver, eng, d, rank, world, lambda: (eng.set_poses_se3(Rt0), eng.set_points(d['pts0'])))
if rank == 0:
    print(json.dumps({'n_gpus': world, 'track': track, 'iterations_per_s': round(1e3 / ms, 1), 'us_per_iteration': round(1e3 * ms, 1),
                      'mode': solver.mode, 'transport': getattr(solver, 'transport', None), 'nvls_reason': getattr(solver, 'nvls_reason', None), 'panels': eng.fused_info(), 'us_unaligned': round(1e3 * ms_unaligned, 1), 'phases_us': {k: round(1e3 * v, 1) for k, v in acc.items() if v > 0}, 'parity_dx': None if par is None else par['dx_rel_err']}))
if world > 1:
    dist.destroy_process_group()

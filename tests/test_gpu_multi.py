"""GPU (>= 2 devices): landmark-sharded iteration over NCCL against the
single-GPU iteration on the same problem."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _build(d, device):
    sys.path.insert(0, os.path.dirname(HERE))
    import bench
    return bench.build_engine(d, device)[0]


def _worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    from pyslam_b200 import synthetic
    from pyslam_b200.dist import ShardedSolver, shard_stereo_ba
    full = synthetic.stereo_ba(40, 3000, track=6, seed=2)
    eng = _build(shard_stereo_ba(full, rank, world), rank)
    solver = ShardedSolver(eng, rank, world)
    res = [solver.eval_cost()]
    for _ in range(3):
        res.append(solver.iterate(0., True))
    out[rank] = (res, eng.get_poses_se3(), eng.get_points())
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_nccl_sharded_matches_single_gpu():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    from pyslam_b200 import synthetic
    from pyslam_b200.dist import shard_range
    world = 2
    out = mp.Manager().dict()
    mp.spawn(_worker, args=(world, 29600 + os.getpid() % 2000, out), nprocs=world, join=True)
    full = synthetic.stereo_ba(40, 3000, track=6, seed=2)
    eng = _build(full, 0)
    ref = [eng.eval_cost()] + [eng.iterate(0., True) for _ in range(3)]
    poses_ref, pts_ref = eng.get_poses_se3(), eng.get_points()
    for rank in range(world):
        res, poses, pts = out[rank]
        lo, hi = shard_range(3000, rank, world)
        assert abs(res[0] - ref[0]) < 1e-12 * ref[0]
        for a, b in zip(res[1:], ref[1:]):
            np.testing.assert_allclose(a, b, rtol=1e-8)
        np.testing.assert_allclose(poses, poses_ref, rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(pts, pts_ref[lo:hi], rtol=1e-9, atol=1e-11)

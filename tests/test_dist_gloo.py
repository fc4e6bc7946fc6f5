"""CPU, world_size 2, gloo: the landmark-sharded iteration of pyslam_b200.dist
(one all-reduce of [S | rhs | cost] + one of two scalars per iteration) must
reproduce the single-process iteration.  The engine is the oracle-backed test
double; the NCCL/GPU version of the same code path runs in bench.py --gpus N
and tests/test_gpu_multi.py."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, out):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from fake_engine import FakeEngine
    from pyslam_b200 import synthetic
    from pyslam_b200.dist import ShardedSolver, shard_stereo_ba, shard_range
    full = synthetic.stereo_ba(6, 41, track=4, seed=3)
    d = shard_stereo_ba(full, rank, world)
    lo, hi = shard_range(41, rank, world)
    assert d['lm_range'] == (lo, hi) and len(d['pts0']) == hi - lo
    eng = FakeEngine()
    eng.set_poses_se3(np.concatenate([d['R0'].reshape(-1, 9), d['t0']], axis=1), d['pose_const'])
    eng.set_points(d['pts0'])
    eng.add_reprojection_blocks(d['pose_idx'], d['pt_idx'], d['obs'], d['stiffness'], d['intr'], 3, 1.5)
    eng.finalize()
    solver = ShardedSolver(eng, rank, world)
    res = [solver.eval_cost()]
    for _ in range(3):
        res.append(solver.iterate(0., True))
    out[rank] = (res, eng.get_poses_se3(), eng.get_points(), lo, hi)
    dist.barrier()
    dist.destroy_process_group()


def _single():
    sys.path.insert(0, HERE)
    from fake_engine import FakeEngine
    from pyslam_b200 import synthetic
    d = synthetic.stereo_ba(6, 41, track=4, seed=3)
    eng = FakeEngine()
    eng.set_poses_se3(np.concatenate([d['R0'].reshape(-1, 9), d['t0']], axis=1), d['pose_const'])
    eng.set_points(d['pts0'])
    eng.add_reprojection_blocks(d['pose_idx'], d['pt_idx'], d['obs'], d['stiffness'], d['intr'], 3, 1.5)
    eng.finalize()
    res = [eng.eval_cost()]
    for _ in range(3):
        res.append(eng.iterate(0., True))
    return res, eng.get_poses_se3(), eng.get_points()


@pytest.mark.timeout(300)
def test_sharded_iteration_matches_single_process():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    ref, poses_ref, pts_ref = _single()
    for rank in range(world):
        res, poses, pts, lo, hi = out[rank]
        assert abs(res[0] - ref[0]) < 1e-10 * ref[0]
        for a, b in zip(res[1:], ref[1:]):
            np.testing.assert_allclose(a, b, rtol=1e-8)           # (cost_lin, cost_new, ||dx||) on every rank
        np.testing.assert_allclose(poses, poses_ref, rtol=1e-9, atol=1e-12)   # replicated poses stay identical
        np.testing.assert_allclose(pts, pts_ref[lo:hi], rtol=1e-9, atol=1e-12)


def test_shard_ranges_cover_everything():
    from pyslam_b200.dist import shard_range
    for n in (0, 1, 7, 100000):
        for world in (1, 2, 3, 8):
            r = [shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_time_contiguous_shards_and_covisibility_pairs():
    """shard_stereo_ba(by_time=True) partitions the landmarks in first-keyframe order (every landmark exactly once, local
    indices consistent with lm_ids); covisibility_pairs = the pose pairs that share a landmark."""
    from pyslam_b200 import synthetic
    from pyslam_b200.dist import covisibility_pairs, shard_stereo_ba
    full = synthetic.stereo_ba(30, 400, track=4, seed=3)
    world, seen = 3, []
    first = np.full(400, 1 << 30)
    np.minimum.at(first, full['pt_idx'], full['pose_idx'])
    prev_max = -1
    for r in range(world):
        d = shard_stereo_ba(full, r, world, by_time=True)
        seen.append(d['lm_ids'])
        assert np.array_equal(d['pts0'], full['pts0'][d['lm_ids']])
        # the observations of the shard are exactly those of its landmarks, renumbered
        keep = np.isin(full['pt_idx'], d['lm_ids'])
        assert np.array_equal(d['lm_ids'][d['pt_idx']], full['pt_idx'][keep])
        assert np.array_equal(d['pose_idx'], full['pose_idx'][keep])
        # time-contiguous: no landmark of this shard starts before one of the previous shard
        assert first[d['lm_ids']].min() >= prev_max
        prev_max = first[d['lm_ids']].max()
    assert np.array_equal(np.sort(np.concatenate(seen)), np.arange(400))
    pairs = covisibility_pairs(full['pose_idx'], full['pt_idx'])
    ref = set()
    for q in range(400):
        ps = sorted(set(full['pose_idx'][full['pt_idx'] == q]))
        ref.update((a, b) for i, a in enumerate(ps) for b in ps[i + 1:])
    assert set(map(tuple, pairs.tolist())) == ref

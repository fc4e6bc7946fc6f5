"""GPU: the landmark-sharded iteration (pyslam_b200/dist.py, csrc/peer.cuh) against the single-GPU iteration
on the same problem.

  * shards as several handles of ONE process on cuda:0 (runs on a 1-GPU lease): the real kernels of the
    sharded schedule -- pack + rendezvous, the Cholesky kernel reading sum_r S_r from the shards' exchange
    regions, the scalar mailbox -- with plain device pointers instead of CUDA-IPC mappings;
  * two PROCESSES on cuda:0 with gloo as the set-up plumbing and CUDA IPC for the mapping (1-GPU lease too);
  * >= 2 GPUs: one process per GPU over NCCL, peer mode and the torch.distributed fallback.
"""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def time_sorted(d):
    """Renumber the landmarks by the first keyframe that sees them (what a SLAM front-end produces): the shards
    then see disjoint parts of the trajectory and would order the reduced system differently on their own."""
    n = len(d['pts0'])
    first = np.full(n, 1 << 30)
    np.minimum.at(first, d['pt_idx'], d['pose_idx'])
    order = np.argsort(first, kind='stable')
    new_id = np.empty(n, np.int32)
    new_id[order] = np.arange(n, dtype=np.int32)
    out = dict(d)
    out.update(pts0=d['pts0'][order], pts_true=d['pts_true'][order], pt_idx=new_id[d['pt_idx']])
    return out


def _single(full, n_iter):
    from pyslam_b200 import configs
    eng, _ = configs.ba_engine(full, 0)
    eng.finalize()
    ref = [eng.eval_cost()] + [eng.iterate(0., True) for _ in range(n_iter)]
    out = ref, eng.get_poses_se3(), eng.get_points()
    eng.close()
    return out


def _local_shards(full, world, couple=True):
    from pyslam_b200 import configs
    from pyslam_b200.dist import covisibility_pairs, shard_stereo_ba
    pairs = covisibility_pairs(full['pose_idx'], full['pt_idx'])
    engines = []
    for r in range(world):
        eng, _ = configs.ba_engine(shard_stereo_ba(full, r, world), 0)
        if couple:
            eng.add_coupling(3, pairs[:, 0], pairs[:, 1])
        eng.finalize()
        engines.append(eng)
    return engines


@pytest.mark.timeout(600)
@pytest.mark.parametrize('world,sort_ids', [(2, False), (2, True), (3, True), (4, False)])
def test_local_shards_match_single_gpu(world, sort_ids):
    from pyslam_b200 import synthetic
    from pyslam_b200.dist import connect_local, iterate_local, shard_range
    full = synthetic.stereo_ba(40, 3000, track=6, seed=2)
    if sort_ids:
        full = time_sorted(full)
    n_iter = 3
    ref, poses_ref, pts_ref = _single(full, n_iter)
    engines = _local_shards(full, world)
    solvers = connect_local(engines)
    assert abs(sum(e.eval_cost() for e in engines) - ref[0]) < 1e-12 * ref[0]
    for it in range(n_iter):
        res = iterate_local(solvers, 0., True)
        for r in res:
            assert r == res[0]                                   # the summed scalars are bit-identical on every shard
            np.testing.assert_allclose(r, ref[1 + it], rtol=1e-8)
    poses0 = engines[0].get_poses_se3()
    for r, eng in enumerate(engines):
        lo, hi = shard_range(len(full['pts0']), r, world)
        assert np.array_equal(eng.get_poses_se3(), poses0)       # replicated solve: bit-identical poses
        np.testing.assert_allclose(eng.get_poses_se3(), poses_ref, rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(eng.get_points(), pts_ref[lo:hi], rtol=1e-9, atol=1e-11)
    for eng in engines:
        eng.close()


def test_shards_without_declared_couplings_are_rejected():
    """Time-sorted landmark ids: each shard alone would pick its own ordering of the reduced system;
    connect_local / ShardedSolver compare bslam_layout_hash and refuse."""
    from pyslam_b200 import synthetic
    from pyslam_b200.dist import connect_local
    full = time_sorted(synthetic.stereo_ba(80, 2000, track=6, seed=3))
    engines = _local_shards(full, 2, couple=False)
    if engines[0].layout_hash() == engines[1].layout_hash():
        pytest.skip('the two shards happened to derive the same layout')
    with pytest.raises(RuntimeError):
        connect_local(engines)
    for eng in engines:
        eng.close()


def test_local_shards_with_unsharded_blocks_and_lm():
    """Pose prior + relative-pose factors next to the sharded landmarks are counted once (shard 0), lambda > 0."""
    from pyslam_b200 import configs, synthetic
    from pyslam_b200.dist import connect_local, covisibility_pairs, iterate_local, shard_stereo_ba
    full = synthetic.stereo_ba(30, 2000, track=5, seed=5)
    Rt_true = np.concatenate([full['R_true'].reshape(-1, 9), full['t_true']], axis=1)
    S6 = 3.0 * np.eye(6)

    def add_edges(eng):
        eng.add_pose_blocks(3, [4], Rt_true[4:5], S6)
        i1, i2 = np.arange(0, 29, dtype=np.int32), np.arange(1, 30, dtype=np.int32)
        rel = []
        for a, b in zip(i1, i2):
            Ra, ta = full['R_true'][a], full['t_true'][a]
            Rb, tb = full['R_true'][b], full['t_true'][b]
            R = Rb @ Ra.T
            rel.append(np.concatenate([R.ravel(), tb - R @ ta]))
        eng.add_pose_to_pose_blocks(3, i1, i2, np.array(rel), S6)

    eng1, _ = configs.ba_engine(full, 0)
    add_edges(eng1)
    eng1.finalize()
    ref = [eng1.iterate(1e-3, True) for _ in range(2)]
    pairs = covisibility_pairs(full['pose_idx'], full['pt_idx'])
    engines = []
    for r in range(2):
        eng, _ = configs.ba_engine(shard_stereo_ba(full, r, 2), 0)
        add_edges(eng)                      # registered on every rank, assembled by shard 0 only
        eng.add_coupling(3, pairs[:, 0], pairs[:, 1])
        eng.finalize()
        engines.append(eng)
    solvers = connect_local(engines)
    for it in range(2):
        for r in iterate_local(solvers, 1e-3, True):
            np.testing.assert_allclose(r, ref[it], rtol=1e-8)
    np.testing.assert_allclose(engines[1].get_poses_se3(), eng1.get_poses_se3(), rtol=1e-9, atol=1e-11)
    for eng in engines + [eng1]:
        eng.close()


# ---------------------------------------------------------------- one process per rank
def _worker(rank, world, port, out, backend, same_device, mode):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dev = 0 if same_device else rank
    torch.cuda.set_device(dev)
    if backend == 'nccl':
        dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', dev))
    else:
        dist.init_process_group('gloo', rank=rank, world_size=world)
    from pyslam_b200 import synthetic
    from pyslam_b200.dist import build_sharded_ba
    full = time_sorted(synthetic.stereo_ba(40, 3000, track=6, seed=2))
    solver, d, _ = build_sharded_ba(full, rank, world, dev, mode=mode)
    eng = solver.engine
    res = [solver.eval_cost()]
    for _ in range(3):
        res.append(solver.iterate(0., True))
    out[rank] = (res, eng.get_poses_se3(), eng.get_points(), solver.mode, np.asarray(d['lm_ids']), getattr(solver, 'tiles_per_rank', None))
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()


def _run_processes(world, backend, same_device, mode):
    import torch.multiprocessing as mp
    from pyslam_b200 import synthetic
    from pyslam_b200.dist import shard_range
    out = mp.Manager().dict()
    mp.spawn(_worker, args=(world, 29600 + os.getpid() % 2000, out, backend, same_device, mode), nprocs=world, join=True)
    full = time_sorted(synthetic.stereo_ba(40, 3000, track=6, seed=2))
    ref, poses_ref, pts_ref = _single(full, 3)
    for rank in range(world):
        res, poses, pts, used, lm_ids, tiles_per_rank = out[rank]
        if mode != 'auto':
            assert used == mode
        if used == 'peer':      # time-contiguous shards: no rank contributes to every tile of the reduced system
            assert tiles_per_rank is not None and min(tiles_per_rank) < max(tiles_per_rank) + 1 and len(tiles_per_rank) == world
        assert abs(res[0] - ref[0]) < 1e-12 * ref[0]
        for a, b in zip(res[1:], ref[1:]):
            np.testing.assert_allclose(a, b, rtol=1e-8)
        np.testing.assert_allclose(poses, poses_ref, rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(pts, pts_ref[lm_ids], rtol=1e-9, atol=1e-11)
    return out


@pytest.mark.timeout(600)
def test_two_processes_one_gpu_ipc():
    """Two ranks = two processes sharing cuda:0: gloo for the set-up, CUDA IPC for the exchange regions."""
    _run_processes(2, 'gloo', True, 'peer')


@pytest.mark.timeout(600)
@pytest.mark.parametrize('mode', ['peer', 'nccl'])
def test_nccl_sharded_matches_single_gpu(mode):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    _run_processes(2, 'nccl', False, mode)


@pytest.mark.timeout(600)
def test_local_shards_general_path_and_losses():
    """Shards whose landmarks do NOT form panels (tracks of 12 poses: materialised-W block kernels, separate retraction
    kernel) under a Cauchy loss and lambda > 0, time-contiguous shards with contributor masks."""
    from pyslam_b200 import configs, synthetic
    from pyslam_b200.dist import connect_local, covisibility_pairs, iterate_local, shard_stereo_ba
    full = synthetic.stereo_ba(30, 1500, track=12, seed=7)
    full['loss'] = ('cauchy', 2.0)
    eng1, _ = configs.ba_engine(full, 0)
    eng1.finalize()
    assert eng1.fused_info()[0] == 0
    ref = [eng1.iterate(1e-2, True) for _ in range(3)]
    pairs = covisibility_pairs(full['pose_idx'], full['pt_idx'])
    engines, shards = [], []
    for r in range(3):
        d = shard_stereo_ba(full, r, 3, by_time=True)
        e, _ = configs.ba_engine(d, 0)
        e.add_coupling(3, pairs[:, 0], pairs[:, 1])
        e.finalize()
        engines.append(e); shards.append(d)
    solvers = connect_local(engines)
    for it in range(3):
        for r in iterate_local(solvers, 1e-2, True):
            np.testing.assert_allclose(r, ref[it], rtol=1e-8)
    pts_ref = eng1.get_points()
    for e, d in zip(engines, shards):
        np.testing.assert_allclose(e.get_poses_se3(), eng1.get_poses_se3(), rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(e.get_points(), pts_ref[d['lm_ids']], rtol=1e-9, atol=1e-11)
    for e in engines + [eng1]:
        e.close()

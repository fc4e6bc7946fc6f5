"""Landmark-sharded bundle adjustment over several GPUs.

One process per GPU; `torch.distributed` (NCCL or gloo) is the set-up plumbing only.  Landmarks -- and the
observations that reference them -- are partitioned contiguously over the ranks; the poses and the reduced
camera system are replicated.

Default schedule (`mode='peer'`, pyslam_b200/csrc/peer.cuh): ONE CUDA graph per rank and iteration, no host code
and no library collective inside it:

    every rank : linearise its observations, eliminate its landmarks  -> partial [S | rhs]
                 gather the structurally non-zero tiles into the rank's exchange region; rendezvous (flags
                 written into the peers' regions over NVLink)
    every rank : tile Cholesky of  sum_r S_r : each tile task reads its operand from ALL ranks' regions
                 (mapped peer memory, fixed rank order => bit-identical on every rank) -- the all-reduce is
                 fused into the factorisation kernel's loads; solve dx_c, back-substitute and retract the
                 rank's own landmarks, cost at the new point
    every rank : partial scalars (cost, new cost, ||dx_p||^2) to every peer's mailbox, rendezvous, sum

Set-up (once): the ranks exchange their pose co-visibility pairs so that every rank derives the SAME reduced
ordering / tile structure from the full coupling graph (`bslam_add_coupling`; checked with `bslam_layout_hash`),
then their CUDA-IPC handles of the exchange regions (`bslam_peer_region` / `bslam_peer_connect`).

Fallback schedule (`mode='nccl'`): two graph replays around a `torch.distributed` all-reduce of the packed
tiles (`bslam_iterate_pre` / `bslam_iterate_post`), used when peer mapping is unavailable and by the CPU test
double of tests/test_dist_gloo.py.

The reference has no distributed path at all; this is the B200-native replacement for the
single-process `spsolve` on the full system.
"""
import os

import numpy as np


def shard_range(n, rank, world):
    """Contiguous block partition of range(n)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_stereo_ba(d, rank, world, by_time=False):
    """Sub-problem of rank `rank`: all poses, its share of the landmarks and their observations (point indices
    renumbered from 0).  `lm_ids` = the global landmark index of every local landmark.

    by_time=False: landmarks [lo, hi) by index.  by_time=True: the landmarks are first ordered by the first keyframe
    that observes them and THEN cut into contiguous ranges: every rank's landmarks then couple ~1/world of the
    trajectory, so its partial reduced system touches ~1/world of the tiles and the fused all-reduce reads a tile only
    from the one or two ranks that can have written it (bslam_peer_set_contributors)."""
    n = len(d['pts0'])
    lo, hi = shard_range(n, rank, world)
    if by_time:
        first = np.full(n, np.iinfo(np.int64).max)
        np.minimum.at(first, d['pt_idx'], d['pose_idx'].astype(np.int64))
        order = np.argsort(first, kind='stable')
    else:
        order = np.arange(n)
    lm_ids = order[lo:hi]
    local_of = np.full(n, -1, np.int64)
    local_of[lm_ids] = np.arange(hi - lo)
    keep = local_of[d['pt_idx']] >= 0
    out = dict(d)
    out.update(pts0=d['pts0'][lm_ids], pts_true=d['pts_true'][lm_ids], pose_idx=d['pose_idx'][keep],
               pt_idx=local_of[d['pt_idx'][keep]].astype(np.int32), obs=d['obs'][keep], n_lm=hi - lo, lm_range=(lo, hi),
               lm_ids=lm_ids)
    return out


def covisibility_pairs(pose_idx, pt_idx):
    """Unique unordered pairs (a < b) of poses that observe a common landmark: the couplings the Schur complement
    of these observations creates in the reduced camera system.  [n_pairs, 2] int32."""
    pose_idx, pt_idx = np.asarray(pose_idx, np.int64), np.asarray(pt_idx, np.int64)
    if len(pose_idx) == 0:
        return np.zeros((0, 2), np.int32)
    order = np.argsort(pt_idx, kind='stable')
    p, q = pose_idx[order], pt_idx[order]
    n_pose = int(p.max()) + 1
    longest = int(np.bincount(q - q.min()).max())
    codes = []
    for dlt in range(1, longest):
        same = q[dlt:] == q[:-dlt]
        a, b = p[:-dlt][same], p[dlt:][same]
        lo, hi = np.minimum(a, b), np.maximum(a, b)
        codes.append(np.unique(lo[lo != hi] * n_pose + hi[lo != hi]))
    if not codes:
        return np.zeros((0, 2), np.int32)
    code = np.unique(np.concatenate(codes))
    return np.stack([code // n_pose, code % n_pose], axis=1).astype(np.int32)


def _all_gather_object(obj, group=None):
    import torch.distributed as dist
    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, obj, group=group)
    return out


def build_sharded_ba(full, rank=0, world=1, device=0, group=None, mode='auto'):
    """Shard a stereo-BA problem (dict of `pyslam_b200.synthetic.stereo_ba` form) over `world` ranks and lower
    this rank's part: (ShardedSolver, this rank's sub-problem, initial pose table)."""
    from . import configs
    d = shard_stereo_ba(full, rank, world, by_time=True) if world > 1 else full
    eng, Rt0 = configs.ba_engine(d, device)
    if world > 1:
        # every rank orders the reduced system from the couplings of ALL shards
        pairs = np.unique(np.concatenate(_all_gather_object(covisibility_pairs(d['pose_idx'], d['pt_idx']), group)), axis=0)
        eng.add_coupling(3, pairs[:, 0], pairs[:, 1])
    eng.finalize()
    return ShardedSolver(eng, rank, world, group, mode), d, Rt0


def _set_contributors(engines_or_engine, all_flags):
    """masks[k] = OR_r (flags_r[k] != 0) << r, handed to every engine."""
    masks = np.zeros(len(all_flags[0]), np.uint8)
    for r, f in enumerate(all_flags):
        masks |= (np.asarray(f, np.uint8) != 0).astype(np.uint8) << r
    for e in (engines_or_engine if isinstance(engines_or_engine, (list, tuple)) else [engines_or_engine]):
        e.peer_set_contributors(masks)
    return masks


def connect_local(engines):
    """Shards held by several handles of ONE process (tests, single-GPU studies): map the exchange regions by
    plain device pointers.  Returns the ShardedSolvers; drive them with `iterate_local`."""
    world = len(engines)
    hashes = {e.layout_hash() for e in engines}
    if len(hashes) != 1:
        raise RuntimeError('the shards derived different reduced layouts: declare the couplings of all shards '
                           '(add_coupling) before finalize')
    ptrs = [e.peer_region()[0] for e in engines]
    solvers = []
    for r, e in enumerate(engines):
        e.peer_connect(world, r, dev_ptrs=ptrs)
        solvers.append(ShardedSolver(e, r, world, mode='connected'))
    _set_contributors(engines, [e.peer_local_slots() for e in engines])
    return solvers


def iterate_local(solvers, lam=0., eval_new_cost=True):
    """One sharded iteration of handles that live in this process: enqueue all, then wait for all."""
    for s in solvers:
        s.engine.iterate_async(lam, eval_new_cost)
    return [s.engine.iterate_wait() for s in solvers]


class ShardedSolver:
    """Drives one engine per rank through the sharded iteration."""

    def __init__(self, engine, rank=0, world=1, group=None, mode='auto'):
        self.engine, self.rank, self.world, self.group = engine, rank, world, group
        self.mode = 'single' if world == 1 else mode
        if world == 1:
            engine.set_shard(0)
            return
        if mode == 'connected':          # connect_local did the mapping
            self.mode = 'peer'
            return
        import torch.distributed as dist
        if mode == 'auto':
            mode = 'peer' if (dist.get_backend(group) == 'nccl' and hasattr(engine, 'peer_region')) else 'nccl'
        engine.set_shard(rank)
        if hasattr(engine, 'layout_hash'):
            hashes = _all_gather_object(engine.layout_hash(), group)
            if len(set(hashes)) != 1:
                raise RuntimeError('ranks derived different reduced-system layouts (%s): every rank must declare the pose '
                                   'couplings of all shards before finalize (build_sharded_ba does)' % hashes)
        if mode == 'peer':
            # 1) symmetric memory with an NVLS multicast mapping (in-switch reduction), when the box offers it
            # (opt-in: measured on 2, 4 and 8 B200s the per-element 8-byte multimem loads are slower than 16-byte loads from
            #  the contributing ranks)
            if os.environ.get('BSLAM_NVLS', '0') == '1' and dist.get_backend(group) == 'nccl' and self._connect_symmetric(group):
                self.mode = 'peer'
                self._exchange_contributors(group)
                return
            # 2) CUDA-IPC mappings of cudaMalloc regions
            # every collective below is entered by every rank whatever failed locally: the ranks fall back together
            ok, err, handle = 1, '', b''
            try:
                handle = engine.peer_region()[2].tobytes()
            except Exception as e:
                ok, err = 0, repr(e)
            got = _all_gather_object((ok, err, handle), group)
            if all(g[0] for g in got):
                try:
                    engine.peer_connect(world, rank, ipc_handles=np.frombuffer(b''.join(g[2] for g in got), np.uint8))
                except Exception as e:      # e.g. no peer access between the devices
                    ok, err = 0, repr(e)
                got = _all_gather_object((ok, err, b''), group)
            if all(g[0] for g in got):
                self.mode = 'peer'
                self._exchange_contributors(group)
                return
            engine.peer_connect(1, 0)
            engine.set_shard(rank)
            self.fallback_reason = [g[1] for g in got if not g[0]][0]
        self.mode = 'nccl'
        self._merge_structure()

    def _exchange_contributors(self, group):
        """Every rank learns which ranks can write which tile of the reduced system (read only those)."""
        flags = _all_gather_object(self.engine.peer_local_slots().tobytes(), group)
        masks = _set_contributors(self.engine, [np.frombuffer(f, np.uint8) for f in flags])
        self.tiles_per_rank = [int(((masks >> r) & 1).sum()) for r in range(self.world)]

    def _connect_symmetric(self, group):
        """Exchange regions in torch symmetric memory; True when every rank got a multicast mapping and connected."""
        import torch
        import torch.distributed as dist
        eng = self.engine
        ok, err = 1, ''
        try:
            import torch.distributed._symmetric_memory as symm
            _, n_bytes, _ = eng.peer_region()
            t = symm.empty(n_bytes // 8, dtype=torch.float64, device='cuda:%d' % eng.device)
            t.zero_()
            torch.cuda.synchronize()
            hdl = symm.rendezvous(t, group if group is not None else dist.group.WORLD)
            mc = int(getattr(hdl, 'multicast_ptr', 0) or 0)
            if mc == 0:
                raise RuntimeError('no multicast mapping')
            hdl.barrier()
            eng.peer_connect_symmetric(self.world, self.rank, [int(p) for p in hdl.buffer_ptrs], mc, n_bytes)
            self._symm = (t, hdl)            # keep the allocation alive
        except Exception as e:
            ok, err = 0, repr(e)
        got = _all_gather_object((ok, err), group)
        if all(g[0] for g in got):
            self.transport = 'symmetric memory + NVLS multicast (multimem.ld_reduce)'
            return True
        if ok:                               # somebody else failed: back to a single-rank handle before the IPC attempt
            eng.peer_connect(1, 0)
            eng.set_shard(self.rank)
            self._symm = None
        self.nvls_reason = [g[1] for g in got if not g[0]][0]
        return False

    def describe(self):
        if self.world == 1:
            return 'single GPU'
        if self.mode == 'peer':
            return ('landmarks sharded over %d GPUs; per iteration ONE CUDA graph per rank: partial reduced systems published '
                    'to peer-mapped exchange regions (%s), all-reduce fused into the tile loads of the Cholesky kernel, '
                    'scalar exchange through peer mailboxes; no NCCL call inside the iteration'
                    % (self.world, getattr(self, 'transport', 'CUDA IPC over NVLink, one load per rank')))
        return ('landmarks sharded over %d GPUs, two CUDA graphs around one torch.distributed all-reduce of the packed '
                'reduced system per iteration (fallback: %s)' % (self.world, getattr(self, 'fallback_reason', 'requested')))

    def _merge_structure(self):
        """The summed reduced matrix has the UNION of the ranks' tile structures."""
        import torch
        import torch.distributed as dist
        m = self.engine.tile_structure()
        t = torch.from_numpy(m.astype(np.int32))
        if dist.get_backend(self.group) == 'nccl':
            t = t.cuda(self.engine.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        self.engine.merge_tile_structure(t.cpu().numpy().astype(np.uint8))

    def eval_cost(self):
        c = self.engine.eval_cost()
        if self.world == 1:
            return c
        import torch
        import torch.distributed as dist
        dev = 'cuda:%d' % self.engine.device if dist.get_backend(self.group) == 'nccl' else 'cpu'
        t = torch.tensor([c], dtype=torch.float64, device=dev)
        dist.all_reduce(t, group=self.group)
        return float(t.item())

    def iterate(self, lam=0., eval_new_cost=True):
        """(cost at the linearisation point, cost at x [+] dx, ||dx||), identical on every rank."""
        eng = self.engine
        if self.mode in ('single', 'peer'):
            return eng.iterate(lam, eval_new_cost)
        import torch.distributed as dist
        stream = eng.torch_stream()
        # two graph replays around the all-reduce of the packed non-zero tiles of S (+ rhs + scalars)
        eng.iterate_pre(lam)
        with _on_stream(stream):
            dist.all_reduce(eng.packed_tensor(), group=self.group)
        eng.iterate_post(eval_new_cost)
        with _on_stream(stream):
            dist.all_reduce(eng.scalars_tensor()[1:3], group=self.group)      # COST_NEW, DX_NORM2
        s = eng.scalars()
        return float(s[0]), float(s[1]), float(np.sqrt(s[2]))

    def iterate_host(self, Rt, xyz, lam=0., eval_new_cost=True):
        """One iteration on HOST (pinned) parameter tables of this rank: upload, iterate, download in place."""
        eng = self.engine
        if self.mode in ('single', 'peer'):
            return eng.iterate_host(Rt, xyz, lam, eval_new_cost)      # one C-ABI call, one synchronisation
        eng.set_poses_se3(Rt)
        eng.set_points(xyz)
        r = self.iterate(lam, eval_new_cost)
        eng.get_poses_se3(Rt)
        eng.get_points(xyz)
        return r


class _on_stream:
    """Make `stream` torch's current stream (no-op for CPU test doubles)."""

    def __init__(self, stream):
        self.stream, self.ctx = stream, None

    def __enter__(self):
        if self.stream is not None:
            import torch
            self.ctx = torch.cuda.stream(self.stream)
            self.ctx.__enter__()

    def __exit__(self, *a):
        if self.ctx is not None:
            self.ctx.__exit__(*a)

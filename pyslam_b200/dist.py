"""Landmark-sharded bundle adjustment over several GPUs.

One process per GPU (torch.distributed, NCCL over NVLink).  Landmarks -- and
the observations that reference them -- are partitioned contiguously over the
ranks; the poses and the reduced camera system are replicated.  Per iteration:

    every rank : linearise its observations, eliminate its landmarks
                 -> partial [S | rhs | cost]     (bslam_iterate_pre: one CUDA graph)
    all ranks  : ONE all-reduce (sum, fp64) of that buffer, packed to its
                 structurally non-zero 32x32 tiles                 (NCCL)
    every rank : factorise S, solve dx_c (redundantly, bit-identical inputs),
                 back-substitute and retract its own landmarks, cost at the
                 new point                      (bslam_iterate_post: one CUDA graph)
    all ranks  : all-reduce of two scalars (new cost, ||dx_p||^2)

The reference has no distributed path at all; this is the
B200-native replacement for the single-process `spsolve` on the full system.
"""
import numpy as np


def shard_range(n, rank, world):
    """Contiguous block partition of range(n)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_stereo_ba(d, rank, world):
    """Sub-problem of rank `rank`: all poses, landmarks [lo, hi) and their
    observations (point indices renumbered from 0)."""
    lo, hi = shard_range(len(d['pts0']), rank, world)
    keep = (d['pt_idx'] >= lo) & (d['pt_idx'] < hi)
    out = dict(d)
    out.update(pts0=d['pts0'][lo:hi], pts_true=d['pts_true'][lo:hi], pose_idx=d['pose_idx'][keep],
               pt_idx=(d['pt_idx'][keep] - lo).astype(np.int32), obs=d['obs'][keep], n_lm=hi - lo, lm_range=(lo, hi))
    return out


class ShardedSolver:
    """Drives one engine per rank through the sharded iteration."""

    def __init__(self, engine, rank=0, world=1, group=None):
        self.engine, self.rank, self.world, self.group = engine, rank, world, group
        engine.set_shard(rank)
        if world > 1:
            self._merge_structure()

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
        import torch
        import torch.distributed as dist
        c = self.engine.eval_cost()
        if self.world == 1:
            return c
        t = torch.tensor([c], dtype=torch.float64, device=self.engine.scalars_tensor().device)
        dist.all_reduce(t, group=self.group)
        return float(t.item())

    def iterate(self, lam=0., eval_new_cost=True):
        """(cost at the linearisation point, cost at x [+] dx, ||dx||), identical on every rank."""
        eng = self.engine
        if self.world == 1:
            return eng.iterate(lam, eval_new_cost)
        import torch.distributed as dist
        stream = eng.torch_stream()
        # two graph replays around the NCCL all-reduce of the packed non-zero tiles of S (+ rhs + scalars)
        eng.iterate_pre(lam)
        with _on_stream(stream):
            dist.all_reduce(eng.packed_tensor(), group=self.group)
        eng.iterate_post(eval_new_cost)
        with _on_stream(stream):
            dist.all_reduce(eng.scalars_tensor()[1:3], group=self.group)      # COST_NEW, DX_NORM2
        s = eng.scalars()
        return float(s[0]), float(s[1]), float(np.sqrt(s[2]))


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

"""Drop-in replacement for `pyslam.problem` (reference pyslam/problem.py:11-409).

Same public surface -- `Options`, `Problem.add_residual_block`,
`initialize_params`, `set_parameters_constant/variable`, `eval_cost`, `solve`,
`solve_one_iter`, `compute_covariance`, `get_covariance_block`, `summary`,
`param_dict`, `_cost_history` -- and the same iteration/termination semantics
(problem.py:130-180), but the Gauss-Newton
iteration itself (linearise, solve, retract, cost) runs on the GPU through
libbslam.so (pyslam_b200/engine.py).  There is no CPU solve path: without the
CUDA library or without a GPU, `solve()` raises `EngineError`.

What happens to a residual block at `solve()`:
  * built-in types (ReprojectionResidual with a StereoCamera, PoseResidual,
    PoseToPoseResidual) with a built-in loss are packed into SoA batches and
    linearised by CUDA kernels -- their Python `evaluate` is never called;
  * anything else (user-defined residuals such as the notebook's
    CubicResidual, QuadraticResidual, user-defined losses) keeps pyslam's
    duck-typed protocol: `block.evaluate(params, compute_jacobians)` and
    `loss.weight/loss.loss` run in Python and the scaled (e, J) rows are
    uploaded; assembly, solve and retraction still happen on the GPU.

Extensions (default to reference behaviour): `Options.device`,
`Options.lm_lambda` (Levenberg-Marquardt damping, 0 = Gauss-Newton),
`Problem.add_reprojection_batch` (bulk block registration for large problems),
`Problem.last_timings`.
"""
import copy
import gc
import operator
import warnings

import numpy as np

from . import engine as _engine
from .lie import group_of
from .losses import L2Loss, loss_descriptor
from .residuals.blocks import (BLOCK_POSE, BLOCK_POSE_TO_POSE, BLOCK_REPROJECTION, PoseResidual, PoseToPoseResidual,
                               ReprojectionResidual)
from .residuals.motion_only import (BLOCK_MOTION_ONLY, BLOCK_MOTION_ONLY_BATCH, BLOCK_ORIENTATION,
                                    PoseToPoseOrientationResidual, ReprojectionMotionOnlyBatchResidual,
                                    ReprojectionMotionOnlyResidual)
from .residuals.photometric import BLOCK_PHOTOMETRIC, PhotometricResidualSE3
from .sensors.rgbd_camera import RGBDCamera
from .sensors.stereo_camera import StereoCamera

_BUILTIN_BLOCKS = {BLOCK_REPROJECTION: ReprojectionResidual, BLOCK_POSE: PoseResidual,
                   BLOCK_POSE_TO_POSE: PoseToPoseResidual, BLOCK_PHOTOMETRIC: PhotometricResidualSE3,
                   BLOCK_MOTION_ONLY: ReprojectionMotionOnlyResidual,
                   BLOCK_MOTION_ONLY_BATCH: ReprojectionMotionOnlyBatchResidual,
                   BLOCK_ORIENTATION: PoseToPoseOrientationResidual}


def _builtin_kind(block):
    """BLOCK_KIND of a residual whose arithmetic IS the built-in class's: the class itself, or a subclass
    that does not override `evaluate`.  A subclass with its own `evaluate` is a user plug-in: the reference
    always calls the object's method (pyslam/problem.py:349), so it must not be lowered to a CUDA kernel."""
    kind = getattr(type(block), 'BLOCK_KIND', None)
    cls = _BUILTIN_BLOCKS.get(kind)
    if cls is None or getattr(type(block), 'evaluate', None) is not cls.evaluate:
        return None
    return kind


def _builtin_camera(camera):
    """A StereoCamera whose projection is the built-in one (a subclass that overrides `project`, e.g. to add
    distortion, keeps the Python plug-in path)."""
    t = type(camera)
    for base in (StereoCamera, RGBDCamera):
        if (isinstance(camera, base) and t.project is base.project and t.triangulate is base.triangulate
                and getattr(t, 'intrinsics', None) is base.intrinsics):
            return True
    return False


class Options:
    """Optimisation options; the first ten attributes and their defaults are
    the reference's (pyslam/problem.py:14-37)."""

    def __init__(self):
        self.max_iters = 100
        self.min_update_norm = 1e-6
        self.min_cost = 1e-12
        self.min_cost_decrease = 0.9

        self.linesearch_alpha = 0.8
        self.linesearch_max_iters = 10
        self.linesearch_min_cost_decrease = 0.9

        self.allow_nondecreasing_steps = False
        self.max_nondecreasing_steps = 3

        self.num_threads = 1          # accepted for compatibility; evaluation is on the GPU

        # --- extensions ---
        self.device = 0               # CUDA device ordinal
        self.lm_lambda = 0.           # lambda * diag(H) damping; 0 = the reference's Gauss-Newton
        self.fused_mode = None        # None: library default; 0/1/2 see bslam_set_fused (include/bslam.h)


class _ReprojectionBatch:
    """Many ReprojectionResidual blocks registered at once (arrays instead of
    one Python object per block)."""

    def __init__(self, camera, pose_keys, point_keys, obs, stiffness, loss):
        self.camera, self.pose_keys, self.point_keys = camera, list(pose_keys), list(point_keys)
        self.obs = np.ascontiguousarray(obs, dtype=float).reshape(-1, 3)
        self.stiffness, self.loss = np.asarray(stiffness, dtype=float), loss
        if not (len(self.pose_keys) == len(self.point_keys) == len(self.obs)):
            raise ValueError('pose_keys, point_keys and obs must have the same length')
        # keys factorised once (first-seen order): the lowering then touches each distinct key once, not each block
        self.pose_uniq, self.pose_inv = self._factorise(self.pose_keys)
        self.point_uniq, self.point_inv = self._factorise(self.point_keys)

    @staticmethod
    def _factorise(keys):
        uniq = list(dict.fromkeys(keys))
        index = dict(zip(uniq, range(len(uniq))))
        return uniq, np.fromiter(map(index.__getitem__, keys), np.int32, len(keys))


class _gc_paused:
    """Bulk construction of 10^5 small objects (row views, key lists, table entries) with the cyclic collector on makes
    every generation-2 pass walk all of them again: the lowering of a BASELINE-size problem then takes 0.2 - 0.4 s
    depending on what else the process holds.  None of the objects built here is cyclic."""

    def __enter__(self):
        self.was = gc.isenabled()
        gc.disable()

    def __exit__(self, *exc):
        if self.was:
            gc.enable()
        return False


def _param_dof(p):
    """problem.py:257-266."""
    if hasattr(p, 'dof'):
        return p.dof
    if hasattr(p, '__len__'):
        return len(p)
    return 1


def _pose_row(T, n):
    return np.concatenate([np.asarray(T.rot.mat, dtype=float).reshape(n * n),
                           np.asarray(T.trans, dtype=float).reshape(n)])


class _Lowered:
    """Result of lowering a Problem onto the engine's tables and batches."""
    pass


class Problem:
    """Builds and solves a non-linear least-squares problem (pyslam/problem.py:40)."""

    def __init__(self, options=None, engine=None):
        """`engine` (extension): an existing `pyslam_b200.engine.Engine` handle to lower this problem onto, so that a
        caller that solves many small problems in a row (the dense pipeline's pyramid loop) creates the device
        context, stream and buffers once."""
        self.options = options if options is not None else Options()
        self.param_dict = dict()
        self.residual_blocks = []
        self.block_param_keys = []
        self.block_loss_functions = []
        self.constant_param_keys = []
        self._update_partition_dict = {}
        self._covariance_matrix = None
        self._cost_history = []
        self._batches = []
        self._point_blocks = []     # (keys, (n, 3) block, row views) of initialize_params
        self._engine = engine
        self._low = None
        self.last_timings = None
        self._timing = False

    @property
    def _update_partition_dict(self):
        """key -> range in the update vector (problem.py:252-277); after a lowering it is materialised on first use
        (10^5 range objects that the hot path never reads)."""
        if self._partition is None and self._partition_src is not None:
            keys, start, dof = self._partition_src
            self._partition = {k: range(b, b + d) for k, b, d in zip(keys, start.tolist(), dof.tolist()) if b >= 0}
        return self._partition

    @_update_partition_dict.setter
    def _update_partition_dict(self, value):
        self._partition, self._partition_src = value, None

    # ------------------------------------------------------------------ building
    def add_residual_block(self, block, param_keys, loss=None):
        """Add a cost block (problem.py:72-81).  `loss` defaults to L2Loss()."""
        if isinstance(param_keys, str):
            param_keys = [param_keys]
        self.residual_blocks.append(block)
        self.block_param_keys.append(param_keys)
        self.block_loss_functions.append(loss if loss is not None else L2Loss())
        self._low = None

    def add_reprojection_batch(self, camera, pose_keys, point_keys, obs, stiffness, loss=None):
        """Extension: register len(obs) ReprojectionResidual blocks at once."""
        self._batches.append(_ReprojectionBatch(camera, pose_keys, point_keys, obs, stiffness,
                                                loss if loss is not None else L2Loss()))
        self._low = None

    def initialize_params(self, param_dict):
        """problem.py:83-86 (values are deep-copied)."""
        # Plain float 3-vectors (the 10^5 landmarks of a BA problem) are copied as the rows of ONE block: each
        # parameter still is its own ndarray object, updated in place as the reference's `+=` does (problem.py:405-409),
        # but the whole set moves to / from the device with a single copy (_upload_params / _download_params).
        # Anything else, and any dict whose values alias each other, goes through copy.deepcopy as in the reference.
        vals = list(param_dict.values())
        if len({id(v) for v in vals}) != len(vals):
            self.param_dict.update(copy.deepcopy(param_dict))
            self._low = None
            return
        with _gc_paused():
            nd, f64 = np.ndarray, np.float64
            new = dict.fromkeys(param_dict)         # insertion order is the update-vector order (problem.py:252-277)
            keys3, vals3 = [], []
            for k, v in param_dict.items():
                if type(v) is nd:
                    if v.shape == (3,) and v.dtype == f64:
                        keys3.append(k)
                        vals3.append(v)
                    else:
                        new[k] = v.copy() if v.dtype != object else copy.deepcopy(v)
                else:
                    new[k] = copy.deepcopy(v)
            if len(keys3) >= 64:
                block = np.concatenate(vals3).reshape(len(keys3), 3)
                views = list(block)
                self._point_blocks.append((keys3, block, views))
            else:
                views = [v.copy() for v in vals3]
            new.update(zip(keys3, views))
            self.param_dict.update(new)
        self._low = None

    def set_parameters_constant(self, param_keys):
        if isinstance(param_keys, str):
            param_keys = [param_keys]
        for key in param_keys:
            if key not in self.constant_param_keys:
                self.constant_param_keys.append(key)
        self._low = None

    def set_parameters_variable(self, param_keys):
        if isinstance(param_keys, str):
            param_keys = [param_keys]
        for key in param_keys:
            if key in self.constant_param_keys:
                self.constant_param_keys.remove(key)
        self._low = None

    # ------------------------------------------------------------------ lowering
    def _get_update_partition_dict(self):
        """key -> range in the update vector, param_dict insertion order,
        constants skipped (problem.py:252-277)."""
        part, stop = {}, 0
        for key, param in self.param_dict.items():
            if key not in self.constant_param_keys:
                dof = _param_dof(param)
                part[key] = range(stop, stop + dof)
                stop += dof
        return part

    def _lower(self):
        """Classify parameters and blocks, fill the engine's tables."""
        if self._engine is None:      # raises EngineError without the CUDA library / a GPU
            self._engine = _engine.Engine(getattr(self.options, 'device', 0))
        try:
            with _gc_paused():
                return self._lower_impl()
        except Exception:
            self._low = None
            raise

    def _lower_impl(self):
        pd = self.param_dict
        const = set(self.constant_param_keys)
        low = _Lowered()

        # which 3-vectors act as landmarks of built-in reprojection blocks
        def fusable_reproj(block, keys, loss):
            if _builtin_kind(block) != BLOCK_REPROJECTION or len(keys) != 2:
                return False
            if loss_descriptor(loss) is None or not _builtin_camera(block.camera):
                return False
            T, p = pd.get(keys[0]), pd.get(keys[1])
            return group_of(T) == 'se3' and group_of(p) is None and np.size(p) == 3 and not np.isscalar(p)

        def fusable_pose(block, keys, loss, kind, nkeys):
            if _builtin_kind(block) != kind or len(keys) != nkeys:
                return None
            if loss_descriptor(loss) is None:
                return None
            obs = block.T_obs if kind == BLOCK_POSE else block.T_2_1_obs
            g = group_of(obs)
            if g not in ('se2', 'se3') or any(group_of(pd.get(k)) != g for k in keys):
                return None
            return g

        def fusable_photo(block, keys, loss):
            """'se3': single SE3 parameter; 'split': the (SO3, t) form of pipelines/dense.py:185-190; else None."""
            if (_builtin_kind(block) != BLOCK_PHOTOMETRIC or loss_descriptor(loss) is None
                    or not _builtin_camera(block.camera)):
                return None
            if len(keys) == 1 and group_of(pd.get(keys[0])) == 'se3':
                return 'se3'
            if (len(keys) == 2 and keys[0] != keys[1] and group_of(pd.get(keys[0])) == 'so3'
                    and hasattr(self._engine, 'add_photometric_block_split')):
                t = pd.get(keys[1])
                if group_of(t) is None and isinstance(t, np.ndarray) and t.size == 3:
                    return 'split'
            return None

        def fusable_motion(block, keys, loss):
            return (_builtin_kind(block) in (BLOCK_MOTION_ONLY, BLOCK_MOTION_ONLY_BATCH) and len(keys) == 1
                    and hasattr(self._engine, 'add_motion_only_blocks') and loss_descriptor(loss) is not None
                    and _builtin_camera(block.camera) and group_of(pd.get(keys[0])) == 'se3'
                    and np.shape(block.stiffness) == (3, 3))

        def fusable_orientation(block, keys, loss):
            return (_builtin_kind(block) == BLOCK_ORIENTATION and len(keys) == 2
                    and hasattr(self._engine, 'add_orientation_blocks') and loss_descriptor(loss) is not None
                    and group_of(block.C_2_1_obs) == 'so3' and all(group_of(pd.get(k)) == 'se3' for k in keys))

        kinds = []      # per block: ('reproj',) | ('pose', g) | ('p2p', g) | ('photo', 'se3' | 'split') | ('dense',)
        point_keys = set()
        for block, keys, loss in zip(self.residual_blocks, self.block_param_keys, self.block_loss_functions):
            for k in keys:
                if k not in pd:
                    raise KeyError('Parameter {} has not been initialized'.format(k))
            if fusable_reproj(block, keys, loss):
                kinds.append(('reproj',))
                point_keys.add(keys[1])
                continue
            ph = fusable_photo(block, keys, loss)
            if ph:
                kinds.append(('photo', ph))
                continue
            if fusable_motion(block, keys, loss):
                kinds.append(('motion',))
                continue
            if fusable_orientation(block, keys, loss):
                kinds.append(('orient',))
                continue
            g = fusable_pose(block, keys, loss, BLOCK_POSE, 1)
            if g:
                kinds.append(('pose', g))
                continue
            g = fusable_pose(block, keys, loss, BLOCK_POSE_TO_POSE, 2)
            if g:
                kinds.append(('p2p', g))
                continue
            kinds.append(('dense',))
        for bt in self._batches:
            if loss_descriptor(bt.loss) is None or not _builtin_camera(bt.camera):
                raise ValueError('add_reprojection_batch needs a built-in camera and loss')
            if not pd.keys() >= set(bt.pose_uniq) or not pd.keys() >= set(bt.point_uniq):
                for k in bt.pose_uniq + bt.point_uniq:
                    if k not in pd:
                        raise KeyError('Parameter {} has not been initialized'.format(k))
            point_keys.update(bt.point_uniq)

        # SO3 parameters live in the library's SO3 table when only fused (SO3, t) photometric blocks use them as
        # their rotation; any other use keeps the host-side (opaque manifold) path for the key AND its blocks
        rot_keys = {keys[0] for k, keys in zip(kinds, self.block_param_keys) if k == ('photo', 'split')}
        for k, keys in zip(kinds, self.block_param_keys):
            if k != ('photo', 'split'):
                rot_keys.difference_update(keys)
        for i, (k, keys) in enumerate(zip(kinds, self.block_param_keys)):
            if k == ('photo', 'split') and (keys[0] not in rot_keys or keys[1] in point_keys):
                kinds[i] = ('dense',)
                rot_keys.discard(keys[0])

        # parameter tables, each in param_dict insertion order.  Plain float arrays (the landmarks: 10^5 of them in a
        # BA problem) are classified in bulk; every other parameter type goes through the per-object checks.
        names = ('se3', 'se2', 'pt', 'vec', 'so3')
        keys, vals = list(pd), list(pd.values())
        n = len(keys)
        nd = np.ndarray
        dof = np.fromiter((len(v) if type(v) is nd else -1 for v in vals), np.int64, n)     # problem.py:257-266
        code = np.full(n, 3, np.int8)                                                       # index into `names`
        if point_keys:
            code[np.fromiter(map(point_keys.__contains__, keys), bool, n)] = 2
        low.opaque = set()
        for i in np.flatnonzero(dof < 0):
            key, p = keys[i], vals[i]
            g = group_of(p)
            dof[i] = _param_dof(p)
            if g in ('se3', 'se2'):
                code[i] = names.index(g)
            elif g == 'so3' and key in rot_keys:
                code[i] = 4
            elif g is None and code[i] == 2:
                pass                        # a 3-vector held in a list / tuple
            else:
                code[i] = 3
                if g is not None or (hasattr(p, 'perturb') and hasattr(p, 'dof')):
                    low.opaque.add(key)     # manifold type the library has no kernel for
        variable = ~np.fromiter(map(const.__contains__, keys), bool, n) if const else np.ones(n, bool)
        vdof = np.where(variable, dof, 0)
        start = np.where(variable, np.cumsum(vdof) - vdof, -1)      # offset in the reference's update vector (problem.py:252-277)
        stop = int(vdof.sum())
        low.keys, low.starts = {}, {}
        pos = np.empty(n, np.int64)
        for c, name in enumerate(names):
            ids = np.flatnonzero(code == c)
            pos[ids] = np.arange(ids.size)
            low.keys[name] = [keys[i] for i in ids] if ids.size != n else keys
            low.starts[name] = np.stack([start[ids], dof[ids]], axis=1)     # per table entry: (start, dof), constants start at -1
        low.table = dict(zip(keys, zip([names[c] for c in code.tolist()], pos.tolist())))     # key -> (table, index)
        self._partition_src = (keys, start, dof)        # `_update_partition_dict` is built from this on first use
        self._partition = None
        low.kinds = kinds
        low.dense_ids = [i for i, k in enumerate(kinds) if k[0] == 'dense']
        low.all_fused = not low.dense_ids
        self._low = low

        eng = self._engine
        eng.clear_blocks()
        if getattr(self.options, 'fused_mode', None) is not None:
            eng.set_fused(self.options.fused_mode)
        self._upload_params(pd, structure=True)

        # --- built-in blocks -> batches grouped by (kind, loss, camera) ---
        KIND = {'se3': _engine.KIND_SE3, 'se2': _engine.KIND_SE2, 'pt': _engine.KIND_POINT, 'vec': _engine.KIND_VEC}
        groups = {}
        for i, (block, keys, loss) in enumerate(zip(self.residual_blocks, self.block_param_keys,
                                                    self.block_loss_functions)):
            k = kinds[i]
            if k[0] == 'dense':
                continue
            ld = loss_descriptor(loss)
            if k[0] == 'photo':
                args = (block.uvd_ref, block.im_ref, block.im_jac, block.im_track, block.camera.intrinsics(),
                        block.intensity_stiffness, block.depth_stiffness, ld[0], ld[1])
                if k[1] == 'split':
                    eng.add_photometric_block_split(low.table[keys[0]][1], low.table[keys[1]][1], *args)
                else:
                    eng.add_photometric_block(low.table[keys[0]][1], *args)
                continue
            if k[0] == 'motion':
                # blocks on the same pose with the same camera / stiffness / loss are concatenated into one batch
                gk = ('motion', ld, tuple(block.camera.intrinsics()), keys[0],
                      tuple(np.asarray(block.stiffness, dtype=float).ravel()))
            elif k[0] == 'orient':
                gk = ('orient', ld)
            elif k[0] == 'reproj':
                gk = ('reproj', ld, tuple(block.camera.intrinsics()))
            else:
                gk = (k[0], k[1], ld)
            groups.setdefault(gk, []).append(i)
        for gk, ids in groups.items():
            if gk[0] == 'motion':
                blocks = [self.residual_blocks[i] for i in ids]
                pts = np.vstack([np.atleast_2d(b.pts_1 if hasattr(b, 'pts_1') else b.pt_1) for b in blocks])
                obs = np.vstack([np.atleast_2d(np.asarray(b.obs_2, dtype=float)) for b in blocks])
                eng.add_motion_only_blocks(low.table[gk[3]][1], pts, obs, np.array(gk[4]).reshape(3, 3), gk[2],
                                           gk[1][0], gk[1][1])
                continue
            if gk[0] == 'orient':
                i1 = [low.table[self.block_param_keys[i][0]][1] for i in ids]
                i2 = [low.table[self.block_param_keys[i][1]][1] for i in ids]
                C = np.array([np.asarray(self.residual_blocks[i].C_2_1_obs.mat, dtype=float).ravel() for i in ids])
                S = np.array([np.asarray(self.residual_blocks[i].stiffness, dtype=float).reshape(3, 3) for i in ids])
                if np.all(S == S[0]):
                    S = S[0]
                eng.add_orientation_blocks(i1, i2, C, S, gk[1][0], gk[1][1])
                continue
            if gk[0] == 'reproj':
                pose_idx = [low.table[self.block_param_keys[i][0]][1] for i in ids]
                pt_idx = [low.table[self.block_param_keys[i][1]][1] for i in ids]
                obs = np.array([np.asarray(self.residual_blocks[i].obs, dtype=float).reshape(3) for i in ids])
                S = np.array([np.asarray(self.residual_blocks[i].stiffness, dtype=float).reshape(3, 3) for i in ids])
                if np.all(S == S[0]):
                    S = S[0]
                eng.add_reprojection_blocks(pose_idx, pt_idx, obs, S, gk[2], gk[1][0], gk[1][1])
            else:
                grp = _engine.SE3 if gk[1] == 'se3' else _engine.SE2
                n = 3 if gk[1] == 'se3' else 2
                dof = 6 if gk[1] == 'se3' else 3
                S = np.array([np.asarray(self.residual_blocks[i].stiffness, dtype=float).reshape(dof, dof) for i in ids])
                if np.all(S == S[0]):
                    S = S[0]
                if gk[0] == 'pose':
                    idx = [low.table[self.block_param_keys[i][0]][1] for i in ids]
                    Tobs = np.array([_pose_row(self.residual_blocks[i].T_obs, n) for i in ids])
                    eng.add_pose_blocks(grp, idx, Tobs, S, gk[2][0], gk[2][1])
                else:
                    i1 = [low.table[self.block_param_keys[i][0]][1] for i in ids]
                    i2 = [low.table[self.block_param_keys[i][1]][1] for i in ids]
                    Tobs = np.array([_pose_row(self.residual_blocks[i].T_2_1_obs, n) for i in ids])
                    eng.add_pose_to_pose_blocks(grp, i1, i2, Tobs, S, gk[2][0], gk[2][1])
        index_of = {name: dict(zip(low.keys[name], range(len(low.keys[name])))) for name in ('se3', 'pt')} if self._batches else {}
        for bt in self._batches:
            ld = loss_descriptor(bt.loss)
            try:
                pose_idx = np.array(list(map(index_of['se3'].__getitem__, bt.pose_uniq)), np.int32)[bt.pose_inv]
            except KeyError as e:
                raise ValueError('reprojection batch pose key {} is not an SE3 parameter'.format(e.args[0]))
            try:
                pt_idx = np.array(list(map(index_of['pt'].__getitem__, bt.point_uniq)), np.int32)[bt.point_inv]
            except KeyError as e:
                raise ValueError('reprojection batch point key {} is not a 3-vector parameter'.format(e.args[0]))
            eng.add_reprojection_blocks(pose_idx, pt_idx, bt.obs, bt.stiffness, bt.camera.intrinsics(), ld[0], ld[1])

        # --- host-evaluated blocks: structure ---
        # (the reference drops blocks whose parameters are all constant from the
        #  normal equations, problem.py:343-348, but still counts them in eval_cost)
        low.dense_active = [i for i in low.dense_ids
                            if any(k not in const for k in self.block_param_keys[i])]
        if low.dense_active:
            rows, pptr, pkind, pindex = [], [0], [], []
            for i in low.dense_active:
                keys = self.block_param_keys[i]
                r = np.atleast_1d(self.residual_blocks[i].evaluate([pd[k] for k in keys]))
                rows.append(r.size)
                for k in keys:
                    name, idx = low.table[k]
                    pkind.append(KIND[name])
                    pindex.append(idx)
                pptr.append(len(pkind))
            eng.set_dense_blocks(rows, pptr, pkind, pindex)
            low.dense_rows = rows
        eng.finalize()

        # --- map the engine's internal update ordering to the reference's ---
        lay = eng.layout()
        D = lay['dim']
        total = stop
        if total > D:      # D may exceed the sum of dofs: the reduced span contains padding entries
            raise RuntimeError('internal layout mismatch: {} vs {}'.format(total, D))
        src = np.empty(total, np.int64)
        for name, dof in (('se3', 6), ('se2', 3), ('so3', 3), ('pt', 3)):     # fixed-dof tables: one gather each
            sd = low.starts[name]
            if not len(sd):
                continue
            var = np.flatnonzero(sd[:, 0] >= 0)
            if np.any(sd[var, 1] != dof):
                raise RuntimeError('internal layout mismatch for {} parameters'.format(name))
            src[sd[var, 0][:, None] + np.arange(dof)] = lay[name][var].astype(np.int64)[:, None] + np.arange(dof)
        for (b, d), o in zip(low.starts['vec'].tolist(), lay['vec'].tolist()):
            if b >= 0:
                src[b:b + d] = np.arange(o, o + d)
        low.ref_from_internal = src
        low.dim = D
        low.layout = lay
        return low

    def _vec_values(self, key, p):
        if key in self._low.opaque:
            return np.zeros(_param_dof(p))
        return np.atleast_1d(np.asarray(p, dtype=float)).ravel()

    def _point_block(self, pd):
        """The (n, 3) array whose rows ARE the point parameters of `pd` (same objects, table order), or None."""
        ks = self._low.keys['pt']
        for keys, block, views in self._point_blocks:
            if len(keys) == len(ks) and keys == ks and all(map(operator.is_, map(pd.__getitem__, ks), views)):
                return block
        return None

    def _upload_params(self, pd, structure=False):
        """Host parameter objects -> device tables."""
        low, eng = self._low, self._engine
        const = set(self.constant_param_keys)
        flags = (lambda keys: [k in const for k in keys]) if structure else (lambda keys: None)
        ks = low.keys
        if ks['se3'] or structure:
            eng.set_poses_se3(np.array([_pose_row(pd[k], 3) for k in ks['se3']]).reshape(-1, 12), flags(ks['se3']))
        if ks['se2'] or structure:
            eng.set_poses_se2(np.array([_pose_row(pd[k], 2) for k in ks['se2']]).reshape(-1, 6), flags(ks['se2']))
        if ks['so3'] or (structure and hasattr(eng, 'set_rotations_so3')):
            eng.set_rotations_so3(np.array([np.asarray(pd[k].mat, dtype=float).ravel() for k in ks['so3']]).reshape(-1, 9),
                                  flags(ks['so3']))
        if ks['pt'] or structure:
            xyz = self._point_block(pd)
            if xyz is None:
                vals = [pd[k] for k in ks['pt']]
                try:        # one pass when every point already is a float64 3-vector
                    xyz = np.concatenate(vals).reshape(len(vals), 3) if vals else np.zeros((0, 3))
                    if xyz.dtype != np.float64:
                        raise ValueError
                except ValueError:
                    xyz = np.array([np.asarray(v, dtype=float).reshape(3) for v in vals]).reshape(-1, 3)
            eng.set_points(xyz, flags(ks['pt']))
        if ks['vec'] or structure:
            vals = [self._vec_values(k, pd[k]) for k in ks['vec']]
            dims = [v.size for v in vals]
            eng.set_vectors(dims, np.concatenate(vals) if vals else np.zeros(0), flags(ks['vec']))

    def _download_params(self, dx_ref=None):
        """Device tables -> the objects in `param_dict` (in place where the type
        allows, as the reference's perturb / += do, problem.py:400-409).
        Opaque manifold parameters are perturbed on the host with `dx_ref`."""
        low, eng, pd = self._low, self._engine, self.param_dict
        const = set(self.constant_param_keys)
        ks = low.keys
        if ks['se3']:
            for k, row in zip(ks['se3'], eng.get_poses_se3()):
                if k not in const:
                    pd[k].rot.mat = row[:9].reshape(3, 3).copy()
                    pd[k].trans = row[9:].copy()
        if ks['se2']:
            for k, row in zip(ks['se2'], eng.get_poses_se2()):
                if k not in const:
                    pd[k].rot.mat = row[:4].reshape(2, 2).copy()
                    pd[k].trans = row[4:].copy()
        if ks['so3']:
            for k, row in zip(ks['so3'], eng.get_rotations_so3()):
                if k not in const:
                    pd[k].mat = row.reshape(3, 3).copy()
        block = self._point_block(pd) if ks['pt'] else None
        if block is not None:
            eng.get_points(out=block)       # the rows are the parameters (constant points come back unchanged)
        elif ks['pt']:
            f64, nd = np.float64, np.ndarray
            for k, row in zip(ks['pt'], eng.get_points()):
                if k not in const:
                    p = pd[k]
                    if type(p) is nd and p.dtype == f64 and p.shape == (3,):
                        p[:] = row                     # in place, as the reference's `+=` (problem.py:405-409)
                    else:
                        self._assign_vector(k, row)
        if ks['vec']:
            vals, pos = eng.get_vectors(), 0
            for k in ks['vec']:
                n = _param_dof(pd[k])
                if k not in const:
                    if k in low.opaque:
                        if dx_ref is not None:
                            pd[k].perturb(dx_ref[self._update_partition_dict[k]])
                    else:
                        self._assign_vector(k, vals[pos:pos + n])
                pos += n

    def _assign_vector(self, key, values):
        p = self.param_dict[key]
        if isinstance(p, np.ndarray) and p.dtype.kind == 'f' and p.shape == np.shape(values):
            p[...] = values
        elif isinstance(p, np.ndarray):
            self.param_dict[key] = np.array(values, dtype=float).reshape(p.shape)
        else:
            # python scalars / lists become float arrays after `+=`, as in the reference
            self.param_dict[key] = np.array(values, dtype=float)

    # ------------------------------------------------ host-evaluated (plug-in) blocks
    def _dense_linearize(self):
        """problem.py:338-360 for the blocks that only exist as Python code."""
        low, pd = self._low, self.param_dict
        const = set(self.constant_param_keys)
        e_parts, J_parts, cost = [], [], 0.
        for i, nrows in zip(low.dense_active, low.dense_rows):
            block, keys, loss = self.residual_blocks[i], self.block_param_keys[i], self.block_loss_functions[i]
            cj = [k not in const for k in keys]
            residual, jacs = block.evaluate([pd[k] for k in keys], cj)
            residual = np.atleast_1d(np.asarray(residual, dtype=float)).ravel()
            if residual.size != nrows:
                raise ValueError('residual block {} changed its size ({} -> {})'.format(i, nrows, residual.size))
            sqrt_w = np.sqrt(np.asarray(loss.weight(residual), dtype=float)).ravel()
            cols = []
            for k, want, jac in zip(keys, cj, jacs):
                dof = _param_dof(pd[k])
                if not want or jac is None:
                    cols.append(np.zeros((nrows, dof)))
                else:
                    cols.append(sqrt_w[:, None] * np.asarray(jac, dtype=float).reshape(nrows, dof))
            J_parts.append(np.hstack(cols).ravel())
            e_parts.append(sqrt_w * residual)
            cost += float(np.sum(loss.loss(residual)))
        self._engine.upload_dense_values(np.concatenate(e_parts), np.concatenate(J_parts), cost)

    def _dense_cost(self, ids, pd):
        cost = 0.
        for i in ids:
            r = self.residual_blocks[i].evaluate([pd[k] for k in self.block_param_keys[i]])
            cost += float(np.sum(self.block_loss_functions[i].loss(np.asarray(r, dtype=float))))
        return cost

    # ------------------------------------------------------------------ public API
    def _ensure_lowered(self, upload=True):
        if self._low is None:
            self._lower()
        elif upload:
            self._upload_params(self.param_dict)
        return self._low

    def eval_cost(self, param_dict=None):
        """Sum of loss(residual) over all blocks (problem.py:110-128)."""
        low = self._ensure_lowered()
        eng = self._engine
        pd = self.param_dict
        if param_dict is not None and param_dict is not self.param_dict:
            pd = dict(self.param_dict)
            pd.update(param_dict)
            self._upload_params(pd)
        cost = eng.eval_cost() + self._dense_cost(low.dense_ids, pd)
        if pd is not self.param_dict:
            self._upload_params(self.param_dict)
        return cost

    def _step(self, apply):
        """One iteration on the device.  Returns (dx in reference order or None,
        ||dx||, cost as `solve_one_iter` reports it)."""
        low, eng, opt = self._low, self._engine, self.options
        linesearch = opt.linesearch_max_iters > 0
        if low.dense_active:
            self._dense_linearize()
        if not apply:
            eng.snapshot()
        cost_lin, cost_new, dx_norm = eng.iterate(getattr(opt, 'lm_lambda', 0.), linesearch)
        self.last_timings = eng.timings() if getattr(self, '_timing', False) else None
        dx_ref = None
        need_host = bool(low.dense_ids) or bool(low.opaque)
        if not apply or need_host:
            dx_ref = eng.get_update(low.dim)[low.ref_from_internal]
        if not apply:
            if linesearch and low.dense_ids:
                trial = copy.deepcopy(self.param_dict)
                saved, self.param_dict = self.param_dict, trial
                self._download_params(dx_ref)
                self.param_dict = saved
                cost_new += self._dense_cost(low.dense_ids, trial)
            eng.restore()
        elif need_host:
            self._download_params(dx_ref)
            if linesearch:
                cost_new += self._dense_cost(low.dense_ids, self.param_dict)
        last = eng.last_scalars() if hasattr(eng, 'last_scalars') else eng.scalars()    # mirror of the iterate's read-back
        if last[_engine.S_CHOL_FAIL] > 0:
            warnings.warn('reduced normal matrix is not positive definite; the update is unreliable')
        if linesearch:
            cost = cost_new if np.isfinite(cost_new) else np.inf     # problem.py:362-398 net effect
        else:
            cost = cost_lin
        return dx_ref, dx_norm, cost

    def solve_one_iter(self):
        """(dx, cost) of one Gauss-Newton iteration without applying it
        (problem.py:182-194).  dx is in the reference's ordering."""
        self._ensure_lowered()
        dx, _, cost = self._step(apply=False)
        return dx, cost

    def solve(self):
        """Gauss-Newton with the reference's termination logic (problem.py:130-180)."""
        opt = self.options
        low = self._ensure_lowered()
        eng = self._engine
        cost = eng.eval_cost() + self._dense_cost(low.dense_ids, self.param_dict)
        iters, nondecreasing = 0, 0
        self._cost_history = [cost]
        best_host = None
        done = False
        while not done:
            iters += 1
            prev_cost = cost
            _, dx_norm, cost = self._step(apply=True)
            self._cost_history.append(cost)
            done = iters > opt.max_iters or dx_norm < opt.min_update_norm or cost < opt.min_cost
            if opt.allow_nondecreasing_steps:
                if nondecreasing == 0:
                    eng.snapshot()
                    if low.opaque:
                        best_host = {k: copy.deepcopy(self.param_dict[k]) for k in low.opaque}
                if cost >= opt.min_cost_decrease * prev_cost:
                    nondecreasing += 1
                else:
                    nondecreasing = 0
                if nondecreasing >= opt.max_nondecreasing_steps:
                    done = True
                    eng.restore()
                    if best_host:
                        self.param_dict.update(best_host)
            else:
                done = done or cost >= opt.min_cost_decrease * prev_cost
        self._download_params()
        return self.param_dict

    def compute_covariance(self):
        """Covariance of the final estimate = inverse of the normal matrix at the
        current parameters (problem.py:196-203)."""
        try:
            low = self._ensure_lowered()
            if low.dense_active:
                self._dense_linearize()
            cov_int = self._engine.covariance(low.dim)
            idx = low.ref_from_internal
            self._covariance_matrix = cov_int[np.ix_(idx, idx)]
        except Exception as e:       # the reference swallows and prints (problem.py:202-203)
            print('Covariance computation failed!\n{}'.format(e))

    def get_covariance_block(self, param0, param1):
        """problem.py:205-216."""
        try:
            r0 = self._update_partition_dict[param0]
            r1 = self._update_partition_dict[param1]
            return np.squeeze(self._covariance_matrix[r0.start:r0.stop, r1.start:r1.stop])
        except KeyError as e:
            print('Cannot compute covariance for constant parameter {}'.format(e.args[0]))
        return None

    def summary(self, format='brief'):
        """Same text as the reference's summary (problem.py:218-250)."""
        if not self._cost_history:
            raise ValueError('solve has not yet been called')
        h = self._cost_history
        if format == 'brief':
            return 'Iterations: {:3} | Cost: {:12e} --> {:12e}'.format(len(h), h[0], h[-1])
        if format == 'full':
            header = '{:>5s} | {:>12s} --> {:>12s} | {:>10s}\n'.format('Iter', 'Initial cost', 'Final cost', 'Rel change')
            lines = [header, '-' * len(header) + '\n']
            for i, (ic, fc) in enumerate(zip(h[:-1], h[1:])):
                lines.append('{:5} | {:12e} --> {:12e} | {:+10f}\n'.format(i, ic, fc, (fc - ic) / ic))
            return ''.join(lines)
        raise ValueError('Invalid summary format \'{}\'.'.format(format) +
                         'Valid formats are \'brief\' and \'full\'')

    def enable_timing(self, on=True):
        """Extension: record per-phase CUDA-event timings into `last_timings`."""
        self._timing = bool(on)
        self._ensure_lowered(upload=False)
        self._engine.enable_timing(on)

"""ctypes binding of libbslam.so (include/bslam.h) -- the only door from the
Python host into the CUDA solver.

There is deliberately no fallback: if the shared library is missing, cannot be
loaded, or no GPU is present, every entry point raises `EngineError`.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# BSLAM_LIB: developer override used to A/B kernel build variants (tools/variants.sh)
LIB_PATH = os.environ.get('BSLAM_LIB') or os.path.join(_HERE, 'libbslam.so')

N_SCALARS = 16
N_TIMINGS = 16
S_COST_LIN, S_COST_NEW, S_DX_NORM2, S_CHOL_FAIL, S_COST_EVAL = 0, 1, 2, 3, 4
TIMING_NAMES = ('linearize', 'reproj', 'schur', 'cholesky', 'trsv', 'backsub', 'retract', 'cost', 'total', 'fused', 'peer_publish',
                'peer_scalars')
SE2, SE3 = 2, 3
KIND_SE3, KIND_SE2, KIND_POINT, KIND_VEC = 0, 1, 2, 3


class EngineError(RuntimeError):
    pass


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_bp = C.POINTER(C.c_uint8)
_h = C.c_void_p

# name -> (restype, argtypes); must list every symbol declared in include/bslam.h
SIGNATURES = {
    'bslam_version': (C.c_int, []),
    'bslam_tile_edge': (C.c_int, []),
    'bslam_create': (C.c_int, [C.POINTER(_h), C.c_int]),
    'bslam_destroy': (None, [_h]),
    'bslam_last_error': (C.c_char_p, [_h]),
    'bslam_set_poses_se3': (C.c_int, [_h, C.c_int, _dp, _bp]),
    'bslam_set_poses_se2': (C.c_int, [_h, C.c_int, _dp, _bp]),
    'bslam_set_points': (C.c_int, [_h, C.c_int, _dp, _bp]),
    'bslam_set_vectors': (C.c_int, [_h, C.c_int, _ip, _dp, _bp]),
    'bslam_set_rotations_so3': (C.c_int, [_h, C.c_int, _dp, _bp]),
    'bslam_get_rotations_so3': (C.c_int, [_h, _dp]),
    'bslam_get_layout_so3': (C.c_int, [_h, _ip]),
    'bslam_add_photometric_block_split': (C.c_int, [_h, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, _dp, C.c_int, C.c_int, _dp,
                                                    C.c_double, C.c_double, C.c_int, C.c_double]),
    'bslam_get_poses_se3': (C.c_int, [_h, _dp]),
    'bslam_get_poses_se2': (C.c_int, [_h, _dp]),
    'bslam_get_points': (C.c_int, [_h, _dp]),
    'bslam_get_vectors': (C.c_int, [_h, _dp]),
    'bslam_add_reprojection_blocks': (C.c_int, [_h, C.c_int, _ip, _ip, _dp, _dp, C.c_int, _dp, C.c_int, C.c_double]),
    'bslam_add_pose_blocks': (C.c_int, [_h, C.c_int, C.c_int, _ip, _dp, _dp, C.c_int, C.c_int, C.c_double]),
    'bslam_add_pose_to_pose_blocks': (C.c_int, [_h, C.c_int, C.c_int, _ip, _ip, _dp, _dp, C.c_int, C.c_int, C.c_double]),
    'bslam_add_photometric_block': (C.c_int, [_h, C.c_int, C.c_int, _dp, _dp, _dp, _dp, C.c_int, C.c_int, _dp, C.c_double,
                                              C.c_double, C.c_int, C.c_double]),
    'bslam_add_motion_only_blocks': (C.c_int, [_h, C.c_int, C.c_int, _dp, _dp, _dp, _dp, C.c_int, C.c_double]),
    'bslam_add_orientation_blocks': (C.c_int, [_h, C.c_int, _ip, _ip, _dp, _dp, C.c_int, C.c_int, C.c_double]),
    'bslam_ransac': (C.c_int, [C.c_int, C.c_int, C.c_int, _ip, _dp, C.c_int, _dp, _dp, _dp, _dp, C.c_double, _dp, _ip, _ip, _bp]),
    'bslam_image_pyramid': (C.c_int, [C.c_int, _bp, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp]),
    'bslam_subsample_pyramid': (C.c_int, [C.c_int, _dp, C.c_int, C.c_int, C.c_int, C.c_double, _dp]),
    'bslam_set_dense_blocks': (C.c_int, [_h, C.c_int, _ip, _ip, _ip, _ip]),
    'bslam_upload_dense_values': (C.c_int, [_h, _dp, C.c_size_t, _dp, C.c_size_t, C.c_double]),
    'bslam_clear_blocks': (C.c_int, [_h]),
    'bslam_finalize': (C.c_int, [_h]),
    'bslam_get_layout': (C.c_int, [_h, _ip, _ip, _ip, _ip, _ip, _ip]),
    'bslam_eval_cost': (C.c_int, [_h, _dp]),
    'bslam_iterate': (C.c_int, [_h, C.c_double, C.c_int, _dp, _dp, _dp]),
    'bslam_iterate_host': (C.c_int, [_h, C.c_double, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp, _dp]),
    'bslam_linearize': (C.c_int, [_h, _dp]),
    'bslam_reduce': (C.c_int, [_h, C.c_double]),
    'bslam_solve_reduced': (C.c_int, [_h]),
    'bslam_retract': (C.c_int, [_h, C.c_int]),
    'bslam_get_scalars': (C.c_int, [_h, _dp]),
    'bslam_last_scalars': (C.c_int, [_h, _dp]),
    'bslam_linearize_reduce': (C.c_int, [_h, C.c_double]),
    'bslam_retract_iterate': (C.c_int, [_h, C.c_int]),
    'bslam_set_fused': (C.c_int, [_h, C.c_int]),
    'bslam_get_fused': (C.c_int, [_h, _ip, _ip]),
    'bslam_reduced_buffer': (C.c_int, [_h, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(C.c_void_p), _ip]),
    'bslam_packed_buffer': (C.c_int, [_h, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    'bslam_iterate_pre': (C.c_int, [_h, C.c_double]),
    'bslam_iterate_post': (C.c_int, [_h, C.c_int]),
    'bslam_pack_reduced': (C.c_int, [_h, C.c_int]),
    'bslam_tile_structure': (C.c_int, [_h, _bp, C.c_size_t, C.c_int]),
    'bslam_set_shard': (C.c_int, [_h, C.c_int]),
    'bslam_add_coupling': (C.c_int, [_h, C.c_int, C.c_int, _ip, _ip]),
    'bslam_layout_hash': (C.c_int, [_h, C.POINTER(C.c_uint64)]),
    'bslam_peer_region': (C.c_int, [_h, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), _bp]),
    'bslam_peer_connect': (C.c_int, [_h, C.c_int, C.c_int, _bp, C.POINTER(C.c_void_p)]),
    'bslam_peer_barrier': (C.c_int, [_h]),
    'bslam_peer_local_slots': (C.c_int, [_h, _bp, C.c_size_t]),
    'bslam_peer_set_contributors': (C.c_int, [_h, _bp, C.c_size_t]),
    'bslam_peer_connect_symmetric': (C.c_int, [_h, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.c_void_p, C.c_size_t]),
    'bslam_iterate_async': (C.c_int, [_h, C.c_double, C.c_int]),
    'bslam_iterate_wait': (C.c_int, [_h, _dp, _dp, _dp]),
    'bslam_stream': (C.c_void_p, [_h]),
    'bslam_snapshot': (C.c_int, [_h]),
    'bslam_restore': (C.c_int, [_h]),
    'bslam_get_update': (C.c_int, [_h, _dp]),
    'bslam_get_normal_equations': (C.c_int, [_h, _dp, _dp]),
    'bslam_get_reduced_system': (C.c_int, [_h, _dp, _dp]),
    'bslam_covariance': (C.c_int, [_h, _dp]),
    'bslam_debug_chol_trace': (C.c_int, [_h, C.POINTER(C.c_int64), C.c_int, _ip]),
    'bslam_enable_timing': (C.c_int, [_h, C.c_int]),
    'bslam_get_timings': (C.c_int, [_h, _dp]),
    'bslam_launch_count': (C.c_int64, [_h]),
}

_lib = None


def load_library(path=None):
    """dlopen libbslam.so and attach the prototypes.  Raises EngineError when
    the library has not been built (run `python -c 'import __graft_entry__ as g;
    g.build()'` or `make -C pyslam_b200/csrc`)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.isfile(path):
        raise EngineError('CUDA extension {} not found -- build it with `make -C pyslam_b200/csrc`; '
                          'there is no CPU fallback'.format(path))
    try:
        lib = C.CDLL(path)
    except OSError as e:
        raise EngineError('cannot load {}: {}'.format(path, e))
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            raise EngineError('{} does not export {}'.format(path, name))
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


def _b(a):
    return None if a is None else a.ctypes.data_as(_bp)


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _flags(a, n):
    if a is None:
        return None
    a = np.ascontiguousarray(np.asarray(a).astype(bool), dtype=np.uint8)
    assert a.shape == (n,)
    return a


class Engine:
    """One libbslam solver handle on one GPU."""

    def __init__(self, device=0):
        self._lib = load_library()
        self._h = _h()
        rc = self._lib.bslam_create(C.byref(self._h), int(device))
        if rc != 0:
            msg = self._lib.bslam_last_error(None)
            self._h = None
            raise EngineError('bslam_create failed ({}): {}'.format(rc, (msg or b'').decode()))
        self.device = int(device)
        self.n = dict(se3=0, se2=0, pt=0, vec=0, vec_entries=0, so3=0)
        self._scal = np.zeros(N_SCALARS)

    def close(self):
        if getattr(self, '_h', None):
            self._lib.bslam_destroy(self._h)
            self._h = None

    __del__ = close

    def _ck(self, rc):
        if rc != 0:
            raise EngineError('libbslam error {}: {}'.format(rc, (self._lib.bslam_last_error(self._h) or b'').decode()))

    # ---- parameter tables -------------------------------------------------
    def set_poses_se3(self, Rt, is_const=None):
        Rt = _f64(Rt, (-1, 12))
        self.n['se3'] = len(Rt)
        self._ck(self._lib.bslam_set_poses_se3(self._h, len(Rt), _d(Rt), _b(_flags(is_const, len(Rt)))))

    def set_poses_se2(self, Rt, is_const=None):
        Rt = _f64(Rt, (-1, 6))
        self.n['se2'] = len(Rt)
        self._ck(self._lib.bslam_set_poses_se2(self._h, len(Rt), _d(Rt), _b(_flags(is_const, len(Rt)))))

    def set_rotations_so3(self, R, is_const=None):
        R = _f64(R, (-1, 9))
        self.n['so3'] = len(R)
        self._ck(self._lib.bslam_set_rotations_so3(self._h, len(R), _d(R), _b(_flags(is_const, len(R)))))

    def get_rotations_so3(self):
        out = np.empty((self.n['so3'], 9))
        self._ck(self._lib.bslam_get_rotations_so3(self._h, _d(out)))
        return out

    def set_points(self, xyz, is_const=None):
        xyz = _f64(xyz, (-1, 3))
        self.n['pt'] = len(xyz)
        self._ck(self._lib.bslam_set_points(self._h, len(xyz), _d(xyz), _b(_flags(is_const, len(xyz)))))

    def set_vectors(self, dims, values, is_const=None):
        dims = _i32(dims)
        values = _f64(values).ravel()
        assert values.size == int(dims.sum())
        self.n['vec'], self.n['vec_entries'] = len(dims), values.size
        self._ck(self._lib.bslam_set_vectors(self._h, len(dims), _i(dims), _d(values), _b(_flags(is_const, len(dims)))))

    def get_poses_se3(self, out=None):
        out = np.empty((self.n['se3'], 12)) if out is None else out
        self._ck(self._lib.bslam_get_poses_se3(self._h, _d(out)))
        return out

    def get_poses_se2(self):
        out = np.empty((self.n['se2'], 6))
        self._ck(self._lib.bslam_get_poses_se2(self._h, _d(out)))
        return out

    def get_points(self, out=None):
        out = np.empty((self.n['pt'], 3)) if out is None else out
        self._ck(self._lib.bslam_get_points(self._h, _d(out)))
        return out

    def get_vectors(self):
        out = np.empty(self.n['vec_entries'])
        self._ck(self._lib.bslam_get_vectors(self._h, _d(out)))
        return out

    # ---- blocks -------------------------------------------------------------
    @staticmethod
    def _stiff(stiffness, n, d):
        S = _f64(stiffness)
        if S.size == d * d:
            return S.reshape(d, d), 0
        if S.size == n * d * d:
            return S.reshape(n, d, d), 1
        raise ValueError('stiffness must hold 1 or {} matrices of {}x{}'.format(n, d, d))

    def add_reprojection_blocks(self, pose_idx, pt_idx, obs, stiffness, intr, loss_kind=0, loss_k=0.):
        pose_idx, pt_idx, obs = _i32(pose_idx), _i32(pt_idx), _f64(obs, (-1, 3))
        n = len(pose_idx)
        assert len(pt_idx) == n and len(obs) == n
        S, per = self._stiff(stiffness, n, 3)
        intr = _f64(intr, (5,))
        self._ck(self._lib.bslam_add_reprojection_blocks(self._h, n, _i(pose_idx), _i(pt_idx), _d(obs), _d(S), per,
                                                         _d(intr), int(loss_kind), float(loss_k)))

    def add_pose_blocks(self, group, pose_idx, T_obs, stiffness, loss_kind=0, loss_k=0.):
        store, dof = (12, 6) if group == SE3 else (6, 3)
        pose_idx, T_obs = _i32(pose_idx), _f64(T_obs, (-1, store))
        n = len(pose_idx)
        S, per = self._stiff(stiffness, n, dof)
        self._ck(self._lib.bslam_add_pose_blocks(self._h, group, n, _i(pose_idx), _d(T_obs), _d(S), per,
                                                 int(loss_kind), float(loss_k)))

    def add_pose_to_pose_blocks(self, group, idx1, idx2, T21_obs, stiffness, loss_kind=0, loss_k=0.):
        store, dof = (12, 6) if group == SE3 else (6, 3)
        idx1, idx2, T21_obs = _i32(idx1), _i32(idx2), _f64(T21_obs, (-1, store))
        n = len(idx1)
        S, per = self._stiff(stiffness, n, dof)
        self._ck(self._lib.bslam_add_pose_to_pose_blocks(self._h, group, n, _i(idx1), _i(idx2), _d(T21_obs), _d(S), per,
                                                         int(loss_kind), float(loss_k)))

    def add_photometric_block(self, pose_idx, uvd_ref, im_ref, im_jac, im_track, intr, intensity_stiffness,
                              depth_stiffness, loss_kind=0, loss_k=0.):
        uvd_ref, im_ref, im_jac = _f64(uvd_ref, (-1, 3)), _f64(im_ref).ravel(), _f64(im_jac, (-1, 2))
        im_track, intr = _f64(im_track), _f64(intr, (5,))
        assert im_track.ndim == 2 and len(im_ref) == len(uvd_ref) == len(im_jac)
        self._ck(self._lib.bslam_add_photometric_block(
            self._h, int(pose_idx), len(uvd_ref), _d(uvd_ref), _d(im_ref), _d(im_jac), _d(im_track), im_track.shape[1],
            im_track.shape[0], _d(intr), float(intensity_stiffness), float(depth_stiffness), int(loss_kind), float(loss_k)))

    def add_photometric_block_split(self, rot_idx, vec_idx, uvd_ref, im_ref, im_jac, im_track, intr, intensity_stiffness,
                                    depth_stiffness, loss_kind=0, loss_k=0.):
        """The (SO3, t) parameter form: rot_idx -> SO3 table, vec_idx -> a 3-vector of the vector table."""
        uvd_ref, im_ref, im_jac = _f64(uvd_ref, (-1, 3)), _f64(im_ref).ravel(), _f64(im_jac, (-1, 2))
        im_track, intr = _f64(im_track), _f64(intr, (5,))
        assert im_track.ndim == 2 and len(im_ref) == len(uvd_ref) == len(im_jac)
        self._ck(self._lib.bslam_add_photometric_block_split(
            self._h, int(rot_idx), int(vec_idx), len(uvd_ref), _d(uvd_ref), _d(im_ref), _d(im_jac), _d(im_track),
            im_track.shape[1], im_track.shape[0], _d(intr), float(intensity_stiffness), float(depth_stiffness),
            int(loss_kind), float(loss_k)))

    def add_motion_only_blocks(self, pose_idx, pts_1, obs_2, stiffness, intr, loss_kind=0, loss_k=0.):
        """ReprojectionMotionOnly(Batch)Residual: n fixed points seen from pose `pose_idx` (T_2_1)."""
        pts_1, obs_2 = _f64(pts_1, (-1, 3)), _f64(obs_2, (-1, 3))
        S, intr = _f64(stiffness, (3, 3)), _f64(intr, (5,))
        assert len(pts_1) == len(obs_2)
        self._ck(self._lib.bslam_add_motion_only_blocks(self._h, int(pose_idx), len(pts_1), _d(pts_1), _d(obs_2), _d(S), _d(intr),
                                                        int(loss_kind), float(loss_k)))

    def add_orientation_blocks(self, idx1, idx2, C21_obs, stiffness, loss_kind=0, loss_k=0.):
        """PoseToPoseOrientationResidual: relative rotation measurements (n x 9) between SE3 poses."""
        i1, i2 = _i32(idx1), _i32(idx2)
        C = _f64(C21_obs, (-1, 9))
        S, per = self._stiff(stiffness, len(i1), 3)
        self._ck(self._lib.bslam_add_orientation_blocks(self._h, len(i1), _i(i1), _i(i2), _d(C), _d(S), per, int(loss_kind), float(loss_k)))

    def set_dense_blocks(self, rows, param_ptr, param_kind, param_index):
        rows, param_ptr = _i32(rows), _i32(param_ptr)
        param_kind, param_index = _i32(param_kind), _i32(param_index)
        self._ck(self._lib.bslam_set_dense_blocks(self._h, len(rows), _i(rows), _i(param_ptr), _i(param_kind),
                                                  _i(param_index)))

    def upload_dense_values(self, e, J, cost):
        e, J = _f64(e).ravel(), _f64(J).ravel()
        self._ck(self._lib.bslam_upload_dense_values(self._h, _d(e), e.size, _d(J), J.size, float(cost)))

    def clear_blocks(self):
        self._ck(self._lib.bslam_clear_blocks(self._h))

    def finalize(self):
        self._ck(self._lib.bslam_finalize(self._h))

    def layout(self):
        """dict(se3=offsets, se2=..., pt=..., vec=..., dim=D, n_reduced=n)."""
        o = {k: np.empty(self.n[k], np.int32) for k in ('se3', 'se2', 'pt', 'vec')}
        dim, nred = C.c_int32(), C.c_int32()
        self._ck(self._lib.bslam_get_layout(self._h, _i(o['se3']), _i(o['se2']), _i(o['pt']), _i(o['vec']),
                                            C.byref(dim), C.byref(nred)))
        o['dim'], o['n_reduced'] = dim.value, nred.value
        o['so3'] = np.empty(self.n['so3'], np.int32)
        self._ck(self._lib.bslam_get_layout_so3(self._h, _i(o['so3'])))
        return o

    # ---- hot path -----------------------------------------------------------
    def eval_cost(self):
        c = C.c_double()
        self._ck(self._lib.bslam_eval_cost(self._h, C.byref(c)))
        return c.value

    def iterate(self, lam=0., eval_new_cost=True):
        """(cost at linearisation point, cost at x [+] dx, ||dx||)."""
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        self._ck(self._lib.bslam_iterate(self._h, float(lam), int(bool(eval_new_cost)), C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def iterate_host(self, Rt, xyz, lam=0., eval_new_cost=True):
        """One iteration on HOST parameter tables (float64, C-contiguous; pinned memory makes the copies
        asynchronous): `Rt` (n_se3 x 12) and `xyz` (n_pt x 3) are uploaded, updated in place and read back
        with a single synchronisation.  Returns (cost at the linearisation point, cost at x [+] dx, ||dx||)."""
        for arr in (Rt, xyz):
            if not (isinstance(arr, np.ndarray) and arr.dtype == np.float64 and arr.flags['C_CONTIGUOUS']):
                raise EngineError('iterate_host needs C-contiguous float64 arrays (it writes the result in place)')
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        self._ck(self._lib.bslam_iterate_host(self._h, float(lam), int(bool(eval_new_cost)), _d(Rt), _d(xyz), _d(Rt), _d(xyz),
                                              C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def linearize(self, fetch_cost=True):
        c = C.c_double()
        self._ck(self._lib.bslam_linearize(self._h, C.byref(c) if fetch_cost else None))
        return c.value if fetch_cost else None

    def reduce(self, lam=0.):
        self._ck(self._lib.bslam_reduce(self._h, float(lam)))

    def solve_reduced(self):
        self._ck(self._lib.bslam_solve_reduced(self._h))

    def linearize_reduce(self, lam=0.):
        """What `iterate` runs before the reduced solve (fused panels included)."""
        self._ck(self._lib.bslam_linearize_reduce(self._h, float(lam)))

    def retract_iterate(self, eval_new_cost=True):
        self._ck(self._lib.bslam_retract_iterate(self._h, int(bool(eval_new_cost))))

    def set_fused(self, mode):
        """0: landmark-block kernels (W materialised); 1: dense panels fused (default); 2: fuse whatever fits."""
        self._ck(self._lib.bslam_set_fused(self._h, int(mode)))

    def fused_info(self):
        """(number of panels, number of landmarks inside panels) after finalize."""
        a, b = C.c_int32(), C.c_int32()
        self._ck(self._lib.bslam_get_fused(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def retract(self, eval_new_cost=True):
        self._ck(self._lib.bslam_retract(self._h, int(bool(eval_new_cost))))

    def scalars(self):
        self._ck(self._lib.bslam_get_scalars(self._h, _d(self._scal)))
        return self._scal.copy()

    def last_scalars(self):
        """Scalars as read back by the last iterate (no device work, no synchronisation)."""
        self._ck(self._lib.bslam_last_scalars(self._h, _d(self._scal)))
        return self._scal.copy()

    def reduced_buffer(self):
        """(device pointer, length in doubles, device pointer of the scalar tail, n_pad)."""
        p, n, ps, npad = C.c_void_p(), C.c_size_t(), C.c_void_p(), C.c_int32()
        self._ck(self._lib.bslam_reduced_buffer(self._h, C.byref(p), C.byref(n), C.byref(ps), C.byref(npad)))
        return p.value, n.value, ps.value, npad.value

    def tile_structure(self):
        """uint8 mask [(nt+1), nt] of the non-zero tiles (bslam_tile_edge() wide) of the reduced system."""
        _, _, _, npad = self.reduced_buffer()
        nt = npad // self._lib.bslam_tile_edge()
        m = np.zeros((nt + 1, nt), np.uint8)
        self._ck(self._lib.bslam_tile_structure(self._h, _b(m), m.size, 0))
        return m

    def merge_tile_structure(self, mask):
        m = np.ascontiguousarray(mask, dtype=np.uint8)
        self._ck(self._lib.bslam_tile_structure(self._h, _b(m), m.size, 1))

    def iterate_pre(self, lam=0.):
        """linearize + reduce + pack of a sharded iteration (one CUDA graph)."""
        self._ck(self._lib.bslam_iterate_pre(self._h, float(lam)))

    def iterate_post(self, eval_new_cost=True):
        """unpack + reduced solve + retract of a sharded iteration (one CUDA graph)."""
        self._ck(self._lib.bslam_iterate_post(self._h, int(bool(eval_new_cost))))

    def pack_reduced(self, unpack=False):
        self._ck(self._lib.bslam_pack_reduced(self._h, int(bool(unpack))))

    def packed_tensor(self):
        """torch.float64 tensor aliasing the compact all-reduce payload (non-zero tiles | rhs | scalars)."""
        import torch
        p, n = C.c_void_p(), C.c_size_t()
        self._ck(self._lib.bslam_packed_buffer(self._h, C.byref(p), C.byref(n)))
        key = ('pack', p.value, n.value)
        if getattr(self, '_pcache_key', None) != key:
            self._pcache = torch.as_tensor(_DevArray(p.value, n.value), device='cuda:%d' % self.device)
            self._pcache_key = key
        return self._pcache

    # ---- sharded iteration over peer memory (csrc/peer.cuh) ----------------------
    def add_coupling(self, group, idx1, idx2):
        """Declare reduced-system couplings (pose pairs) without residuals: rank-independent ordering."""
        i1, i2 = _i32(idx1), _i32(idx2)
        self._ck(self._lib.bslam_add_coupling(self._h, int(group), len(i1), _i(i1), _i(i2)))

    def layout_hash(self):
        h = C.c_uint64()
        self._ck(self._lib.bslam_layout_hash(self._h, C.byref(h)))
        return int(h.value)

    def peer_region(self):
        """(device pointer, bytes, 64-byte CUDA-IPC handle as uint8 array) of this handle's exchange region."""
        p, n = C.c_void_p(), C.c_size_t()
        handle = np.zeros(64, np.uint8)
        self._ck(self._lib.bslam_peer_region(self._h, C.byref(p), C.byref(n), _b(handle)))
        return p.value, n.value, handle

    def peer_connect(self, world, rank, ipc_handles=None, dev_ptrs=None):
        """Map the exchange regions of all ranks (IPC handles [world, 64] uint8, or device pointers of handles
        of this process) and switch `iterate` to the sharded schedule."""
        hb = None if ipc_handles is None else np.ascontiguousarray(ipc_handles, dtype=np.uint8).reshape(world, 64)
        pa = None
        if dev_ptrs is not None:
            pa = (C.c_void_p * world)(*[C.c_void_p(int(x) if x else 0) for x in dev_ptrs])
        self._ck(self._lib.bslam_peer_connect(self._h, int(world), int(rank), _b(hb), pa))

    def peer_connect_symmetric(self, world, rank, region_ptrs, multicast_ptr, n_bytes):
        """Exchange regions in caller-provided symmetric memory (+ optional NVLS multicast mapping)."""
        pa = (C.c_void_p * world)(*[C.c_void_p(int(x)) for x in region_ptrs])
        self._ck(self._lib.bslam_peer_connect_symmetric(self._h, int(world), int(rank), pa, C.c_void_p(int(multicast_ptr) or None),
                                                        int(n_bytes)))

    def n_packed_tiles(self):
        p, n = C.c_void_p(), C.c_size_t()
        self._ck(self._lib.bslam_packed_buffer(self._h, C.byref(p), C.byref(n)))
        _, _, _, npad = self.reduced_buffer()
        return (n.value - npad - N_SCALARS) // (self._lib.bslam_tile_edge() ** 2)

    def peer_local_slots(self):
        """uint8 flags per packed tile: this handle's blocks can write it."""
        f = np.zeros(self.n_packed_tiles(), np.uint8)
        self._ck(self._lib.bslam_peer_local_slots(self._h, _b(f), f.size))
        return f

    def peer_set_contributors(self, masks):
        m = np.ascontiguousarray(masks, dtype=np.uint8)
        self._ck(self._lib.bslam_peer_set_contributors(self._h, _b(m), m.size))

    def peer_barrier(self):
        """Enqueue a device-side rendezvous of all connected ranks on the handle's stream."""
        self._ck(self._lib.bslam_peer_barrier(self._h))

    def iterate_async(self, lam=0., eval_new_cost=True):
        self._ck(self._lib.bslam_iterate_async(self._h, float(lam), int(bool(eval_new_cost))))

    def iterate_wait(self):
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        self._ck(self._lib.bslam_iterate_wait(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def set_shard(self, rank):
        self._ck(self._lib.bslam_set_shard(self._h, int(rank)))

    def stream(self):
        return self._lib.bslam_stream(self._h)

    def snapshot(self):
        self._ck(self._lib.bslam_snapshot(self._h))

    def restore(self):
        self._ck(self._lib.bslam_restore(self._h))

    # ---- inspection -----------------------------------------------------------
    def get_update(self, dim):
        out = np.empty(dim)
        self._ck(self._lib.bslam_get_update(self._h, _d(out)))
        return out

    def get_normal_equations(self, dim):
        H, b = np.empty((dim, dim)), np.empty(dim)
        self._ck(self._lib.bslam_get_normal_equations(self._h, _d(H), _d(b)))
        return H, b

    def get_reduced_system(self, n):
        S, r = np.empty((n, n)), np.empty(n)
        self._ck(self._lib.bslam_get_reduced_system(self._h, _d(S), _d(r)))
        return S, r

    def covariance(self, dim):
        out = np.empty((dim, dim))
        self._ck(self._lib.bslam_covariance(self._h, _d(out)))
        return out

    def chol_trace(self, max_tasks=20000):
        """[(tile_row|-1, tile_col, start, deps_ready, end, sm)] of one traced reduced solve (ns)."""
        out = np.zeros((max_tasks, 6), np.int64)
        n = C.c_int32()
        self._ck(self._lib.bslam_debug_chol_trace(self._h, out.ctypes.data_as(C.POINTER(C.c_int64)), max_tasks, C.byref(n)))
        return out[:n.value]

    def enable_timing(self, on=True):
        self._ck(self._lib.bslam_enable_timing(self._h, int(bool(on))))

    def timings(self):
        out = np.zeros(N_TIMINGS)
        self._ck(self._lib.bslam_get_timings(self._h, _d(out)))
        return dict(zip(TIMING_NAMES, out[:len(TIMING_NAMES)]))

    def launch_count(self):
        return int(self._lib.bslam_launch_count(self._h))


# ---- torch interop (multi-GPU plumbing and event timing only) -----------------
class _DevArray:
    """Minimal __cuda_array_interface__ carrier so torch can alias library memory."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = dict(shape=(int(n),), typestr='<f8', data=(int(ptr), False), version=2)


def _engine_reduced_tensor(self):
    """torch.float64 tensor aliasing the device buffer [S | rhs | scalars] that is
    all-reduced across GPUs between `reduce` and `solve_reduced`."""
    import torch
    p, n, _, _ = self.reduced_buffer()
    key = ('red', p, n)
    if getattr(self, '_tcache_key', None) != key:
        self._tcache = torch.as_tensor(_DevArray(p, n), device='cuda:%d' % self.device)
        self._tcache_key = key
    return self._tcache


def _engine_scalars_tensor(self):
    return self.reduced_tensor()[-N_SCALARS:]


def _engine_torch_stream(self):
    """The handle's CUDA stream as a torch stream (order collectives / record events on it)."""
    import torch
    if getattr(self, '_tstream', None) is None:
        self._tstream = torch.cuda.ExternalStream(self.stream(), device='cuda:%d' % self.device)
    return self._tstream


Engine.reduced_tensor = _engine_reduced_tensor
Engine.scalars_tensor = _engine_scalars_tensor
Engine.torch_stream = _engine_torch_stream


# ---- stand-alone device routines (no solver handle) ---------------------------------------------
def _ck_global(lib, rc):
    if rc != 0:
        raise EngineError('libbslam error {}: {}'.format(rc, (lib.bslam_last_error(None) or b'').decode()))


def ransac(pts_1, obs_2, intr, thresh, sample_idx=None, pts_2=None, T_21=None, device=0):
    """FrameToFrameRANSAC on the device (csrc/ransac.cuh).  Either `sample_idx` [n_hyp, n_min] + `pts_2` (the
    hypotheses are computed from the minimal sets) or `T_21` [n_hyp, 4, 4].  Returns (T_21 [n_hyp,4,4],
    inlier counts [n_hyp], index of the first best hypothesis, its inlier mask [n_pts] bool)."""
    lib = load_library()
    pts_1, obs_2 = _f64(pts_1, (-1, 3)), _f64(obs_2, (-1, 3))
    n_pts = len(pts_1)
    if sample_idx is not None:
        idx = _i32(sample_idx)
        n_hyp, n_min = idx.shape
        p2, Tin = _f64(pts_2, (-1, 3)), None
    else:
        Tin = _f64(T_21, (-1, 16))
        n_hyp, n_min, idx, p2 = len(Tin), 0, None, None
    T_out = np.empty((n_hyp, 16))
    counts, best, mask = np.zeros(n_hyp, np.int32), np.zeros(2, np.int32), np.zeros(n_pts, np.uint8)
    intr = _f64(intr, (5,))
    _ck_global(lib, lib.bslam_ransac(int(device), n_hyp, n_min, _i(idx), _d(Tin), n_pts, _d(pts_1), _d(p2), _d(obs_2), _d(intr),
                                     float(thresh), _d(T_out), _i(counts), _i(best), _b(mask)))
    return T_out.reshape(n_hyp, 4, 4), counts, int(best[0]), mask.astype(bool)


def _level_shapes(w, h, levels):
    out = []
    for _ in range(levels):
        out.append((h, w))
        w, h = (w + 1) // 2, (h + 1) // 2
    return out


def image_pyramid(image_u8, levels, gradients=True, device=0):
    """cv2.pyrDown chain of an 8-bit image as float / 255 per level, with 0.5 * Sobel gradients per level
    (csrc/image.cuh).  Returns (list of images, list of [gradx, grady] arrays or None)."""
    lib = load_library()
    im = np.ascontiguousarray(image_u8, dtype=np.uint8)
    h, w = im.shape
    shapes = _level_shapes(w, h, levels)
    total = sum(a * b for a, b in shapes)
    out = np.empty(total)
    gx, gy = (np.empty(total), np.empty(total)) if gradients else (None, None)
    _ck_global(lib, lib.bslam_image_pyramid(int(device), _b(im.ravel()), w, h, int(levels), _d(out), _d(gx), _d(gy)))
    ims, jac, off = [], [], 0
    for hh, ww in shapes:
        n = hh * ww
        ims.append(out[off:off + n].reshape(hh, ww).copy())
        if gradients:
            jac.append(np.array([gx[off:off + n].reshape(hh, ww), gy[off:off + n].reshape(hh, ww)]))
        off += n
    return ims, (jac if gradients else None)


def subsample_pyramid(depth_map, levels, scale_per_level=1., device=0):
    """Level l = map[::2^l, ::2^l] * scale_per_level^l (disparity: 0.5; depth: 1)."""
    lib = load_library()
    m = np.ascontiguousarray(depth_map, dtype=np.float64)
    h, w = m.shape
    shapes = _level_shapes(w, h, levels)
    out = np.empty(sum(a * b for a, b in shapes))
    _ck_global(lib, lib.bslam_subsample_pyramid(int(device), _d(m.ravel()), w, h, int(levels), float(scale_per_level), _d(out)))
    res, off = [], 0
    for hh, ww in shapes:
        res.append(out[off:off + hh * ww].reshape(hh, ww).copy())
        off += hh * ww
    return res

"""TEST DOUBLE for pyslam_b200.engine.Engine, backed by the CPU oracle.

Lets the CPU test-suite (-m "not gpu") exercise the host logic of
pyslam_b200.Problem -- lowering, update-vector ordering, termination rules,
plug-in blocks, parameter write-back -- without a GPU.  It lives under tests/
and is never importable from the product."""
import numpy as np

from oracle import gn_oracle as O
from oracle import liegroups as OL

N_SCALARS = 16


def _se3(row):
    return OL.SE3(OL.SO3(row[:9].reshape(3, 3).copy()), row[9:].copy())


def _se2(row):
    return OL.SE2(OL.SO2(row[:4].reshape(2, 2).copy()), row[4:].copy())


def _row(T):
    return np.concatenate([T.rot.mat.ravel(), T.trans])


class FakeEngine:
    def __init__(self, device=0):
        self.n = dict(se3=0, se2=0, pt=0, vec=0, vec_entries=0)
        self.tab = dict(se3=np.zeros((0, 12)), se2=np.zeros((0, 6)), pt=np.zeros((0, 3)), vec=np.zeros(0))
        self.const = dict(se3=np.zeros(0, bool), se2=np.zeros(0, bool), pt=np.zeros(0, bool), vec=np.zeros(0, bool))
        self.vec_dims = np.zeros(0, int)
        self.finalized = False
        self.clear_blocks()
        self._scal = np.zeros(N_SCALARS)
        self.calls = []

    # ---- tables
    def _set(self, name, vals, is_const, width):
        vals = np.array(vals, dtype=float).reshape(-1, width)
        if self.finalized:
            assert len(vals) == self.n[name]
        else:
            self.n[name] = len(vals)
            self.const[name] = np.zeros(len(vals), bool) if is_const is None else np.asarray(is_const, bool)
        self.tab[name] = vals

    def set_poses_se3(self, Rt, is_const=None):
        self._set('se3', Rt, is_const, 12)

    def set_poses_se2(self, Rt, is_const=None):
        self._set('se2', Rt, is_const, 6)

    def set_points(self, xyz, is_const=None):
        self._set('pt', xyz, is_const, 3)

    def set_vectors(self, dims, values, is_const=None):
        dims = np.asarray(dims, int)
        if not self.finalized:
            self.vec_dims = dims
            self.n['vec'], self.n['vec_entries'] = len(dims), int(dims.sum())
            self.const['vec'] = np.zeros(len(dims), bool) if is_const is None else np.asarray(is_const, bool)
        self.tab['vec'] = np.array(values, dtype=float).ravel()

    def get_poses_se3(self):
        return self.tab['se3'].copy()

    def get_poses_se2(self):
        return self.tab['se2'].copy()

    def get_points(self, out=None):
        if out is not None:
            out[...] = self.tab['pt']
            return out
        return self.tab['pt'].copy()

    def get_vectors(self):
        return self.tab['vec'].copy()

    # ---- blocks
    def clear_blocks(self):
        self.reproj, self.pose_b, self.p2p_b, self.photo_b, self.dense = [], [], [], [], None
        self.finalized = False

    @staticmethod
    def _loss(kind, k):
        return O.loss_from_kind(['l2', 'l1', 'cauchy', 'huber', 'tukey', 'tdist'][kind], k)

    def add_reprojection_blocks(self, pose_idx, pt_idx, obs, stiffness, intr, loss_kind=0, loss_k=0.):
        S = np.asarray(stiffness, float)
        n = len(pose_idx)
        S = np.broadcast_to(S.reshape(-1, 3, 3), (n, 3, 3)) if S.size in (9, 9 * n) else None
        cam = O.StereoCamera(*intr, 1, 1)
        for i in range(n):
            self.reproj.append((int(pose_idx[i]), int(pt_idx[i]), np.asarray(obs, float).reshape(-1, 3)[i], S[i], cam,
                                self._loss(loss_kind, loss_k)))

    def add_pose_blocks(self, group, pose_idx, T_obs, stiffness, loss_kind=0, loss_k=0.):
        d = 6 if group == 3 else 3
        n = len(pose_idx)
        S = np.broadcast_to(np.asarray(stiffness, float).reshape(-1, d, d), (n, d, d))
        T = np.asarray(T_obs, float).reshape(n, -1)
        for i in range(n):
            self.pose_b.append((group, int(pose_idx[i]), T[i], S[i], self._loss(loss_kind, loss_k)))

    def add_pose_to_pose_blocks(self, group, idx1, idx2, T21_obs, stiffness, loss_kind=0, loss_k=0.):
        d = 6 if group == 3 else 3
        n = len(idx1)
        S = np.broadcast_to(np.asarray(stiffness, float).reshape(-1, d, d), (n, d, d))
        T = np.asarray(T21_obs, float).reshape(n, -1)
        for i in range(n):
            self.p2p_b.append((group, int(idx1[i]), int(idx2[i]), T[i], S[i], self._loss(loss_kind, loss_k)))

    def add_photometric_block(self, pose_idx, uvd_ref, im_ref, im_jac, im_track, intr, intensity_stiffness,
                              depth_stiffness, loss_kind=0, loss_k=0.):
        im_track = np.asarray(im_track, float)
        blk = O.PhotometricResidualSE3.__new__(O.PhotometricResidualSE3)
        blk.camera = O.StereoCamera(*intr, im_track.shape[1], im_track.shape[0])
        blk.uvd_ref, blk.im_ref = np.asarray(uvd_ref, float).reshape(-1, 3), np.asarray(im_ref, float).ravel()
        blk.im_jac, blk.im_track = np.asarray(im_jac, float).reshape(-1, 2), im_track
        blk.intensity_covar, blk.depth_covar = intensity_stiffness ** -2, depth_stiffness ** -2
        blk.pt_ref, blk.triang_jac = blk.camera.triangulate(blk.uvd_ref, compute_jacobians=True)
        self.photo_b.append((int(pose_idx), blk, self._loss(loss_kind, loss_k)))

    def set_dense_blocks(self, rows, param_ptr, param_kind, param_index):
        self.dense = (list(rows), list(param_ptr), list(param_kind), list(param_index))
        self.dense_vals = None

    def upload_dense_values(self, e, J, cost):
        self.dense_vals = (np.array(e, float).ravel(), np.array(J, float).ravel(), float(cost))

    # ---- layout: SE3 | SE2 | vectors | points (any consistent layout is legal)
    def finalize(self):
        off = 0
        self.off = {}
        for name, dof in (('se3', 6), ('se2', 3)):
            o = np.full(self.n[name], -1, np.int32)
            for i in range(self.n[name]):
                if not self.const[name][i]:
                    o[i] = off
                    off += dof
            self.off[name] = o
        o = np.full(self.n['vec'], -1, np.int32)
        for i in range(self.n['vec']):
            if not self.const['vec'][i]:
                o[i] = off
                off += int(self.vec_dims[i])
        self.off['vec'] = o
        self.n_reduced = off
        # points last and in REVERSE order, to prove the host does not assume an ordering
        o = np.full(self.n['pt'], -1, np.int32)
        for i in reversed(range(self.n['pt'])):
            if not self.const['pt'][i]:
                o[i] = off
                off += 3
        self.off['pt'] = o
        self.dim = off
        self.dx = np.zeros(off)
        self.finalized = True

    def layout(self):
        out = {k: v.copy() for k, v in self.off.items()}
        out['dim'], out['n_reduced'] = self.dim, self.n_reduced
        return out

    # ---- maths (oracle)
    def _objects(self):
        return ([_se3(r) for r in self.tab['se3']], [_se2(r) for r in self.tab['se2']])

    def _blocks(self):
        """yield (residual object, [(table, index, param object)], loss)"""
        se3, se2 = self._objects()
        for ci, qi, obs, S, cam, loss in self.reproj:
            yield O.ReprojectionResidual(cam, obs, S), [('se3', ci, se3[ci]), ('pt', qi, self.tab['pt'][qi])], loss
        for grp, i, T, S, loss in self.pose_b:
            name, objs, mk = ('se3', se3, _se3) if grp == 3 else ('se2', se2, _se2)
            yield O.PoseResidual(mk(T), S), [(name, i, objs[i])], loss
        for grp, i, j, T, S, loss in self.p2p_b:
            name, objs, mk = ('se3', se3, _se3) if grp == 3 else ('se2', se2, _se2)
            yield O.PoseToPoseResidual(mk(T), S), [(name, i, objs[i]), (name, j, objs[j])], loss
        for i, blk, loss in self.photo_b:
            yield blk, [('se3', i, se3[i])], loss

    def eval_cost(self):
        return float(sum(np.sum(loss.loss(blk.evaluate([p for _, _, p in ps]))) for blk, ps, loss in self._blocks()))

    def _assemble(self):
        D = self.dim
        H, b, cost = np.zeros((D, D)), np.zeros(D), 0.
        dofs = dict(se3=6, se2=3, pt=3)
        for blk, ps, loss in self._blocks():
            cj = [self.off[n][i] >= 0 for n, i, _ in ps]
            if not any(cj):
                continue
            r, jac = blk.evaluate([p for _, _, p in ps], cj)
            sw = np.sqrt(loss.weight(r))
            J = np.zeros((r.size, D))
            for (n, i, _), want, j in zip(ps, cj, jac):
                if want:
                    o = self.off[n][i]
                    J[:, o:o + dofs[n]] = sw[:, None] * j
            H += J.T @ J
            b -= J.T @ (sw * r)
            cost += float(np.sum(loss.loss(r)))
        if self.dense is not None and self.dense[0]:
            assert self.dense_vals is not None, 'dense values not uploaded'
            rows, pptr, pkind, pindex = self.dense
            e, Jv, c = self.dense_vals
            names = ['se3', 'se2', 'pt', 'vec']
            ro, jo = 0, 0
            for bi, m in enumerate(rows):
                cols = []
                for k in range(pptr[bi], pptr[bi + 1]):
                    n, i = names[pkind[k]], pindex[k]
                    dof = int(self.vec_dims[i]) if n == 'vec' else dofs[n]
                    o = self.off[n][i]
                    cols += [(o + c if o >= 0 else -1) for c in range(dof)]
                Jb = Jv[jo:jo + m * len(cols)].reshape(m, len(cols))
                J = np.zeros((m, D))
                for c, gi in enumerate(cols):
                    if gi >= 0:
                        J[:, gi] += Jb[:, c]
                H += J.T @ J
                b -= J.T @ e[ro:ro + m]
                ro += m
                jo += m * len(cols)
            cost += c
            self.dense_vals = None
        return H, b, cost

    def linearize(self, fetch_cost=True):
        self.H, self.b, c = self._assemble()
        self._scal[0] = c
        return c

    def _retract(self):
        dx = self.dx
        for name, mk, dof in (('se3', _se3, 6), ('se2', _se2, 3)):
            for i in range(self.n[name]):
                o = self.off[name][i]
                if o >= 0:
                    T = mk(self.tab[name][i])
                    T.perturb(dx[o:o + dof])
                    self.tab[name][i] = _row(T)
        for i in range(self.n['pt']):
            o = self.off['pt'][i]
            if o >= 0:
                self.tab['pt'][i] += dx[o:o + 3]
        pos = 0
        for i in range(self.n['vec']):
            d = int(self.vec_dims[i])
            o = self.off['vec'][i]
            if o >= 0:
                self.tab['vec'][pos:pos + d] += dx[o:o + d]
            pos += d

    def iterate(self, lam=0., eval_new_cost=True):
        self.calls.append('iterate')
        c = self.linearize()
        H = self.H + lam * np.diag(np.diag(self.H))
        self.dx = np.linalg.solve(H, self.b) if self.dim else np.zeros(0)
        self._retract()
        cn = self.eval_cost() if eval_new_cost else 0.
        self._scal[:4] = [c, cn, float(self.dx @ self.dx), 0.]
        return c, cn, float(np.linalg.norm(self.dx))

    # ---- phase-wise iteration (multi-GPU plumbing) ----
    def set_shard(self, rank):
        self.rank = rank

    def tile_structure(self):
        return np.ones((2, 1), np.uint8)

    def merge_tile_structure(self, mask):
        self.merged_mask = np.array(mask)

    def reduce(self, lam=0.):
        """Schur complement of the points out of (H, b) -> torch buffer [S | rhs | scalars]."""
        import torch
        n = self.n_reduced
        H = self.H + lam * np.diag(np.diag(self.H))
        Hcc, Hcp, Hpp = H[:n, :n], H[:n, n:], H[n:, n:]
        self._Hpp, self._Hpc, self._bp = Hpp, Hcp.T, self.b[n:]
        if Hpp.size:
            X = np.linalg.solve(Hpp, np.column_stack([Hcp.T, self.b[n:]]))
            S = Hcc - Hcp @ X[:, :n]
            rhs = self.b[:n] - Hcp @ X[:, n]
        else:
            S, rhs = Hcc.copy(), self.b[:n].copy()
        buf = np.concatenate([S.ravel(), rhs, np.zeros(N_SCALARS)])
        buf[n * n + n] = self._scal[0]
        self._buf = torch.from_numpy(buf)

    def reduced_tensor(self):
        return self._buf

    # the product engine's sharded-iteration surface (bslam_iterate_pre / bslam_packed_buffer / bslam_iterate_post):
    # here the dense buffer doubles as the "packed" payload
    def iterate_pre(self, lam=0.):
        self.linearize(fetch_cost=False)
        self.reduce(lam)

    def packed_tensor(self):
        return self._buf

    def iterate_post(self, eval_new_cost=True):
        self.solve_reduced()
        self.retract(eval_new_cost)

    def scalars_tensor(self):
        import torch
        if getattr(self, '_buf', None) is None:
            return torch.zeros(N_SCALARS, dtype=torch.float64)
        n = self.n_reduced
        return self._buf[n * n + n:]

    def torch_stream(self):
        return None

    def solve_reduced(self):
        n = self.n_reduced
        buf = self._buf.numpy()
        S, rhs = buf[:n * n].reshape(n, n), buf[n * n:n * n + n]
        dxc = np.linalg.solve(S, rhs) if n else np.zeros(0)
        dxp = np.linalg.solve(self._Hpp, self._bp - self._Hpc @ dxc) if self._Hpp.size else np.zeros(0)
        self.dx = np.concatenate([dxc, dxp])

    def retract(self, eval_new_cost=True):
        n = self.n_reduced
        self._retract()
        tail = self.scalars_tensor().numpy()
        tail[1] = self.eval_cost() if eval_new_cost else 0.
        tail[2] = float(self.dx[n:] @ self.dx[n:]) + (float(self.dx[:n] @ self.dx[:n]) if getattr(self, 'rank', 0) == 0 else 0.)

    def scalars(self):
        if getattr(self, '_buf', None) is not None:
            return self.scalars_tensor().numpy().copy()
        return self._scal.copy()

    def snapshot(self):
        self.calls.append('snapshot')
        self._snap = {k: v.copy() for k, v in self.tab.items()}

    def restore(self):
        self.calls.append('restore')
        self.tab = {k: v.copy() for k, v in self._snap.items()}

    def get_update(self, dim):
        return self.dx.copy()

    def get_normal_equations(self, dim):
        return self.H.copy(), self.b.copy()

    def covariance(self, dim):
        H, _, _ = self._assemble()
        return np.linalg.inv(H)

    def enable_timing(self, on=True):
        pass

    def timings(self):
        return {}

    def close(self):
        pass

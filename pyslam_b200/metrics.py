"""Trajectory error metrics on SE2 / SE3 pose lists -- same class, methods, arguments and .mat format as
reference pyslam/metrics.py:7-300 (`poses_gt` / `poses_est` stored M x M x N, metrics.py:95-110).

Evaluation-format code, not on the hot path: re-expressed on stacked homogeneous matrices (one batched numpy
operation per metric instead of a Python loop of liegroups products).
"""
import numpy as np

from .lie import SE2, SE3, SO2, SO3


def _stack(poses):
    return np.array([T.as_matrix() for T in poses])


def _inv(M):
    """Batched inverse of rigid transforms [N, d+1, d+1]."""
    R, t = M[:, :-1, :-1], M[:, :-1, -1]
    out = np.zeros_like(M)
    out[:, :-1, :-1] = np.transpose(R, (0, 2, 1))
    out[:, :-1, -1] = -np.einsum('nji,nj->ni', R, t)
    out[:, -1, -1] = 1.
    return out


def _rot_log(R):
    """SO2 / SO3 logarithm of every rotation in the stack (liegroups semantics incl. the small-angle branch)."""
    if R.shape[-1] == 2:
        return np.array([[SO2(r).log()] for r in R]).reshape(len(R), 1)
    return np.array([SO3(r).log() for r in R]).reshape(len(R), 3)


class TrajectoryMetrics:
    """convention='Twv': poses are vehicle-to-world transforms; 'Tvw': world-to-vehicle (converted to Twv internally)."""

    def __init__(self, poses_gt, poses_est, convention='Twv'):
        if convention == 'Twv':
            Twv_gt, Twv_est = list(poses_gt), list(poses_est)
        elif convention == 'Tvw':
            Twv_gt, Twv_est = [T.inv() for T in poses_gt], [T.inv() for T in poses_est]
        else:
            raise ValueError('convention must be \'Tvw\' or \'Twv\'')
        if len(Twv_gt) != len(Twv_est):
            n = min(len(Twv_gt), len(Twv_est))
            print('WARNING: poses_gt has length {} but poses_est has length {}. Truncating to {}.'.format(
                len(Twv_gt), len(Twv_est), n))
            Twv_gt, Twv_est = Twv_gt[:n], Twv_est[:n]
        self.convention = convention
        self.Twv_gt, self.Twv_est = Twv_gt, Twv_est
        self.pose_type = type(Twv_gt[0])
        self.num_poses = len(Twv_gt)
        self._gt, self._est = _stack(Twv_gt), _stack(Twv_est)
        self.rel_dists, self.cum_dists = self._compute_distances()

    def _compute_distances(self):
        pos = np.zeros((self.num_poses, 3))
        d = self._gt.shape[1] - 1
        pos[:, :d] = self._gt[:, :-1, -1]
        rel = np.append([0.], np.linalg.norm(np.diff(pos, axis=0), axis=1))
        return rel, np.cumsum(rel)

    @staticmethod
    def _convert_meters(meters, unit):
        return {'m': 1., 'dm': 10., 'cm': 100., 'mm': 1000.}[unit] * meters

    @staticmethod
    def _convert_radians(radians, unit):
        return {'rad': 1, 'deg': 180. / np.pi}[unit] * radians

    def savemat(self, filename, extras=None):
        import scipy.io
        gt, est = (self._gt, self._est) if self.convention == 'Twv' else (_inv(self._gt), _inv(self._est))
        mdict = {'poses_gt': np.transpose(gt, [1, 2, 0]), 'poses_est': np.transpose(est, [1, 2, 0]),
                 'convention': self.convention, 'pose_type': self.pose_type.__name__, 'num_poses': self.num_poses,
                 'rel_dists': self.rel_dists, 'cum_dists': self.cum_dists}
        if extras is not None:
            mdict.update(extras)
        scipy.io.savemat(filename, mdict, do_compression=True)

    @classmethod
    def loadmat(cls, filename):
        import scipy.io
        mdict = scipy.io.loadmat(filename, verify_compressed_data_integrity=True)
        name = str(np.squeeze(mdict['pose_type']))
        if name not in ('SE2', 'SE3'):
            raise ValueError('Got invalid pose type: {}'.format(mdict['pose_type']))
        pose_type = SE2 if name == 'SE2' else SE3
        n = int(np.squeeze(mdict['num_poses']))
        gt = [pose_type.from_matrix(mdict['poses_gt'][:, :, i], normalize=True) for i in range(n)]
        est = [pose_type.from_matrix(mdict['poses_est'][:, :, i], normalize=True) for i in range(n)]
        tm = cls(gt, est, convention=str(np.squeeze(mdict['convention'])))
        tm.mdict = mdict
        return tm

    def _errors(self, err, trans_unit, rot_unit):
        return (self._convert_meters(err[:, :-1, -1], trans_unit),
                self._convert_radians(_rot_log(err[:, :-1, :-1]), rot_unit))

    def endpoint_error(self, segment_range=None, trans_unit='m', rot_unit='rad'):
        """Translational and rotational error at the endpoint of a segment."""
        if segment_range is None:
            segment_range = range(self.num_poses)
        a, b = segment_range[0], segment_range[-1]
        d_gt = _inv(self._gt[a:a + 1]) @ self._gt[b:b + 1]
        d_est = _inv(self._est[a:a + 1]) @ self._est[b:b + 1]
        t, r = self._errors(_inv(d_est) @ d_gt, trans_unit, rot_unit)
        return np.linalg.norm(t[0]), np.linalg.norm(r[0])

    def segment_errors(self, segment_lengths, trans_unit='m', rot_unit='rad'):
        """All endpoint errors of all segments of the given lengths, and their averages per length:
        rows (length, proportional translation error, proportional rotation error)."""
        errs = []
        for length in segment_lengths:
            length = self._convert_meters(length, trans_unit)
            for start in range(self.num_poses):
                stop = np.searchsorted(self.cum_dists - self.cum_dists[start], length, side='right')
                if stop < self.num_poses:
                    te, re = self.endpoint_error(range(start, stop + 1), trans_unit, rot_unit)
                    errs.append([length, te / length, re / length])
        errs = np.array(errs)
        avg = np.array([np.mean(errs[errs[:, 0] == self._convert_meters(l, trans_unit)], axis=0) for l in segment_lengths])
        return errs, avg

    def traj_errors(self, segment_range=None, trans_unit='m', rot_unit='rad'):
        """Errors in all degrees of freedom relative to the first ground-truth pose of the segment."""
        if segment_range is None:
            segment_range = range(self.num_poses)
        idx = np.asarray(list(segment_range))
        g0 = _inv(self._gt[idx[0]:idx[0] + 1])
        d_gt, d_est = g0 @ self._gt[idx], g0 @ self._est[idx]
        return self._errors(_inv(d_est) @ d_gt, trans_unit, rot_unit)

    def rel_errors(self, segment_range=None, trans_unit='m', rot_unit='rad', delta=1):
        """Relative pose errors (Sturm et al., eq. 1)."""
        if segment_range is None:
            segment_range = range(self.num_poses)
        idx = np.asarray(list(segment_range))[:-delta]
        d_gt = _inv(self._gt[idx]) @ self._gt[idx + delta]
        d_est = _inv(self._est[idx]) @ self._est[idx + delta]
        return self._errors(_inv(d_gt) @ d_est, trans_unit, rot_unit)

    def error_norms(self, segment_range=None, trans_unit='m', rot_unit='rad', error_type='traj', delta=1):
        if error_type == 'traj':
            t, r = self.traj_errors(segment_range, trans_unit, rot_unit)
        elif error_type == 'rel':
            t, r = self.rel_errors(segment_range, trans_unit, rot_unit, delta)
        else:
            raise ValueError('error_type must be either `traj` or `rel`.')
        return np.sqrt(np.sum(t ** 2, axis=1)), np.sqrt(np.sum(r ** 2, axis=1))

    def mean_err(self, segment_range=None, trans_unit='m', rot_unit='rad', error_type='traj'):
        t, r = self.error_norms(segment_range, trans_unit, rot_unit, error_type)
        return np.mean(t), np.mean(r)

    def cum_err(self, segment_range=None, trans_unit='m', rot_unit='rad', error_type='traj'):
        t, r = self.error_norms(segment_range, trans_unit, rot_unit, error_type)
        return np.cumsum(t), np.cumsum(r)

    def rms_err(self, segment_range=None, trans_unit='m', rot_unit='rad', error_type='traj', delta=1):
        t, r = self.error_norms(segment_range, trans_unit, rot_unit, error_type, delta)
        return np.sqrt(np.mean(t ** 2)), np.sqrt(np.mean(r ** 2))

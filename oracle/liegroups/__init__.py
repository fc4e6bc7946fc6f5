"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the `liegroups` numpy backend.

pyslam's parameter type and all of its manifold arithmetic come from the
third-party package `liegroups` (github.com/utiasSTARS/liegroups), which is an
UNPINNED dependency (`/root/reference/setup.py:12`), is not vendored under
/root/reference and is not installable here (no network).  This module restates
the handful of methods pyslam's hot path calls, from upstream's published
formulas.  It exists so that

  * the unmodified reference can be imported in the build container
    (`oracle/reference_loader.py`) to generate golden vectors, and
  * the numpy oracle (`oracle/gn_oracle.py`) has the same manifold semantics.

Parity status: the *exact* numeric outputs of exp/log/adjoint are not pinned by
any golden vector in /root/reference; they are pinned indirectly by the
reference's own tests (tests/test_problem.py:163-198, 239-282, 294-321 and
tests/test_costs.py), all 36 of which pass against this restatement, and by
pyslam's own element-wise restatement of SE3.odot
(pyslam/residuals/photometric_residual.py:14-35).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
arm may import this package.  The product (pyslam_b200/) never does.
"""
from .so2 import SO2
from .se2 import SE2
from .so3 import SO3
from .se3 import SE3

__all__ = ['SO2', 'SE2', 'SO3', 'SE3']

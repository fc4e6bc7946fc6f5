"""pytest plugin (-p reference_alias_plugin): make `import pyslam...` / `import liegroups` resolve to the PRODUCT
(pyslam_b200, pyslam_b200.lie), so that the reference's own test files run unmodified against libbslam.so.
Used by tests/test_reference_suite.py only."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import pyslam_b200  # noqa: E402
import pyslam_b200.lie  # noqa: E402
import pyslam_b200.losses  # noqa: E402
import pyslam_b200.metrics  # noqa: E402
import pyslam_b200.pipelines  # noqa: E402
import pyslam_b200.problem  # noqa: E402
import pyslam_b200.residuals  # noqa: E402
import pyslam_b200.sensors  # noqa: E402
import pyslam_b200.utils  # noqa: E402

sys.modules['liegroups'] = pyslam_b200.lie
sys.modules['pyslam'] = pyslam_b200
for _name in ('problem', 'residuals', 'losses', 'sensors', 'utils', 'pipelines', 'metrics'):
    sys.modules['pyslam.' + _name] = getattr(pyslam_b200, _name)

"""GPU: the reference's OWN test files (tests/test_problem.py, test_costs.py, test_sensors.py, test_utils.py of
utiasSTARS/pyslam), collected unmodified, with `pyslam` -> pyslam_b200 and `liegroups` -> pyslam_b200.lie
(tests/reference_alias_plugin.py).  Every Problem.solve() in them runs through libbslam.so on the GPU.

The files are read from /root/reference/tests (build container) or baseline/_ref/tests (a git-ignored verbatim copy
made by oracle/install_reference.sh that travels to the GPU box); nothing of them is committed."""
import os
import re
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _reference_tests():
    for d in ('/root/reference/tests', os.path.join(ROOT, 'baseline', '_ref', 'tests')):
        if os.path.isfile(os.path.join(d, 'test_problem.py')):
            return d
    return None


@pytest.mark.timeout(900)
def test_reference_test_files_pass_against_the_product(tmp_path):
    d = _reference_tests()
    if d is None:
        pytest.skip('reference test files not present (run oracle/install_reference.sh in the build container)')
    env = dict(os.environ)
    env['PYTHONPATH'] = os.pathsep.join([os.path.join(ROOT, 'tests'), ROOT, env.get('PYTHONPATH', '')])
    # a scratch rootdir + cache dir: the reference tree is read-only and must not pick up this repo's conftest
    r = subprocess.run([sys.executable, '-m', 'pytest', d, '-q', '-p', 'reference_alias_plugin', '-p', 'no:cacheprovider',
                        '--rootdir', str(tmp_path), '-W', 'ignore'], capture_output=True, text=True, env=env, cwd=str(tmp_path))
    tail = (r.stdout + r.stderr)[-3000:]
    m = re.search(r'(\d+) passed', r.stdout)
    assert r.returncode == 0, tail
    assert m and int(m.group(1)) >= 36, tail
    assert 'failed' not in r.stdout.splitlines()[-1], tail

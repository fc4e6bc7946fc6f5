"""TEST INFRASTRUCTURE -- import the UNMODIFIED reference (/root/reference).

Used only in the build container (where /root/reference exists) by
oracle/make_golden.py and by the optional "reference present" CPU tests.  The
GPU box has no /root/reference; nothing on the -m gpu / smoke / bench path
calls this.

Three shims make the verbatim tree import on py3.12 / numpy 2.3:
  1. oracle/ on sys.path so `import liegroups` resolves to oracle/liegroups;
  2. numpy.int = int                      (pyslam/utils.py:44,47);
  3. FileFinder.find_module restored      (pyslam/residuals/__init__.py:7);
plus a writable NUMBA_CACHE_DIR (the kernels use cache=True and
/root/reference is read-only).
"""
import importlib.machinery
import os
import sys
import tempfile

REFERENCE_ROOT = os.environ.get('PYSLAM_REFERENCE_ROOT', '/root/reference')


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'pyslam', 'problem.py'))


def load_reference():
    """Returns the imported verbatim `pyslam` package."""
    if not reference_available():
        raise RuntimeError('reference tree not present at ' + REFERENCE_ROOT)
    os.environ.setdefault('NUMBA_CACHE_DIR',
                          os.path.join(tempfile.gettempdir(), 'numba_cache_pyslam_ref'))
    here = os.path.dirname(os.path.abspath(__file__))
    if here not in sys.path:
        sys.path.insert(0, here)           # -> `import liegroups` = oracle/liegroups
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import numpy as np
    if not hasattr(np, 'int'):
        np.int = int
    FF = importlib.machinery.FileFinder
    if not hasattr(FF, 'find_module'):
        def find_module(self, name):
            spec = self.find_spec(name)
            return spec.loader if spec is not None else None
        FF.find_module = find_module
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        import pyslam
        import pyslam.problem
        import pyslam.losses
        import pyslam.utils
        import pyslam.sensors
        import pyslam.residuals
    return pyslam

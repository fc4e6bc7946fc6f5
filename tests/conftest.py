import os
import sys

import numpy as np
import pytest

# several solver handles (one CUDA stream each) of one process rendezvous on the device in tests/test_gpu_multi.py:
# more hardware queues than the default 8, so that two streams never share one (read at CUDA context creation)
os.environ.setdefault('CUDA_DEVICE_MAX_CONNECTIONS', '32')

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')
    config.addinivalue_line('markers', 'real_engine: do not replace the engine with the CPU test double')


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False)


@pytest.fixture
def golden():
    return load_golden


def rel_err(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))

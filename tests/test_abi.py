"""CPU: the C-ABI shared library loads and exports exactly the entry points
declared in include/bslam.h, and the ctypes binding covers all of them.  No
compute call is made (there is no GPU here)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from pyslam_b200 import engine as E

HEADER = os.path.join(ROOT, 'include', 'bslam.h')


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r'^BSLAM_API\s+[\w\s\*]+?\b(bslam_\w+)\s*\(', src, flags=re.M)))


def test_header_declares_the_expected_surface():
    names = declared_symbols()
    assert len(names) >= 35
    for must in ('bslam_create', 'bslam_destroy', 'bslam_add_reprojection_blocks', 'bslam_add_pose_blocks',
                 'bslam_add_pose_to_pose_blocks', 'bslam_iterate', 'bslam_eval_cost', 'bslam_get_update'):
        assert must in names


def test_library_exports_every_declared_symbol():
    if not os.path.isfile(E.LIB_PATH):
        pytest.fail('libbslam.so has not been built: run __graft_entry__.build()')
    lib = ctypes.CDLL(E.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert lib.bslam_version() >= 100


def test_binding_covers_header_exactly():
    assert sorted(E.SIGNATURES) == declared_symbols()
    E.load_library()


def test_header_constants_match_binding():
    src = open(HEADER).read()
    defs = dict(re.findall(r'#define\s+(BSLAM_\w+)\s+(-?\d+)', src))
    assert int(defs['BSLAM_N_SCALARS']) == E.N_SCALARS
    assert int(defs['BSLAM_N_TIMINGS']) == E.N_TIMINGS
    assert int(defs['BSLAM_SE2']) == E.SE2 and int(defs['BSLAM_SE3']) == E.SE3
    from pyslam_b200 import losses as L
    for k, v in dict(L2=L.LOSS_L2, L1=L.LOSS_L1, CAUCHY=L.LOSS_CAUCHY, HUBER=L.LOSS_HUBER, TUKEY=L.LOSS_TUKEY,
                     TDIST=L.LOSS_TDIST).items():
        assert int(defs['BSLAM_LOSS_' + k]) == v
    for i, name in enumerate(E.TIMING_NAMES):
        assert int(defs['BSLAM_T_' + name.upper()]) == i


def test_create_without_gpu_reports_error():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    lib = E.load_library()
    h = ctypes.c_void_p()
    rc = lib.bslam_create(ctypes.byref(h), 0)
    assert rc == -2 and not h.value
    assert b'no CPU fallback' in lib.bslam_last_error(None)

"""CPU: the C-ABI library builds, loads and exports every symbol include/b200trk.h declares; argument
validation errors are reported through status codes + b200trk_last_error (no exit(), no device needed)."""
import ctypes as C
import os
import re

import pytest

from pytracking_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "b200trk.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(b200trk_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    h = C.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(h, n), "missing export: " + n
    assert sorted(_lib.SIGNATURES) == names, "ctypes table out of sync with include/b200trk.h"


def test_version_and_error_reporting():
    l = _lib.lib()
    assert l.b200trk_version() == 100
    # invalid arguments are rejected before any CUDA call
    st = l.b200trk_apply_filter(None, None, None, 1, 512, 18, 18, 4, None, None, None)
    assert st != 0 and b"null" in l.b200trk_last_error()
    one = C.c_void_p(16)
    st = l.b200trk_apply_filter(one, one, one, 1, 512, 18, 18, 3, None, None, None)
    assert st != 0 and b"filter size" in l.b200trk_last_error()
    st = l.b200trk_apply_filter(one, one, one, 0, 512, 18, 18, 4, None, None, None)
    assert st != 0 and b"empty" in l.b200trk_last_error()
    st = l.b200trk_dimp_sd_gn(one, one, one, one, None, 5, 512, 19, 19, 4, 2, one, one, one, 100, 0.1, 16.0, 1.0, 0.01, 0.0,
                              None, None, None)
    assert st != 0 and b"feature size" in l.b200trk_last_error()
    with pytest.raises(RuntimeError):
        _lib.check(st, "dimp_sd_gn")


def test_ops_refuse_cpu_tensors():
    import torch
    from pytracking_b200 import ops
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.apply_filter(torch.zeros(1, 512, 18, 18), torch.zeros(1, 512, 4, 4))

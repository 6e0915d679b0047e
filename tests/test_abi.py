"""CPU tests of the drop-in boundary: the C-ABI shared library builds, loads and exports every symbol
that include/unipose_b200.h declares; argument validation fails loudly without touching a GPU."""
import ctypes
import os
import re

import pytest

from unipose_b200 import _lib

from conftest import ROOT


def _declared_in_header():
    src = open(os.path.join(ROOT, "include", "unipose_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(up_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_loads():
    lib = _lib.load()
    assert lib.up_version() == 100


def test_every_declared_symbol_is_exported():
    lib = _lib.load()
    declared = _declared_in_header()
    assert len(declared) >= 20
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, "declared in include/unipose_b200.h but not exported: %s" % missing
    # and the Python binding table covers the header
    unbound = [s for s in declared if s not in _lib.DECLARED_SYMBOLS]
    assert not unbound, "no ctypes signature for: %s" % unbound


def test_conv_desc_struct_layout_matches_header():
    # 28 int32 + 4 int64, no padding surprises
    assert ctypes.sizeof(_lib.UpConvDesc) == 28 * 4 + 4 * 8 + 2 * 4
    assert _lib.UpConvDesc.x_plane_stride.offset == 28 * 4


def test_argument_validation_reports_errors():
    lib = _lib.load()
    d = _lib.UpConvDesc()
    rc = lib.up_conv2d_fwd(ctypes.byref(d), None, None, None, None, None, None, None, None)
    assert rc == -1
    assert b"null" in lib.up_last_error()
    with pytest.raises(RuntimeError, match="up_conv2d_fwd failed"):
        _lib.call("up_conv2d_fwd", ctypes.byref(d), None, None, None, None, None, None, None, None)
    with pytest.raises(RuntimeError, match="bad"):
        _lib.call("up_argmax2d", ctypes.c_void_p(8), ctypes.c_void_p(8), ctypes.c_void_p(8), ctypes.c_void_p(8),
                  0, 0, 0, 0, None)

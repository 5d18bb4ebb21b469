"""C-ABI boundary: the shared library loads and exports every symbol include/rick_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

from conftest import ROOT
from rick_b200 import _lib


def _declared():
    src = open(os.path.join(ROOT, "include", "rick_b200.h")).read()
    return sorted(set(re.findall(r"RICK_API\s+[\w\s\*]+?\b(rick_\w+)\s*\(", src)))


def test_header_declares_entry_points():
    names = _declared()
    assert "rick_upfirdn2d" in names and "rick_bias_act" in names and "rick_mask_apply" in names
    assert len(names) >= 14


def test_library_exports_every_declared_symbol():
    assert os.path.isfile(_lib.LIB_PATH), "build the library first: python -c 'import __graft_entry__ as g; g.build()'"
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared():
        assert hasattr(handle, name), f"{name} declared in include/rick_b200.h but not exported"


def test_ctypes_table_matches_header():
    assert _lib.exported_symbols() == _declared()


def test_no_compute_entry_points_work_without_gpu():
    lib = _lib.lib()
    assert lib.rick_abi_version() == 1
    assert lib.rick_status_string(0) == b"ok"
    assert lib.rick_status_string(2) == b"unsupported configuration"
    # (in*up + pad0 + pad1 - k) // down + 1, op/upfirdn2d.py:103-104
    assert lib.rick_upfirdn2d_out_size(16, 4, 2, 1, 2, 1) == 32
    assert lib.rick_upfirdn2d_out_size(33, 4, 1, 1, 1, 1) == 32
    assert lib.rick_upfirdn2d_out_size(9, 5, 2, 2, -1, 3) == 8
    assert lib.rick_bias_act_bwd_workspace(2, 128, 65536) > 0


def test_argument_validation_returns_status_not_crash():
    lib = _lib.lib()
    assert lib.rick_upfirdn2d(None, None, None, 1, 4, 4, 1, 4, 4, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, None) == 1
    assert lib.rick_bias_act(None, None, None, None, 4, 1, 1, 3, 0, 0.2, 1.0, 0, None) == 1
    assert lib.rick_percentile(None, None, 0, None, 0, None) == 1


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under rick_b200/ may import it."""
    pkg = os.path.join(ROOT, "rick_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports oracle"
                assert "upfirdn2d_native" not in src or "lives in" in src

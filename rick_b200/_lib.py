"""ctypes binding of ``librick_b200.so`` -- the C-ABI boundary declared in ``include/rick_b200.h``.

The shared object is built in-tree by ``rick_b200/csrc/Makefile`` (``__graft_entry__.build()``); importing this
module never compiles anything (the reference JIT-compiles at import, op/upfirdn2d.py:10-16).  There is no CPU
fallback anywhere in the package: if the library is missing, ``lib()`` raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int64, c_ulonglong, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RICK_B200_LIB") or os.path.join(_HERE, "librick_b200.so")   # (override: A/B of builds)

RICK_F32, RICK_BF16 = 0, 1
ACT_LINEAR, ACT_LRELU = 1, 3

_PROTOTYPES = {
    # name: (restype, [argtypes])
    "rick_abi_version": (c_int, []),
    "rick_status_string": (c_char_p, [c_int]),
    "rick_last_cuda_error": (c_char_p, []),
    "rick_launch_count": (c_ulonglong, []),
    "rick_upfirdn2d_out_size": (c_int, [c_int] * 6),
    "rick_upfirdn2d": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                               c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "rick_bias_act": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int, c_int, c_float,
                              c_float, c_int, c_void_p]),
    "rick_bias_act_bwd_workspace": (c_int64, [c_int64, c_int64, c_int64]),
    "rick_bias_act_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_float,
                                  c_float, c_int, c_void_p]),
    "rick_bias_act_bwd_nhwc_workspace": (c_int64, [c_int64, c_int64]),
    "rick_bias_act_bwd_nhwc": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_float,
                                       c_float, c_void_p]),
    "rick_colsum_workspace": (c_int64, [c_int, c_int64, c_int]),
    "rick_modulate_nhwc": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int, c_void_p]),
    "rick_modulate_bwd_nhwc": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int,
                                       c_void_p]),
    "rick_styled_epilogue_nhwc": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64,
                                          c_int, c_float, c_float, c_void_p]),
    "rick_styled_epilogue_bwd_nhwc": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                              c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int, c_float, c_float,
                                              c_void_p]),
    "rick_fisher_accum": (c_int, [POINTER(c_void_p), POINTER(c_void_p), POINTER(c_int64), c_int, c_int, c_void_p]),
    "rick_fisher_divide": (c_int, [POINTER(c_void_p), POINTER(c_int64), c_int, c_float, c_void_p]),
    "rick_filter_fim": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p]),
    "rick_percentile": (c_int, [c_void_p, c_void_p, c_int64, POINTER(c_double), c_int, c_void_p]),
    "rick_decide": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int, c_int, c_void_p]),
    "rick_mask_apply": (c_int, [POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p),
                                POINTER(c_int64), POINTER(c_int64), c_int, c_void_p]),
    "rick_scale_multi": (c_int, [POINTER(c_void_p), POINTER(c_void_p), POINTER(c_float), POINTER(c_int64), c_int,
                                 c_void_p]),
    "rick_linear_multi": (c_int, [POINTER(c_void_p)] * 4 + [POINTER(c_int64), POINTER(c_int), POINTER(c_float),
                                  POINTER(c_float), c_int, c_int, c_int, c_int, c_float, c_float, c_int, c_void_p]),
    "rick_linear_multi_wgrad": (c_int, [POINTER(c_void_p)] * 4 + [POINTER(c_int64), POINTER(c_int), POINTER(c_float),
                                        POINTER(c_float), c_int, c_int, c_int, c_void_p]),
    "rick_from_rgb_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int, c_int, c_float, c_int,
                                  c_float, c_float, c_void_p]),
    "rick_from_rgb_bwd_data": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int, c_int, c_float,
                                       c_int, c_float, c_float, c_void_p]),
    "rick_weight_sqsum": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int64, c_int64, c_int64, c_void_p]),
    "rick_demod_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_float, c_float,
                               c_void_p]),
    "rick_demod_bwd": (c_int, [c_void_p] * 7 + [c_int, c_int, c_int, c_float, c_float, c_void_p]),
    "rick_add_scale": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_int64, c_void_p]),
    "rick_adam_mask_ema": (c_int, [POINTER(c_void_p)] * 8 + [POINTER(c_int64), POINTER(c_int64), c_int,
                                   c_float, c_float, c_float, c_float, c_float, c_void_p]),
}

# entry points added by later source files (conv_*.cu); bound when present in the header list below
_OPTIONAL: dict = {}

_lib = None


class RickError(RuntimeError):
    pass


def exported_symbols():
    """Names include/rick_b200.h declares (kept in sync by tests/test_abi.py)."""
    return sorted(list(_PROTOTYPES) + list(_OPTIONAL))


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RickError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C rick_b200/csrc`).  rick_b200 has no CPU / PyTorch fallback.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in {**_PROTOTYPES, **_OPTIONAL}.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        if handle.rick_abi_version() != 1:
            raise RickError("librick_b200.so ABI version mismatch; rebuild")
        _lib = handle
    return _lib


def check(status: int, what: str) -> None:
    if status != 0:
        l = lib()
        msg = l.rick_status_string(status).decode()
        if status == 4:
            msg += ": " + l.rick_last_cuda_error().decode()
        raise RickError(f"{what} failed: {msg}")


def ptr_table(ptrs):
    """Host array of device pointers for the multi-tensor entry points."""
    arr = (c_void_p * len(ptrs))()
    for i, p in enumerate(ptrs):
        arr[i] = p if p else None
    return arr


def i64_table(vals):
    arr = (c_int64 * len(vals))()
    for i, v in enumerate(vals):
        arr[i] = int(v)
    return arr

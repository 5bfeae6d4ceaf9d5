"""ctypes binding of libshiftedprox.so -- the C ABI declared in include/shiftedprox.h.

This is the only way the host layer reaches the GPU: there is no CPU fallback and no
other backend.  Importing succeeds without a GPU (the symbols are bound lazily), but any
compute call without the library or without a B200 raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.dirname(_HERE)
# SPX_LIB: developer override used by tools/ to A/B kernel build variants
LIB_PATH = os.environ.get("SPX_LIB") or os.path.join(PKG_DIR, "libshiftedprox.so")

SPX_OK = 0
SPX_E_INVALID = -1
SPX_E_ASSERT_D = -2
SPX_E_BOUNDS = -3
SPX_E_NOROOT = -4
SPX_E_UNSUPPORTED = -5

SEL_ALL, SEL_RANGE, SEL_MASK = 0, 1, 2
H_L1, H_L0, H_LHALF, H_INDBALLL0, H_GROUPL2 = 0, 1, 2, 3, 4
BOX_L1, BOX_L0, BOX_LHALF = 0, 1, 2


class Bound(C.Structure):
    _fields_ = [("vec", C.c_void_p), ("val", C.c_double)]


class Sel(C.Structure):
    _fields_ = [("kind", C.c_int32), ("start", C.c_int64), ("step", C.c_int64), ("stop", C.c_int64),
                ("mask", C.c_void_p), ("list", C.c_void_p), ("nlist", C.c_int64)]


ALLREDUCE_FN = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.POINTER(C.c_double), C.c_int32)


class SpxError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"libshiftedprox status {status}: {msg}")
        self.status = status


_lib = None


def lib() -> C.CDLL:
    """Load libshiftedprox.so (built in-tree by __graft_entry__.build()); fail loudly if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)")
        _lib = C.CDLL(LIB_PATH)
        _lib.spx_last_error.restype = C.c_char_p
    return _lib


def check(status: int) -> None:
    if status == SPX_OK:
        return
    msg = lib().spx_last_error().decode("utf-8", "replace")
    if status == SPX_E_ASSERT_D:
        raise AssertionError(msg)
    raise SpxError(status, msg)


def call(name: str, *args) -> None:
    f = getattr(lib(), name)
    f.restype = C.c_int32
    check(f(*args))


class BoxJobF64(C.Structure):
    """spx_box_job_f64 (include/shiftedprox.h)"""
    _fields_ = [("op", C.c_int32), ("reserved", C.c_int32), ("y_host", C.c_void_p), ("q_or_g_host", C.c_void_p),
                ("d_host", C.c_void_p), ("lambda_", C.c_double), ("sigma", C.c_double)]


BoxJobF32 = BoxJobF64  # same layout: the pointers are untyped here

"""Host-buffer entry points: Box prox!/iprox! on vectors that live in (pinned) HOST memory.

`box_host` is one operation (spx_box_host_*); `box_multi_host` evaluates several operations at the same
shifted point (xk, sj, l, u) in one pass over the host data (spx_box_multi_host_*): every distinct input
vector crosses PCIe once per call.  Arguments are CPU torch tensors (pinned for full speed) or numpy arrays
of one dtype; bounds are tensors/arrays or Python scalars.  The arithmetic is the same CUDA kernels as
the device-resident path -- there is no CPU implementation here.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L

_OPS = {"l1": L.BOX_L1, "l0": L.BOX_L0, "lhalf": L.BOX_LHALF}


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return C.c_void_p(a.ctypes.data)
    assert a.device.type == "cpu" and a.is_contiguous(), "host-path vectors live in host memory"
    return C.c_void_p(a.data_ptr())


def _suffix(a):
    dt = str(a.dtype)
    if dt.endswith("float64"):
        return "f64"
    if dt.endswith("float32"):
        return "f32"
    raise TypeError(f"unsupported element type {dt}")


def _bound(b):
    if isinstance(b, (int, float)):
        return None, float(b)
    return b, 0.0


def box_host(ctx, op, y, xk, sj, q_or_g, l, u, lam, sigma=1.0, d=None, chunk=1 << 22, want_value=False):
    """y <- prox!(ψ, q, σ) (d is None) or iprox!(ψ, g, d) for ψ = shifted(h, xk, l, u) shifted by sj."""
    lv, ls = _bound(l)
    uv, us = _bound(u)
    psi = C.c_double(0.0)
    L.call(f"spx_box_host_{_suffix(xk)}", ctx, C.c_int32(_OPS[op]), C.c_int64(xk.shape[0]), _ptr(y), _ptr(xk),
           _ptr(sj), _ptr(q_or_g), _ptr(d), _ptr(lv), C.c_double(ls), _ptr(uv), C.c_double(us), C.c_double(lam),
           C.c_double(sigma), C.c_int64(chunk), C.byref(psi) if want_value else None)
    return psi.value if want_value else None


def box_multi_host(ctx, jobs, xk, sj, l, u, chunk=1 << 22, want_value=False):
    """jobs: list of dicts {op, y, q (or g), d (None -> prox!), lam, sigma}; one pass over the host data."""
    lv, ls = _bound(l)
    uv, us = _bound(u)
    arr = (L.BoxJobF64 * len(jobs))()
    for j, job in enumerate(jobs):
        arr[j].op = _OPS[job["op"]]
        arr[j].y_host = _ptr(job["y"]).value
        arr[j].q_or_g_host = _ptr(job["q"]).value
        dptr = _ptr(job.get("d"))
        arr[j].d_host = dptr.value if dptr is not None else None
        arr[j].lambda_ = float(job["lam"])
        arr[j].sigma = float(job.get("sigma", 1.0))
    psi = (C.c_double * len(jobs))()
    L.call(f"spx_box_multi_host_{_suffix(xk)}", ctx, C.c_int32(len(jobs)), arr, C.c_int64(xk.shape[0]), _ptr(xk),
           _ptr(sj), _ptr(lv), C.c_double(ls), _ptr(uv), C.c_double(us), C.c_int64(chunk),
           psi if want_value else None)
    return list(psi) if want_value else None

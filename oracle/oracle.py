"""ctypes binding of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY.  May be imported from tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs; never from the product
package.  See the header of shifted_prox_oracle.cpp for what is pinned and what
is "parity unpinned".
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "shifted_prox_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "liboracle.so"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
    return _lib


_DT = {"f64": (np.float64, C.c_double), "f32": (np.float32, C.c_float)}


def _suf(dtype) -> str:
    dtype = np.dtype(dtype)
    if dtype == np.float64:
        return "f64"
    if dtype == np.float32:
        return "f32"
    raise TypeError(f"unsupported dtype {dtype}")


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _vec(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def _bound(b, dtype):
    """Return (vec_or_None, scalar) for a scalar-or-vector bound."""
    if np.ndim(b) == 0:
        return None, float(b)
    return _vec(b, dtype), 0.0


def _sel(selected):
    """selected: None (all) or an iterable of 0-based indices (any order, duplicates kept)."""
    if selected is None:
        return 0, None, 0
    lst = np.ascontiguousarray(np.asarray(list(selected) if not isinstance(selected, np.ndarray) else selected), dtype=np.int64)
    return 1, lst, lst.size


def _call(name, restype, *args):
    f = getattr(lib(), name)
    f.restype = restype
    return f(*args)


i64 = C.c_int64
i32 = C.c_int32
f64 = C.c_double


# ------------------------------------------------------------------ prox ---
def prox_l1(xk, sj, q, lam, sigma):
    s = _suf(q.dtype); xk, sj, q = (_vec(a, q.dtype) for a in (xk, sj, q))
    y = np.empty_like(q)
    _call(f"orc_prox_l1_{s}", None, i64(q.size), _p(y), _p(xk), _p(sj), _p(q), f64(lam), f64(sigma))
    return y


def iprox_l1(xk, sj, g, d, lam):
    s = _suf(g.dtype); xk, sj, g, d = (_vec(a, g.dtype) for a in (xk, sj, g, d))
    y = np.empty_like(g)
    bad = _call(f"orc_iprox_l1_{s}", i64, i64(g.size), _p(y), _p(xk), _p(sj), _p(g), _p(d), f64(lam))
    if bad >= 0:
        raise AssertionError(f"d[{bad}] > 0")
    return y


def prox_l0(xk, sj, q, lam, sigma):
    s = _suf(q.dtype); xk, sj, q = (_vec(a, q.dtype) for a in (xk, sj, q))
    y = np.empty_like(q)
    _call(f"orc_prox_l0_{s}", None, i64(q.size), _p(y), _p(xk), _p(sj), _p(q), f64(lam), f64(sigma))
    return y


def iprox_l0(xk, sj, g, d, lam):
    s = _suf(g.dtype); xk, sj, g, d = (_vec(a, g.dtype) for a in (xk, sj, g, d))
    y = np.empty_like(g)
    bad = _call(f"orc_iprox_l0_{s}", i64, i64(g.size), _p(y), _p(xk), _p(sj), _p(g), _p(d), f64(lam))
    if bad >= 0:
        raise AssertionError(f"d[{bad}] > 0")
    return y


def prox_lhalf(xk, sj, q, lam, sigma):
    s = _suf(q.dtype); xk, sj, q = (_vec(a, q.dtype) for a in (xk, sj, q))
    y = np.empty_like(q); sol = np.empty_like(q)
    _call(f"orc_prox_lhalf_{s}", None, i64(q.size), _p(y), _p(sol), _p(xk), _p(sj), _p(q), f64(lam), f64(sigma))
    return y


def prox_rootlhalf_unshifted(x, lam, gamma):
    s = _suf(x.dtype); x = _vec(x, x.dtype)
    y = np.empty_like(x)
    v = _call(f"orc_prox_rootlhalf_unshifted_{s}", f64, i64(x.size), _p(y), _p(x), f64(lam), f64(gamma))
    return y, v


_BOX_OP = {"l1": 0, "l0": 1, "lhalf": 2}


def prox_box(op, xk, sj, q, l, u, lam, sigma, selected=None):
    dt = q.dtype; s = _suf(dt); xk, sj, q = (_vec(a, dt) for a in (xk, sj, q))
    lv, ls = _bound(l, dt); uv, us = _bound(u, dt)
    kind, lst, nsel = _sel(selected)
    y = np.empty_like(q); sol = np.empty_like(q)
    _call(f"orc_prox_box_{s}", None, i32(_BOX_OP[op]), i64(q.size), _p(y), _p(sol), _p(xk), _p(sj), _p(q),
          _p(lv), f64(ls), _p(uv), f64(us), i32(kind), _p(lst), i64(nsel), f64(lam), f64(sigma))
    return y


def prox_lhalfbox_dbg(xk, sj, q, l, u, lam, sigma, selected=None):
    """ShiftedRootNormLhalfBox prox! plus, per element, the candidate `findmin` chose (0 left edge, 1 right edge,
    2 the kink -xs, 3 the stationary point; -1 not selected), the four objectives (obj[:, :4]) and the stationary
    candidate `val - xs` itself (obj[:, 4]).  Returns (y, pick, obj[n, 5])."""
    dt = q.dtype; s = _suf(dt); xk, sj, q = (_vec(a, dt) for a in (xk, sj, q))
    lv, ls = _bound(l, dt); uv, us = _bound(u, dt)
    kind, lst, nsel = _sel(selected)
    y = np.empty_like(q); sol = np.empty_like(q)
    pick = np.empty(q.size, np.int8); obj = np.full((q.size, 5), np.nan)
    _call(f"orc_prox_lhalfbox_dbg_{s}", None, i64(q.size), _p(y), _p(sol), _p(xk), _p(sj), _p(q),
          _p(lv), f64(ls), _p(uv), f64(us), i32(kind), _p(lst), i64(nsel), f64(lam), f64(sigma), _p(pick), _p(obj))
    return y, pick, obj


def iprox_box(op, xk, sj, g, d, l, u, lam, selected=None):
    dt = g.dtype; s = _suf(dt); xk, sj, g, d = (_vec(a, dt) for a in (xk, sj, g, d))
    lv, ls = _bound(l, dt); uv, us = _bound(u, dt)
    kind, lst, nsel = _sel(selected)
    y = np.zeros_like(g)
    _call(f"orc_iprox_box_{s}", None, i32(_BOX_OP[op]), i64(g.size), _p(y), _p(xk), _p(sj), _p(g), _p(d),
          _p(lv), f64(ls), _p(uv), f64(us), i32(kind), _p(lst), i64(nsel), f64(lam))
    return y


def prox_l1b2(xk, sj, q, lam, sigma, delta, chi_lambda=1.0, return_evals=False):
    dt = q.dtype; s = _suf(dt); xk, sj, q = (_vec(a, dt) for a in (xk, sj, q))
    y = np.empty_like(q)
    ev = _call(f"orc_prox_l1b2_{s}", i32, i64(q.size), _p(y), _p(xk), _p(sj), _p(q), f64(lam), f64(sigma),
               f64(delta), f64(chi_lambda))
    return (y, ev) if return_evals else y


def _offsets(offs):
    return np.ascontiguousarray(offs, dtype=np.int64)


def prox_groupl2(xk, sj, q, offs, lam_g, sigma):
    dt = q.dtype; s = _suf(dt); xk, sj, q, lam_g = (_vec(a, dt) for a in (xk, sj, q, lam_g))
    offs = _offsets(offs); y = np.empty_like(q); sol = np.empty_like(q)
    _call(f"orc_prox_groupl2_{s}", None, i64(q.size), _p(y), _p(sol), _p(xk), _p(sj), _p(q),
          i64(offs.size - 1), _p(offs), _p(lam_g), f64(sigma))
    return y


def prox_groupl2_unshifted(x, offs, lam_g, gamma):
    dt = x.dtype; s = _suf(dt); x, lam_g = _vec(x, dt), _vec(lam_g, dt)
    offs = _offsets(offs); y = np.empty_like(x)
    v = _call(f"orc_prox_groupl2_unshifted_{s}", f64, i64(x.size), _p(y), _p(x), i64(offs.size - 1), _p(offs),
              _p(lam_g), f64(gamma))
    return y, v


def prox_groupl2binf(xk, sj, q, offs, lam_g, sigma, delta):
    dt = q.dtype; s = _suf(dt); xk, sj, q, lam_g = (_vec(a, dt) for a in (xk, sj, q, lam_g))
    offs = _offsets(offs); y = np.empty_like(q); sol = np.empty_like(q)
    _call(f"orc_prox_groupl2binf_{s}", None, i64(q.size), _p(y), _p(sol), _p(xk), _p(sj), _p(q),
          i64(offs.size - 1), _p(offs), _p(lam_g), f64(sigma), f64(delta))
    return y


def prox_groupl2binf_dbg(xk, sj, q, offs, lam_g, sigma, delta, ulp_shift=0):
    """ShiftedGroupNormL2Binf prox! plus per-group diagnostics: the root the bisection ended on (NaN when the
    bracket held no sign change) and whether the group was zeroed.  `ulp_shift` moves every root by that many
    ulps before the final formula (conditioning probe).  Returns (y, nroot[ng], zeroed[ng])."""
    dt = q.dtype; s = _suf(dt); xk, sj, q, lam_g = (_vec(a, dt) for a in (xk, sj, q, lam_g))
    offs = _offsets(offs); y = np.empty_like(q); sol = np.empty_like(q)
    ng = offs.size - 1
    nroot = np.empty(ng, np.float64); zg = np.empty(ng, np.int8)
    _call(f"orc_prox_groupl2binf_dbg_{s}", None, i64(q.size), _p(y), _p(sol), _p(xk), _p(sj), _p(q),
          i64(ng), _p(offs), _p(lam_g), f64(sigma), f64(delta), _p(nroot), _p(zg), i64(int(ulp_shift)))
    return y, nroot, zg.astype(bool)


def prox_indballl0(xk, sj, q, r, delta=None):
    dt = q.dtype; s = _suf(dt); xk, sj, q = (_vec(a, dt) for a in (xk, sj, q))
    y = np.empty_like(q)
    _call(f"orc_prox_indballl0_{s}", None, i64(q.size), _p(y), _p(xk), _p(sj), _p(q), i64(int(r)),
          i32(0 if delta is None else 1), f64(0.0 if delta is None else delta))
    return y


# ---------------------------------------------------------------- values ---
H_KIND = {"l1": 0, "l0": 1, "lhalf": 2, "indballl0": 3, "groupl2": 4}


def value_plain(kind, xk, sj, y, lam=0.0, r=0):
    dt = y.dtype; s = _suf(dt); xk, sj, y = (_vec(a, dt) for a in (xk, sj, y))
    xsy = np.empty_like(y)
    return _call(f"orc_value_plain_{s}", f64, i32(H_KIND[kind]), i64(y.size), _p(xsy), _p(xk), _p(sj), _p(y),
                 f64(lam), i64(int(r)))


def value_box(kind, xk, sj, y, l, u, lam, selected=None):
    dt = y.dtype; s = _suf(dt); xk, sj, y = (_vec(a, dt) for a in (xk, sj, y))
    lv, ls = _bound(l, dt); uv, us = _bound(u, dt)
    k, lst, nsel = _sel(selected)
    return _call(f"orc_value_box_{s}", f64, i32(H_KIND[kind]), i64(y.size), _p(xk), _p(sj), _p(y), _p(lv), f64(ls),
                 _p(uv), f64(us), i32(k), _p(lst), i64(nsel), f64(lam))


def value_l1b2(xk, sj, y, lam, delta):
    dt = y.dtype; s = _suf(dt); xk, sj, y = (_vec(a, dt) for a in (xk, sj, y))
    return _call(f"orc_value_l1b2_{s}", f64, i64(y.size), _p(xk), _p(sj), _p(y), f64(lam), f64(delta))


def value_binf(kind, xk, sj, y, delta, r=0, offs=None, lam_g=None):
    dt = y.dtype; s = _suf(dt); xk, sj, y = (_vec(a, dt) for a in (xk, sj, y))
    ng = 0
    if offs is not None:
        offs = _offsets(offs); lam_g = _vec(lam_g, dt); ng = offs.size - 1
    return _call(f"orc_value_binf_{s}", f64, i32(H_KIND[kind]), i64(y.size), _p(xk), _p(sj), _p(y), f64(delta),
                 i64(int(r)), i64(ng), _p(offs), _p(lam_g))


def value_groupl2(xk, sj, y, offs, lam_g):
    dt = y.dtype; s = _suf(dt); xk, sj, y, lam_g = (_vec(a, dt) for a in (xk, sj, y, lam_g))
    offs = _offsets(offs)
    return _call(f"orc_value_groupl2_{s}", f64, i64(y.size), _p(xk), _p(sj), _p(y), i64(offs.size - 1), _p(offs),
                 _p(lam_g))


def solver_step(op, xk, sj, grad, lam, nu, l=None, u=None, selected=None):
    """The sweeps a solver iteration of the reference's callers wraps around one prox! (R2 / TR of
    RegularizedOptimization.jl, reference README.md:17 -- that package is NOT under /root/reference, so the step
    is DEFINED here as the composition of the reference's own prox! and ψ(y); "parity unpinned" beyond that):

        q = -ν .* ∇f;  s = prox!(ψ, q, ν);  xsy = xk .+ sj .+ s;  ψ(s);  ‖s‖₂;  ∇f's

    op: "l1" | "l0" | "lhalf"; with bounds (l, u) the Box form of the same h.
    Returns (s, xsy, ψ(s), ‖s‖₂, ∇f's) -- the two sums in Float64."""
    dt = grad.dtype
    q = (dt.type(-dt.type(nu)) * grad).astype(dt)
    if l is None:
        s = {"l1": prox_l1, "l0": prox_l0, "lhalf": prox_lhalf}[op](xk, sj, q, lam, nu)
        psi = value_plain(op, xk, sj, s, lam)
    else:
        s = prox_box(op, xk, sj, q, l, u, lam, nu, selected)
        psi = value_box(op, xk, sj, s, l, u, lam, selected)
    xsy = ((xk + sj) + s).astype(dt)
    s64 = s.astype(np.float64)
    return s, xsy, psi, float(np.sqrt(np.sum(s64 * s64))), float(np.sum(grad.astype(np.float64) * s64))


def prox_zero(q, l, u, dtype=np.float64):
    return _call(f"orc_prox_zero_{_suf(dtype)}", f64, f64(q), f64(l), f64(u))


def iprox_zero(d, g, l, u, dtype=np.float64):
    return _call(f"orc_iprox_zero_{_suf(dtype)}", f64, f64(d), f64(g), f64(l), f64(u))


# ------------------------------------------------------ synthetic inputs ---
SEED = 20261018


def uniform(n, stream, dtype=np.float64, scale=1.0, shift=0.0, i0=0, seed=SEED):
    """u(i,k) of SURVEY.md §8d, bit-identical to the device generator spx_fill_uniform_*."""
    dt = np.dtype(dtype); s = _suf(dt)
    out = np.empty(n, dtype=dt)
    ft = C.c_double if s == "f64" else C.c_float
    _call(f"orc_fill_uniform_{s}", None, _p(out), i64(n), i64(i0), C.c_uint64(seed), C.c_uint64(stream),
          ft(scale), ft(shift))
    return out


# ------------------------------------------ spectral thresholding stage ---
def spectral_threshold(kind, U, S, lam, sigma, theta=0.0):
    """The stage between the SVD and the `mul!` of ShiftedRank / ShiftedNuclearnorm / ShiftedCappedl1 prox!, in the
    element type of S (numpy restatement; the SVD itself is LAPACK in the reference and out of scope):
      "rank"     shiftedRank.jl:72-80         c = sqrt(2λσ); S_i <= c -> U[:, i] = 0, else U[:, i] *= S_i (S untouched)
      "nuclear"  shiftedNuclearnorm.jl:72-77  S = max.(0, S .- λσ); U[:, i] *= S_i
      "cappedl1" shiftedCappedl1.jl:71-82     x1 = max(θ, S_i); x2 = min(θ, max(0, S_i - λσ));
                                              S_i = ((x1-S_i)^2/2 + λσθ < (x2-S_i)^2/2 + λσ x2) ? x1 : x2; U[:, i] *= S_i
    Returns (U', S')."""
    dt = S.dtype.type
    U = np.array(U, dtype=dt, copy=True)
    S = np.array(S, dtype=dt, copy=True)
    lam, sigma, theta = dt(lam), dt(sigma), dt(theta)
    if kind == "rank":
        c = np.sqrt(dt(2) * lam * sigma)
        for i in range(S.size):
            if S[i] <= c:
                U[:, i] = 0
            else:
                U[:, i] = U[:, i] * S[i]
    elif kind == "nuclear":
        S = np.maximum(dt(0), S - lam * sigma).astype(dt)
        for i in range(S.size):
            U[:, i] = U[:, i] * S[i]
    elif kind == "cappedl1":
        for i in range(S.size):
            x1 = max(theta, S[i])
            x2 = min(theta, max(dt(0), S[i] - lam * sigma))
            d1, d2 = dt(x1 - S[i]), dt(x2 - S[i])
            if dt(d1 * d1) / dt(2) + lam * sigma * theta < dt(d2 * d2) / dt(2) + lam * sigma * x2:
                S[i] = x1
            else:
                S[i] = x2
            U[:, i] = U[:, i] * S[i]
    else:
        raise ValueError(kind)
    return U, S

"""Helpers shared by the `-m gpu` parity tests: the CUDA path is always reached through the host
mirror `shiftedprox`, i.e. through the C ABI of libshiftedprox.so; the oracle is the checker."""
import numpy as np
import torch

import shiftedprox as sp  # noqa: F401
from oracle import oracle as orc

DEV = "cuda:0"


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def N(t):
    return t.detach().cpu().numpy()


def inputs(n, dtype=np.float64, base=0, seed=orc.SEED):
    """xk = 4u-2, sj = u-0.5, q = 4u-2, streams base..base+2."""
    xk = orc.uniform(n, base + 0, dtype, 4.0, -2.0, seed=seed)
    sj = orc.uniform(n, base + 1, dtype, 1.0, -0.5, seed=seed)
    q = orc.uniform(n, base + 2, dtype, 4.0, -2.0, seed=seed)
    return xk, sj, q


def bounds(n, dtype=np.float64, base=3, seed=orc.SEED):
    """l = -(0.25+u), u = 0.25+u'."""
    l = -(dtype(0.25) + orc.uniform(n, base, dtype, seed=seed))
    u = dtype(0.25) + orc.uniform(n, base + 1, dtype, seed=seed)
    return l.astype(dtype), u.astype(dtype)


def diag(n, dtype=np.float64, base=5, seed=orc.SEED):
    """d: 80 % 0.5+u, 10 % -(0.5+u), 10 % exactly 0 (by hash bucket)."""
    u = orc.uniform(n, base, dtype, seed=seed)
    b = orc.uniform(n, base + 1, np.float64, seed=seed)
    d = dtype(0.5) + u
    d = np.where(b < 0.1, -d, d)
    d = np.where((b >= 0.1) & (b < 0.2), dtype(0), d)
    return d.astype(dtype)


def ulp_diff(a, b):
    """distance in units of the last place between two float arrays of the same dtype (NaN == NaN)."""
    a = np.asarray(a); b = np.asarray(b)
    it = np.int64 if a.dtype == np.float64 else np.int32
    ia = a.view(it).astype(np.int64); ib = b.view(it).astype(np.int64)
    m = np.int64(np.iinfo(it).min)
    ia = np.where(ia < 0, m - ia, ia); ib = np.where(ib < 0, m - ib, ib)
    d = np.abs(ia - ib)
    both_nan = np.isnan(a) & np.isnan(b)
    return np.where(both_nan, 0, d)

"""Helpers shared by the `-m gpu` parity tests: the CUDA path is always reached through the host
mirror `shiftedprox`, i.e. through the C ABI of libshiftedprox.so; the oracle is the checker."""
import json
import os

import numpy as np
import torch

import shiftedprox as sp  # noqa: F401
from oracle import oracle as orc

DEV = "cuda:0"

_STATS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "parity_stats.jsonl")


def log_stat(name, **kw):
    """Append one line of parity statistics (flip counts, ulp-distance percentiles) to gpurun_out/parity_stats.jsonl;
    the builder copies the file of its last full run to profiles/.  Never fails a test."""
    try:
        os.makedirs(os.path.dirname(_STATS), exist_ok=True)
        with open(_STATS, "a") as f:
            f.write(json.dumps({"stat": name, **kw}) + "\n")
    except Exception:
        pass


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


def check_lhalfbox(got, xk, sj, q, l, u, lam, sigma, selected=None, label=""):
    """ShiftedRootNormLhalfBox prox! against the oracle, every element (shiftedRootNormLhalfBox.jl:108, `findmin`:
    first minimal index).  Two assertions:

      1. the CHOSEN CANDIDATE equals the oracle's, except where the objectives of the two picks differ by less
         than 8 ulp(Float64) of their magnitude -- a tie at the precision the reference itself evaluates RNorm in
         (candidate 4 differs from the oracle's by the few ulp its transcendental chain is off, so a true tie can
         legitimately fall either way);
      2. values: candidates 1-3 (the two edges, the kink) bit-exact, candidate 4 (`val - xs`) within 4 ulp(R) of
         the un-shifted value scale, on EVERY element (a flipped pick is compared with the candidate it picked).

    Returns the number of flipped picks (logged by the callers; 0 on every seeded input of the suite)."""
    dt = q.dtype.type
    n = q.size
    ref, pick, obj = orc.prox_lhalfbox_dbg(xk, sj, q, l, u, lam, sigma, selected=selected)
    lv = np.broadcast_to(np.asarray(l, dt), (n,))
    uv = np.broadcast_to(np.asarray(u, dt), (n,))
    xs = xk + sj
    cand = np.stack([lv - sj, uv - sj, -xs], axis=1)  # candidates 1-3 in R, as the reference forms them
    scale = np.abs(xk) + np.abs(sj) + np.abs(q) + 1.0
    tol = 4 * np.finfo(dt).eps * scale
    g64, r64 = got.astype(np.float64), ref.astype(np.float64)
    same_bits = (got.view(np.uint64 if dt == np.float64 else np.uint32)
                 == ref.view(np.uint64 if dt == np.float64 else np.uint32))
    # elements whose value agrees with the oracle's pick
    ok = same_bits | ((pick == 3) & (np.abs(g64 - r64) <= tol)) | (np.isnan(got) & np.isnan(ref))
    idx = np.flatnonzero(~ok)
    flips = 0
    for i in idx:
        assert pick[i] >= 0, (label, "unselected element differs", int(i), got[i], ref[i])
        # which candidate did the CUDA path pick?  an exact edge / kink, else the stationary point
        gp = next((k for k in range(3) if got[i] == cand[i, k] and np.isfinite(obj[i, k])), 3)
        assert gp != pick[i], (label, "same pick, value off", int(i), got[i], ref[i], int(pick[i]))
        a, b = obj[i, gp], obj[i, pick[i]]
        assert np.isfinite(a) and np.isfinite(b), (label, "picked an excluded candidate", int(i), gp, obj[i])
        gap = abs(a - b)
        assert gap <= 8 * np.finfo(np.float64).eps * max(abs(a), abs(b)), \
            (label, "wrong candidate (not a tie)", int(i), int(gp), int(pick[i]), obj[i].tolist())
        if gp == 3:  # the stationary point it picked must be the oracle's candidate 4 to 4 ulp(R) of the scale
            assert abs(g64[i] - obj[i, 4]) <= tol[i], (label, "stationary point off", int(i), got[i], obj[i, 4])
        flips += 1
    if flips:
        print(f"[lhalfbox {label}] {flips} tied picks out of {n} fell on the other candidate (objective gap < 8 ulp)")
    log_stat("lhalfbox_flips", label=label, dtype=np.dtype(dt).name, n=int(n), tied_picks_on_other_candidate=int(flips),
             elements_off_oracle_pick_value=int(idx.size))
    return flips


def check_groupl2binf(got, xk, sj, q, offs, lam_g, sigma, delta, label="", floor=1.0):
    """ShiftedGroupNormL2Binf prox! against the oracle (shiftedGroupNormL2Binf.jl:80-117), every element.

    * support: the groups the reference zeroes (`fl*fm > 0` or n == σλ, :102,107) are exactly the groups the CUDA
      path zeroes (y_g == 0 - (xk+sj)_g bit for bit);
    * values: the oracle ends its bisection on two adjacent floats around the sign change of ITS rounding of froot;
      any other summation order moves that sign change by a few ulps of the root n*.  y depends on n* through
      c(n) = n/(σ(n-σλ)):  δy_i = κ_g (|xk_i| + Δ + |v_i|) δn/n with κ_g = σλ_g/(n*-σλ_g) the conditioning of the
      group (SURVEY.md §8c "ill-conditioned groups": κ -> ∞ as the root approaches the pole of c at σλ).  Bound:
          |y_gpu - y_orc|_i <= eps(R) [ 8 scale_i + 2 κ_g (scale_i + n*_g) ]
      i.e. 8 ulp of the value scale plus 2 ulps of root displacement through the conditioning term; and the
      99.9th percentile of the error over well-conditioned groups (κ_g <= 1) must be <= 4 eps·scale outright.
      Measured (profiles/r02_parity_stats.jsonl): max 1.6 eps·scale on C4, 99.9th percentile <= 1 for κ <= 1, and
      never more than 25 % of this bound in any regime, the ill-conditioned ones (κ up to 65) included."""
    dt = q.dtype.type
    n = q.size
    offs = np.asarray(offs, np.int64)
    ref, nroot, zeroed = orc.prox_groupl2binf_dbg(xk, sj, q, offs, lam_g, sigma, delta)
    sizes = np.diff(offs)
    gid = np.repeat(np.arange(sizes.size), sizes)
    eps = float(np.finfo(dt).eps)
    scale = np.abs(xk).astype(np.float64) + np.abs(sj) + np.abs(q) + floor
    sl = lam_g.astype(np.float64) * float(dt(sigma))
    with np.errstate(divide="ignore", invalid="ignore"):
        kappa = np.where(zeroed | ~np.isfinite(nroot), 0.0, sl / np.abs(nroot - sl))
    kappa = np.nan_to_num(kappa, nan=0.0, posinf=1e300)
    nr = np.nan_to_num(np.where(zeroed, 0.0, nroot), nan=0.0, posinf=0.0)
    err = np.abs(got.astype(np.float64) - ref.astype(np.float64))
    tol = eps * (8.0 * scale + 2.0 * kappa[gid] * (scale + nr[gid]))
    zero_pat = np.zeros(n, dt) - (xk + sj)
    zr, zg = ref == zero_pat, got == zero_pat
    assert np.array_equal(zr, zg), (label, "support differs", np.flatnonzero(zr != zg)[:5])
    bad = np.flatnonzero(~(err <= tol))
    assert bad.size == 0, (label, "value off", bad[:5], err[bad[:5]] / (eps * scale[bad[:5]]), kappa[gid[bad[:5]]])
    rel = err / (eps * scale)
    well = kappa[gid] <= 1.0
    p999 = float(np.percentile(rel[well], 99.9)) if well.any() else 0.0
    assert p999 <= 4.0, (label, "99.9th percentile of the error over well-conditioned groups", p999)
    log_stat("groupl2binf_ulp", label=label, dtype=np.dtype(dt).name, n=int(n), groups=int(sizes.size),
             err_over_eps_scale={"p50": float(np.percentile(rel, 50)), "p99": float(np.percentile(rel, 99)),
                                 "p99.9": float(np.percentile(rel, 99.9)), "max": float(rel.max())},
             well_conditioned_p999=p999, kappa_max=float(kappa.max()), zeroed_groups=int(zeroed.sum()),
             worst_tol_fraction=float((err / tol).max()))
    return rel

#!/usr/bin/env python
"""numpy model of the ShiftedNormL1B2 root search of csrc/spx_l1b2.cu: a ladder pass, then clusters of four trial
values around the root of the cubic Hermite interpolant of the residual on the current bracket.  Counts passes until an
evaluated point has |residual| <= 4 ulp (or the bracket closes)."""
import sys
import numpy as np


def hermite_root(a, fa, da, b, fb, db):
    """root in (a, b) of the cubic with the given values / slopes at both ends (bisection on the cubic)"""
    h = b - a
    def p(x):
        t = (x - a) / h
        h00 = 2*t**3 - 3*t**2 + 1; h10 = t**3 - 2*t**2 + t; h01 = -2*t**3 + 3*t**2; h11 = t**3 - t**2
        return h00*fa + h10*h*da + h01*fb + h11*h*db
    lo, hi = a, b
    for _ in range(200):
        m = 0.5*(lo+hi)
        if not (lo < m < hi): break
        if (p(m) < 0) == (fa < 0): lo = m
        else: hi = m
    return 0.5*(lo+hi)


def run(n, frac, seed=0, dt=np.float64):
    rng = np.random.default_rng(seed)
    xk = (4*rng.random(n)-2).astype(dt); sj = (rng.random(n)-0.5).astype(dt); q = (4*rng.random(n)-2).astype(dt)
    ls = dt(0.1); mid = sj+q; lo = mid-ls; hi = mid+ls
    def ev(eta, Delta):
        s = dt(eta/Delta)
        z = -xk*s; below = z < lo; above = z > hi
        w = np.where(below, lo, np.where(above, hi, z)).astype(np.float64)
        dw = np.where(below | above, 0.0, -xk.astype(np.float64))
        nw = np.sqrt(np.sum(w*w)); dot = np.sum(w*dw)
        return float(eta - nw), float(1 - dot/(nw*Delta))
    n0 = -ev(0.0, 1.0)[0] if False else None
    z = -xk; w = np.clip(z, lo, hi); full = float(np.sqrt(np.sum(w.astype(np.float64)**2)))
    Delta = dt(frac*full)
    ulp = lambda x: float(np.spacing(dt(x)))
    passes = 0
    pts = [float(Delta)*m for m in (1, 2, 4, 16)]
    vals = [ev(p, Delta) for p in pts]; passes += 1
    ev_pts = sorted(zip(pts, vals))
    a = max(p for p, v in ev_pts if v[0] < 0)
    pos = [p for p, v in ev_pts if v[0] > 0]
    if not pos: return None
    b = min(pos)
    D = dict(ev_pts)
    k = 6
    while True:
        fa, da = D[a]; fb, db = D[b]
        best = min(D.items(), key=lambda kv: abs(kv[1][0]))
        if abs(best[1][0]) <= 4*ulp(best[0]) or not (a < dt(a + (b-a)/2) < b):
            return passes, best[0], best[1][0]/best[0], (b-a)/a
        xh = hermite_root(a, fa, da, b, fb, db)
        e = max(4*ulp(xh), (b-a)*2.0**-k)
        cand = [xh - e, xh, xh + e, xh + 3*e if (b - xh) > (xh - a) else xh - 3*e]
        cand = sorted(set(float(dt(c)) for c in cand if a < c < b))
        for c in cand: D[c] = ev(c, Delta)
        passes += 1
        a = max(p for p, v in D.items() if v[0] < 0)
        b = min(p for p, v in D.items() if v[0] > 0) if any(v[0] > 0 for v in D.values()) else b
        k = 16
        if passes > 30: return passes, None, None, None


for n in (1000, 10**5, 4*10**6):
    for frac in (0.9, 0.5, 0.1, 0.01):
        for dt in (np.float64, np.float32):
            print(n, frac, dt.__name__, run(n, frac, dt=dt))

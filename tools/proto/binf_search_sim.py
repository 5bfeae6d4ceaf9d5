#!/usr/bin/env python
"""numpy model of the GroupNormL2Binf root search of csrc/spx_group.cu (binf_solve) against the reference's
bisection (shiftedGroupNormL2Binf.jl:87-107, Roots' fzero restated as bisection to adjacent floats): evaluations
of froot per group and agreement of the roots, over the regimes tests/test_gpu_parity.py exercises on the GPU.

    python tools/proto/binf_search_sim.py [groups-per-regime]
"""
import sys

import numpy as np

eps = np.finfo(float).eps
END_ON_ROOT_RULE = True  # False: the search before the rule (big_lambda: ~60 evaluations per group)


def make(regime, G, m=64, seed=1):
    rng = np.random.default_rng(seed)
    xk = 4 * rng.random((G, m)) - 2; sj = rng.random((G, m)) - 0.5; q = 4 * rng.random((G, m)) - 2
    lam = 0.5 + rng.random(G); sigma = 0.3; delta = 0.5
    if regime == "big_lambda": lam = lam * 200
    if regime == "tiny_lambda": lam = lam * 1e-6
    if regime == "tiny_delta": delta = 1e-4
    if regime == "huge_delta": delta = 50.0
    if regime == "tiny_shift": xk *= 1e-3; sj *= 1e-3; q *= 1e-3; lam = lam * 5
    return xk, sj, q, lam, sigma, delta


def run(regime, G):
    xk, sj, q, lam, sigma, delta = make(regime, G)
    sol = (q + xk) + sj

    def froot(g, n):  # value and derivative, as the kernel's pass returns them
        sl = lam[g] * sigma; gap = n - sl; sc = n / gap
        t = sol[g] - sc * xk[g]; a = np.abs(t) - delta * sc; act = a > 0
        w = np.where(act, np.copysign(a, t) - sol[g], sol[g]); dw = np.where(act, -xk[g] - np.copysign(delta, t), 0.0)
        nw = np.sqrt(np.sum(w * w))
        with np.errstate(all="ignore"):
            df = 1 + (np.sum(w * dw) * sl) / (nw * gap * gap)
        return n - nw, df

    res = {}
    for strat in ("reference", "kernel"):
        out, evs = [], []
        for g in range(G):
            sl = lam[g] * sigma; lmin = sl * (1 + eps); ans = lmin + 1; step = ans / (sigma * (ans - sl))
            u = sol[g] / sigma - step * xk[g]; z = np.sign(u) * np.maximum(0, np.abs(u) - delta * step)
            nsol = np.linalg.norm(sol[g])
            lmax = nsol + sigma * (np.linalg.norm(z) + lam[g] * np.linalg.norm(xk[g]))
            ev = 0
            fa, _ = froot(g, lmin); fb, _ = froot(g, lmax); ev = 2
            if fa * fb > 0:
                out.append(None); evs.append(ev); continue
            if strat == "reference":
                a, bb = lmin, lmax
                while True:
                    mid = a + (bb - a) / 2
                    if not (a < mid < bb): break
                    f, _ = froot(g, mid); ev += 1
                    if f == 0: a = bb = mid; fa = fb = 0; break
                    if (f < 0) == (fa < 0): a, fa = mid, f
                    else: bb, fb = mid, f
                out.append(a if abs(fa) <= abs(fb) else bb); evs.append(ev); continue
            # the kernel: both ends first (as the reference), then safeguarded Newton from lmax with ulp probes
            a, bb = lmin, lmax; x, fx = lmax, fb; _, dx = froot(g, lmax); kulp = 1; tiny = 4096 * eps
            for it in range(400):
                mid = a + (bb - a) / 2
                if not (a < mid < bb): break
                xn = mid; probed = False
                if it < 40:
                    with np.errstate(all="ignore"): xs = x - fx / dx
                    if a < xs < bb: xn = xs
                    elif END_ON_ROOT_RULE and abs(fa) <= tiny * a: xn = a
                    elif END_ON_ROOT_RULE and abs(fb) <= tiny * bb: xn = bb
                    ak = a + kulp * np.spacing(a); bk = bb - kulp * np.spacing(bb)
                    if xn <= ak: xn = ak if ak < mid else mid; probed = True
                    elif xn >= bk: xn = bk if bk > mid else mid; probed = True
                f, d = froot(g, xn); ev += 1; x, fx, dx = xn, f, d
                if f == 0: a = bb = xn; fa = fb = 0; break
                if (f < 0) == (fa < 0): a, fa = xn, f
                else: bb, fb = xn, f
                if probed: kulp = min(kulp * 4, 1 << 20)
            out.append(a if abs(fa) <= abs(fb) else bb); evs.append(ev)
        res[strat] = (out, np.array(evs))
    r, k = res["reference"][0], res["kernel"][0]
    zmis = sum((a is None) != (b is None) for a, b in zip(r, k))
    d = [abs(a - b) / abs(a) for a, b in zip(r, k) if a is not None and b is not None]
    ke = res["kernel"][1]
    lock = ke[: len(ke) // 4 * 4].reshape(-1, 4).max(1).mean()
    print(f"{regime:12s} y = 0 groups {sum(a is None for a in r):4d} (mismatches {zmis})  max rel root difference "
          f"{max(d) if d else 0:.2e}  evaluations per group: bisection {res['reference'][1].mean():.1f}, kernel "
          f"{ke.mean():.2f} (max over 4 in lockstep {lock:.2f})")


if __name__ == "__main__":
    G = int(sys.argv[1]) if len(sys.argv) > 1 else 400
    for reg in ("base", "big_lambda", "tiny_lambda", "tiny_delta", "huge_delta", "tiny_shift"):
        run(reg, G)

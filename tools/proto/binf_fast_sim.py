#!/usr/bin/env python
"""numpy model of the uniform-group fast path of ShiftedGroupNormL2Binf (csrc/spx_group.cu, binf_fast_*):
Float32 Newton search -> one Float64 evaluation + Halley step -> final pass that doubles as the residual check.
Counts evaluations per group, how many groups the fast path accepts, and the distance (in ulps) between the
accepted root and the root of bisection to adjacent floats (the oracle's end state).

    python tools/proto/binf_fast_sim.py [groups-per-regime]
"""
import sys

import numpy as np

eps = np.finfo(np.float64).eps
MODE = "f"
TOLMIN = 4.0
f32 = np.float32


def make(regime, G, m=64, seed=1):
    rng = np.random.default_rng(seed)
    xk = 4 * rng.random((G, m)) - 2; sj = rng.random((G, m)) - 0.5; q = 4 * rng.random((G, m)) - 2
    lam = 0.5 + rng.random(G); sigma = 0.3; delta = 0.5
    if regime == "big_lambda": lam = lam * 200
    if regime == "mid_lambda": lam = lam * 20
    if regime == "tiny_lambda": lam = lam * 1e-6
    if regime == "tiny_delta": delta = 1e-4
    if regime == "huge_delta": delta = 50.0
    if regime == "tiny_shift": xk *= 1e-3; sj *= 1e-3; q *= 1e-3; lam = lam * 5
    return xk, sj, q, lam, sigma, delta


def ev(sol, xk, tau, delta, dt):
    """ss = ||w||^2, dot = sum w dw/dtau at tau (arrays over groups), arithmetic in dt"""
    tau = tau.astype(dt)[:, None]
    t = sol - tau * xk
    sdc = dt(delta) * tau
    a = np.abs(t) - sdc
    act = a > 0
    w = np.where(act, np.copysign(a, t) - sol, sol)
    dw = np.where(act, -xk - np.copysign(dt(delta), t), dt(0))
    return np.sum(w * w, axis=1, dtype=dt), np.sum(w * dw, axis=1, dtype=dt)


def bisect_ref(sol, xk, lam, sigma, delta):
    """bisection to adjacent floats per group (float64), froot in the division-free form"""
    G = sol.shape[0]
    sl = lam * sigma
    lmin = sl * (1 + eps)
    ans = lmin + 1; step = ans / (sigma * (ans - sl))
    u = sol / sigma - step[:, None] * xk
    z = np.sign(u) * np.maximum(0, np.abs(u) - delta * step[:, None])
    lmax = np.linalg.norm(sol, axis=1) + sigma * (np.linalg.norm(z, axis=1) + lam * np.linalg.norm(xk, axis=1))

    def f(n):
        ss, _ = ev(sol, xk, n / (n - sl), delta, np.float64)
        return n - np.sqrt(ss)
    a, b = lmin.copy(), lmax.copy()
    fa, fb = f(a), f(b)
    zero = fa * fb > 0
    for _ in range(200):
        mid = a + (b - a) / 2
        live = (a < mid) & (mid < b)
        if not live.any():
            break
        fm = f(mid)
        left = ((fm < 0) == (fa < 0)) & live
        right = ~left & live
        a = np.where(left, mid, a); fa = np.where(left, fm, fa)
        b = np.where(right, mid, b); fb = np.where(right, fm, fb)
    root = np.where(np.abs(fa) <= np.abs(fb), a, b)
    return root, zero, lmin, lmax


def fast(sol, xk, lam, sigma, delta, lmin, lmax, tol_lo=1e-5, maxit=12):
    G = sol.shape[0]
    sl = lam * sigma
    solf, xkf = sol.astype(f32), xk.astype(f32)
    x = lmax.copy()
    a, b = lmin.copy(), lmax.copy()
    done = np.zeros(G, bool)
    its = np.zeros(G, int)
    for it in range(maxit):
        tau = x / (x - sl)
        ss, dot = ev(solf, xkf, tau, delta, f32)
        nw = np.sqrt(ss.astype(np.float64))
        fx = x - nw
        dfx = 1 + dot.astype(np.float64) * sl / (nw * (x - sl) ** 2)
        its += ~done
        a = np.where(~done & (fx < 0), x, a); b = np.where(~done & (fx >= 0), x, b)
        if MODE == "f":
            step = fx / dfx
        else:  # Newton on h(n) = (n - sl) f(n) / n: linear in n when every entry is thresholded (B = 0)
            gap = x - sl
            h = gap * fx / x
            dh = fx / x + gap * (dfx * x - fx) / (x * x)
            use_h = (MODE == "h") | ((tau * dot.astype(np.float64)) > 0.5 * ss.astype(np.float64))
            step = np.where(use_h, h / dh, fx / dfx)
        xn = x - step
        conv = (np.abs(step) <= tol_lo * x) | (fx == 0)
        inside = (a <= xn) & (xn <= b)
        xn = np.where(inside, xn, a + (b - a) / 2)
        x = np.where(done, x, xn)
        done |= conv
        if done.all():
            break
    lo_fail = ~done
    # one Float64 evaluation + Halley
    tau = x / (x - sl)
    ss, dot = ev(sol, xk, tau, delta, np.float64)
    phi = np.sqrt(ss)
    f = x - phi
    A = dot / tau; B = ss - tau * dot
    gap = x - sl
    tp = -sl / gap ** 2; tpp = 2 * sl / gap ** 3
    php = dot / phi            # tau A / phi
    phpp = A * B / phi ** 3
    fp = 1 - php * tp
    fpp = -(phpp * tp * tp + php * tpp)
    n1 = x - 2 * f * fp / (2 * fp * fp - f * fpp)
    n1_newton = x - f / fp
    # final pass residual (the reference-formula pass of the kernel): f(n1)
    ss1, _ = ev(sol, xk, n1 / (n1 - sl), delta, np.float64)
    res = n1 - np.sqrt(ss1)
    # chord correction with the stale slope where the residual is too large
    kappa = sl / (n1 - sl)
    tolf = np.maximum(TOLMIN, np.minimum(32.0, TOLMIN / kappa)) * eps * n1
    need = np.abs(res) > tolf
    n2 = np.where(need, n1 - res / fp, n1)
    ss2, _ = ev(sol, xk, n2 / (n2 - sl), delta, np.float64)
    res2 = n2 - np.sqrt(ss2)
    still = np.abs(res2) > tolf
    return dict(n=n2, its=its, lo_fail=lo_fail, need_chord=need, fallback=still | lo_fail, n_halley=n1, n_newton=n1_newton,
                kappa=kappa)


def ulps(a, b):
    return np.abs(a.view(np.int64) - b.view(np.int64))


def main():
    global MODE
    G = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
    MODE = sys.argv[2] if len(sys.argv) > 2 else 'adaptive'
    for regime in ("base", "mid_lambda", "big_lambda", "tiny_lambda", "tiny_delta", "huge_delta", "tiny_shift"):
        xk, sj, q, lam, sigma, delta = make(regime, G)
        sol = (q + xk) + sj
        root, zero, lmin, lmax = bisect_ref(sol, xk, lam, sigma, delta)
        with np.errstate(all="ignore"):
            r = fast(sol, xk, lam, sigma, delta, lmin, lmax)
        live = ~zero
        ok = live & ~r["fallback"]
        d = ulps(r["n"][ok], root[ok]) if ok.any() else np.array([0])
        dn = ulps(r["n_newton"][ok], root[ok]) if ok.any() else np.array([0])
        print(f"{regime:12s} zeroed {zero.mean():.3f}  lo its mean {r['its'][live].mean() if live.any() else 0:.2f} max {r['its'][live].max() if live.any() else 0}  "
              f"lo_fail {r['lo_fail'][live].mean() if live.any() else 0:.4f} chord {r['need_chord'][live].mean() if live.any() else 0:.4f} "
              f"fallback {r['fallback'][live].mean() if live.any() else 0:.4f}  kappa med {np.median(r['kappa'][live]) if live.any() else 0:.3g}  "
              f"|root-ref| ulps p50 {np.percentile(d, 50):.0f} p99 {np.percentile(d, 99):.0f} max {d.max()}  (newton-only max {dn.max()})")


if __name__ == "__main__":
    main()

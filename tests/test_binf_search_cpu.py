"""CPU check of the Float32 Newton search of ShiftedGroupNormL2Binf's fast kernels (csrc/spx_group.cu: binf_fast_search,
binf_big_group): Newton on h(n) = (n - σλ) froot(n)/n from lmax in Float32, stopped at a step below kBinfSearchTol·n,
then ONE evaluation in Float64 with the Halley step.  Restated in numpy; the claim pinned here is the one the stopping
rule rests on: the root after the Halley step is within a few Float64 ulps of the root of froot
(shiftedGroupNormL2Binf.jl:80-107 finds it by bisection), in three evaluations or fewer on the bench's data, for short
and long groups and for the sparse-solution regime (λ × 200).  The kernels do not depend on this for correctness --
their final pass is an acceptance test -- only for how many groups fall through to the bracketing search."""
import numpy as np

f32 = np.float32
TOL = f32(2e-3)  # kBinfSearchTol


def search(rng, m, lam_scale):
    xk = rng.uniform(-2, 2, m); sj = rng.uniform(-.5, .5, m); q = rng.uniform(-2, 2, m)
    lam = (0.5 + rng.uniform()) * lam_scale
    sigma, delta = 0.3, 0.5
    sol = (q + xk) + sj
    sl = lam * sigma
    eps = np.finfo(np.float64).eps
    lmin = sl * (1 + eps)

    def froot(n):  # Float64, the reference's formula
        tau = n / (n - sl)
        t = sol - tau * xk
        a = np.abs(t) - tau * delta
        w = np.where(a > 0, np.copysign(a, t) - sol, sol)
        return n - np.sqrt((w * w).sum())

    so, xg = sol.astype(f32), xk.astype(f32)
    slf, delf = f32(sl), f32(delta)
    ans = lmin + 1
    tau_a = f32(sigma) * f32(ans / (sigma * (ans - sl)))
    t = so - tau_a * xg
    a = np.maximum(np.abs(t) - tau_a * delf, 0)
    lmax = np.sqrt((so * so).sum(dtype=f32)) + np.sqrt((a * a).sum(dtype=f32)) + slf * np.sqrt((xg * xg).sum(dtype=f32))
    if froot(float(lmax) * 1.01) < 0 or froot(lmin * (1 + 1e-9)) > 0:
        return None  # no sign change: the group is zeroed, nothing to search
    lo, hi = lmin, float(lmax) * 1.01
    for _ in range(200):
        mid = 0.5 * (lo + hi)
        lo, hi = (mid, hi) if froot(mid) < 0 else (lo, mid)
    root = 0.5 * (lo + hi)

    def ev32(tau):
        t = so - tau * xg
        a = np.abs(t) - tau * delf
        act = a > 0
        z = np.copysign(a, t) - so
        return f32((z[act] ** 2).sum(dtype=f32)), f32((so[~act] ** 2).sum(dtype=f32))

    x, a_, b_, nev = f32(lmax), f32(lmin), f32(lmax), 0
    for it in range(8):
        gap = x - slf
        ssA, ssB = ev32(x / gap)
        nev += 1
        nw = np.sqrt(ssA + ssB)
        fx = x - nw
        dfx = f32(1) + ssA * slf / (x * nw * gap)
        stp = gap * fx / (gap * (dfx - fx / x) + fx)
        xn = x - stp
        a_, b_ = (x, b_) if fx < 0 else (a_, x)
        if not (a_ <= xn <= b_):
            xn = f32(0.5) * (a_ + b_)
        conv = abs(stp) <= TOL * x or fx == 0
        x = f32(xn)
        if conv:
            break
    xd = float(x)
    gapd = xd - sl
    taud = xd / gapd
    t = sol - taud * xk
    a = np.abs(t) - delta * taud
    act = a > 0
    z = np.copysign(a, t) - sol
    ssA, ssB = (z[act] ** 2).sum(), (sol[~act] ** 2).sum()
    phi = np.sqrt(ssA + ssB)
    f = xd - phi
    rD = 1 / (xd * phi * gapd)
    p1 = ssA * sl * rD
    fp = 1 + p1
    fpp = -(p1 * (ssB * sl * rD) * (rD * xd * gapd) + 2 * p1 / gapd)
    n1 = xd - 2 * f * fp / (2 * fp * fp - f * fpp)
    return nev, abs(n1 - root) / root / eps


def test_float32_search_with_the_early_stop_lands_within_ulps_of_the_root():
    rng = np.random.default_rng(1)
    for lam_scale, max_ev in ((1.0, 3), (200.0, 4)):
        for sizes in (np.full(300, 64), np.exp(rng.uniform(np.log(300), np.log(4096), 120)).astype(int)):
            res = [r for r in (search(rng, int(m), lam_scale) for m in sizes) if r is not None]
            assert len(res) > 50
            nev = np.array([r[0] for r in res]); ulps = np.array([r[1] for r in res])
            assert nev.max() <= max_ev, (lam_scale, nev.max())
            assert ulps.max() <= 8.0, (lam_scale, ulps.max())

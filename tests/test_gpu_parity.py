"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): identical support pattern; bit-exact wherever the arithmetic is
+,-,*,/,sqrt and comparisons (L1, L0, their Box forms, iprox, top-r); a few ulp of the value scale for
the transcendental RootNormLhalf forms (Julia's and CUDA's cos/acos are each < 1-2 ulp but not
correctly rounded -- the tolerance is written at each test); solver-limited for L1B2 / GroupNormL2Binf.
"""
import numpy as np
import pytest
import torch

from gpu_util import DEV, N, T, bounds, check_groupl2binf, check_lhalfbox, diag, inputs, log_stat, orc, sp, ulp_diff
from test_oracle_golden import BOX9, GOLD, IPROX14, NU, Q5, isapprox

pytestmark = pytest.mark.gpu

DT = [np.float64, np.float32]
SIZES = [1, 7, 1000, 262_147]


def eps(dt):
    return np.finfo(dt).eps


# ------------------------------------------------------------------ separable ---
@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("lam", [1.0, 10.0])
def test_prox_l1_l0_bit_exact(dt, n, lam):
    xk, sj, q = inputs(n, dt)
    sigma = 0.1
    for h, f in ((sp.NormL1(lam), orc.prox_l1), (sp.NormL0(lam), orc.prox_l0)):
        psi = sp.shifted(sp.shifted(h, T(xk)), T(sj))
        y = torch.empty(n, dtype=T(q).dtype, device=DEV)
        sp.prox_(y, psi, T(q), sigma)
        assert np.array_equal(N(y), f(xk, sj, q, lam, sigma), equal_nan=True)
        # prox(ψ, q, σ) writes ψ.sol
        assert sp.prox(psi, T(q), sigma) is psi.sol
        assert np.array_equal(N(psi.sol), N(y))


@pytest.mark.parametrize("dt", DT)
def test_unaligned_views_take_the_scalar_path(dt):
    n = 10_001
    xk, sj, q = inputs(n + 1, dt)
    txk, tsj, tq = T(xk)[1:], T(sj)[1:], T(q)[1:]
    y = torch.empty(n + 1, dtype=tq.dtype, device=DEV)[1:]
    psi = sp.shifted(sp.shifted(sp.NormL1(1.0), txk), tsj)
    sp.prox_(y, psi, tq, 0.1)
    assert np.array_equal(N(y), orc.prox_l1(xk[1:], sj[1:], q[1:], 1.0, 0.1))


@pytest.mark.parametrize("dt", DT)
def test_prox_aliasing_y_is_q(dt):
    # test_allocs.jl:108 calls prox!(y, ψ, y, 1.0); single-pass semantics (documented in DESIGN.md)
    n = 4099
    xk, sj, q = inputs(n, dt)
    psi = sp.shifted(sp.shifted(sp.NormL0(1.0), T(xk)), T(sj))
    y = T(q).clone()
    sp.prox_(y, psi, y, 0.1)
    assert np.array_equal(N(y), orc.prox_l0(xk, sj, q, 1.0, 0.1))


@pytest.mark.parametrize("dt", DT)
def test_prox_l1_aliasing_y_is_q_documented_deviation(dt):
    """prox!(y, ψ, y, σ) for ShiftedNormL1 (the call of test/test_allocs.jl:108).  The reference's two-pass body
    (shiftedNormL1.jl:47-51) overwrites y with -xk - sj BEFORE it reads q = y, so under aliasing it returns
    t = -(xk + sj) whatever q held; the single-pass kernel reads q first and returns the soft-threshold the
    operator defines.  Pinned here as the documented deviation (DESIGN.md §3.1 "Aliasing"): the
    kernel equals the oracle's un-aliased result, and differs from the reference's aliased artefact exactly where
    the soft-threshold is not clamped at t."""
    n = 4099
    xk, sj, q = inputs(n, dt)
    lam, sigma = 1.0, 0.1
    psi = sp.shifted(sp.shifted(sp.NormL1(lam), T(xk)), T(sj))
    y = T(q).clone()
    sp.prox_(y, psi, y, sigma)
    unaliased = orc.prox_l1(xk, sj, q, lam, sigma)
    assert np.array_equal(N(y), unaliased)
    t = (-xk) - sj  # what the reference's aliased call returns: min(max(t, t - a), t + a) = t
    assert np.array_equal(orc.prox_l1(xk, sj, t, lam, sigma), t)
    differs = N(y) != t
    assert differs.any() and np.array_equal(differs, unaliased != t)


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("n", SIZES)
def test_iprox_l1_l0_bit_exact_and_assertion(dt, n):
    xk, sj, g = inputs(n, dt)
    d = (dt(0.5) + orc.uniform(n, 5, dt)).astype(dt)
    for h, f in ((sp.NormL1(1.3), orc.iprox_l1), (sp.NormL0(1.3), orc.iprox_l0)):
        psi = sp.shifted(sp.shifted(h, T(xk)), T(sj))
        y = torch.empty(n, dtype=T(g).dtype, device=DEV)
        sp.iprox_(y, psi, T(g), T(d))
        assert np.array_equal(N(y), f(xk, sj, g, d, 1.3))
        # @assert d[i] > 0 (shiftedNormL1.jl:70, partial_prox.jl:61)
        dbad = d.copy()
        k = n // 2
        dbad[k] = 0
        with pytest.raises(AssertionError, match=rf"d\[{k}\]"):
            sp.iprox_(y, psi, T(g), T(dbad))


@pytest.mark.parametrize("dt", DT)
def test_iprox_equals_prox_identity(dt):
    # partial_prox.jl:58-72: unboxed iprox(ψ, q, d·1) == prox(ψ, q, σ = d) on ... exactly, d ∈ {1, 2}
    n = 5
    rng = np.random.default_rng(2)
    x = rng.random(n).astype(dt); q = (rng.random(n) - 0.5).astype(dt)
    z = np.zeros(n, dt)
    for h, fi, fp in ((sp.NormL0(3.14), orc.iprox_l0, orc.prox_l0), (sp.NormL1(3.14), orc.iprox_l1, orc.prox_l1)):
        psi = sp.shifted(h, T(x))
        for dv in (1.0, 2.0):
            a = N(sp.iprox(psi, T(q), T(np.full(n, dv, dt)))).copy()
            b = N(sp.prox(psi, T(q), dv)).copy()
            assert np.array_equal(a, fi(x, z, q, np.full(n, dv, dt), 3.14))
            assert np.array_equal(b, fp(x, z, q, 3.14, dv))


def lhalf_scale(xk, sj, q):
    return np.abs(xk) + np.abs(sj) + np.abs(q) + 1.0


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("lam", [1.0, 10.0])
def test_prox_lhalf(dt, n, lam):
    xk, sj, q = inputs(n, dt)
    sigma = 0.1
    psi = sp.shifted(sp.shifted(sp.RootNormLhalf(lam), T(xk)), T(sj))
    y = torch.empty(n, dtype=T(q).dtype, device=DEV)
    sp.prox_(y, psi, T(q), sigma)
    ref = orc.prox_lhalf(xk, sj, q, lam, sigma)
    got = N(y)
    # support: y == -(xk+sj) exactly where the oracle thresholds to zero
    zero_ref = ref == (np.zeros(n, dt) - (xk + sj))
    zero_got = got == (np.zeros(n, dt) - (xk + sj))
    assert np.array_equal(zero_ref, zero_got)
    # values: the closed form is evaluated in Float64 then rounded to R; tolerance 4 ulp(R) of the
    # magnitude of the un-shifted value (y + xs cancels, so ulps of y itself are meaningless)
    tol = 4 * eps(dt) * lhalf_scale(xk, sj, q)
    assert np.all(np.abs(got.astype(np.float64) - ref.astype(np.float64)) <= tol)


def test_lhalf_closed_form_against_mpmath():
    """Both the oracle and the GPU are within 4 ulp of the exact closed form (x = s = 0)."""
    mp = pytest.importorskip("mpmath")
    mp.mp.prec = 200
    n = 600
    q = orc.uniform(n, 7, np.float64, 8.0, -4.0)
    lam, sigma = 1.0, 0.1
    z = np.zeros(n)
    psi = sp.shifted(sp.RootNormLhalf(lam), T(z))
    got = N(sp.prox(psi, T(q), sigma))
    ref = orc.prox_lhalf(z, z, q, lam, sigma)
    nl = mp.mpf(sigma) * mp.mpf(lam)
    for i in range(n):
        if ref[i] == 0.0:
            assert got[i] == 0.0
            continue
        a = mp.mpf(abs(float(q[i])))
        t = nl / 4 * (a / 3) ** mp.mpf(-1.5)
        exact = mp.mpf(2) / 3 * a * (1 + mp.cos(2 * mp.pi / 3 - 2 * mp.acos(t) / 3)) * (1 if q[i] > 0 else -1)
        u = np.spacing(abs(float(exact)))
        assert abs(mp.mpf(float(got[i])) - exact) <= 4 * u, (i, q[i], got[i], exact)
        assert abs(mp.mpf(float(ref[i])) - exact) <= 4 * u


# ------------------------------------------------------------------------ values ---
@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("n", [1, 1000, 262_147])
def test_values_and_fused_values(dt, n):
    xk, sj, q = inputs(n, dt)
    y0 = orc.uniform(n, 9, dt, 1.0, -0.5)
    rtol = 1e-13 if dt == np.float64 else 2e-6
    for name, h, kind in (("l1", sp.NormL1(1.7), "l1"), ("l0", sp.NormL0(1.7), "l0"),
                          ("lhalf", sp.RootNormLhalf(1.7), "lhalf")):
        psi = sp.shifted(sp.shifted(h, T(xk)), T(sj))
        v = psi(T(y0))
        vref = orc.value_plain(kind, xk, sj, y0, 1.7)
        assert v == pytest.approx(vref, rel=rtol)
        # fused: prox! + ψ(y) in one pass equals the stand-alone value at the same y
        y = torch.empty(n, dtype=T(q).dtype, device=DEV)
        _, vf = sp.prox_(y, psi, T(q), 0.1, want_value=True)
        assert vf == pytest.approx(psi(y), rel=rtol)
        assert vf == pytest.approx(orc.value_plain(kind, xk, sj, N(y), 1.7), rel=rtol)
    # exact identities of runtests.jl:175-194: ψ(0) == h(x)
    psi = sp.shifted(sp.NormL0(1.2), T(np.ones(3, dt)))
    assert psi(T(np.zeros(3, dt))) == float(dt(1.2) * dt(3))
    psi = sp.shifted(sp.IndBallL0(2), T(np.ones(3, dt)))
    assert psi(T(np.zeros(3, dt))) == np.inf
    assert psi(T(np.array([-1, 0, 0], dt))) == 0.0


# --------------------------------------------------------------------------- Box ---
SELS = [None, "range", "list"]


def make_sel(kind, n):
    if kind is None:
        return None, None
    if kind == "range":
        return range(0, n, 2), np.arange(0, n, 2)
    rng = np.random.default_rng(7)
    lst = rng.integers(0, n, size=max(1, n // 2))  # unsorted, duplicates (test_allocs.jl:78)
    return lst.tolist(), lst


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("n", [1, 9, 65_539])
@pytest.mark.parametrize("vecb", [True, False])
@pytest.mark.parametrize("selk", SELS)
def test_box_prox(dt, n, vecb, selk):
    xk, sj, q = inputs(n, dt)
    l, u = bounds(n, dt) if vecb else (dt(-1.0), dt(1.0))
    sel_dev, sel_host = make_sel(selk, n)
    lam, sigma = 1.0, 0.1
    tl, tu = (T(l), T(u)) if vecb else (float(l), float(u))
    for name, h in (("l1", sp.NormL1(lam)), ("l0", sp.NormL0(lam)), ("lhalf", sp.RootNormLhalf(lam))):
        psi = sp.shifted(sp.shifted(h, T(xk), tl, tu, sel_dev), T(sj))
        y = torch.empty(n, dtype=T(q).dtype, device=DEV)
        sp.prox_(y, psi, T(q), sigma)
        got = N(y)
        if name != "lhalf":
            ref = orc.prox_box(name, xk, sj, q, l, u, lam, sigma, selected=sel_host)
            assert np.array_equal(got, ref, equal_nan=True), (name, np.flatnonzero(got != ref)[:5])
        else:
            # every element: same candidate as the oracle's findmin (ties to < 8 ulp of the objective excepted and
            # counted), candidates 1-3 bit-exact, candidate 4 within 4 ulp(R) of the value scale
            check_lhalfbox(got, xk, sj, q, l, u, lam, sigma, selected=sel_host, label=f"{dt.__name__} n={n}")


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("case", ["ties_zero", "ties_symmetric", "near_boundary_t1", "wide_range", "inf_bounds"])
def test_lhalfbox_tie_and_edge_inputs(dt, case):
    """Inputs built to sit ON the decisions of shiftedRootNormLhalfBox.jl:108-114: exact ties between candidates
    (`findmin` keeps the first), |xsq| at the real/complex boundary of `val` (t = 1), operands far outside the
    Float32 range of the kernel's filter, infinite bounds."""
    n = 40_003
    xk, sj, q = inputs(n, dt)
    l, u = bounds(n, dt)
    lam, sigma = 1.0, 0.1
    if case == "ties_zero":  # x = s = q = 0, symmetric bounds: left and right edge tie exactly, kink is best
        xk[:] = 0; sj[:] = 0; q[:] = 0
        l, u = -u, u
    elif case == "ties_symmetric":  # q = 0 and xs = 0 on half of the entries: both edges tie, stationary point absent
        half = np.arange(n) % 2 == 0
        xk[half] = 0; sj[half] = 0; q[half] = 0
        l, u = -u, u
    elif case == "near_boundary_t1":  # |xs + q| within a few ulp of 3 (σλ/4)^(2/3), where t crosses 1
        thr = 3.0 * (sigma * lam / 4.0) ** (2.0 / 3.0)
        k = np.arange(n) - n // 2
        xsq = (thr * (1.0 + k * 4.0 * np.finfo(dt).eps)).astype(dt) * np.where(np.arange(n) % 3 == 0, -1, 1).astype(dt)
        q = (xsq - (xk + sj)).astype(dt)
        l = np.full(n, -10.0, dt); u = np.full(n, 10.0, dt)
    elif case == "wide_range":  # magnitudes from 1e-30 to 1e30 (Float64) / 1e-12 to 1e12 (Float32)
        ex = np.linspace(-30, 30, n) if dt == np.float64 else np.linspace(-12, 12, n)
        m = (10.0 ** ex).astype(dt)
        xk = (xk * m).astype(dt); sj = (sj * m).astype(dt); q = (q * m).astype(dt)
        l = (l * m).astype(dt); u = (u * m).astype(dt)
    elif case == "inf_bounds":
        l = np.where(np.arange(n) % 2 == 0, -np.inf, l).astype(dt)
        u = np.where(np.arange(n) % 3 == 0, np.inf, u).astype(dt)
    psi = sp.shifted(sp.shifted(sp.RootNormLhalf(lam), T(xk), T(l), T(u)), T(sj))
    y = torch.empty(n, dtype=T(q).dtype, device=DEV)
    sp.prox_(y, psi, T(q), sigma)
    check_lhalfbox(N(y), xk, sj, q, l, u, lam, sigma, label=f"{dt.__name__} {case}")


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("n", [1, 9, 65_539])
@pytest.mark.parametrize("vecb", [True, False])
@pytest.mark.parametrize("selk", SELS)
def test_box_iprox_bit_exact(dt, n, vecb, selk):
    xk, sj, g = inputs(n, dt)
    d = diag(n, dt)
    l, u = bounds(n, dt) if vecb else (dt(-1.0), dt(1.0))
    sel_dev, sel_host = make_sel(selk, n)
    tl, tu = (T(l), T(u)) if vecb else (float(l), float(u))
    for name, h in (("l1", sp.NormL1(0.8)), ("l0", sp.NormL0(0.8))):
        psi = sp.shifted(sp.shifted(h, T(xk), tl, tu, sel_dev), T(sj))
        y = torch.empty(n, dtype=T(g).dtype, device=DEV)
        sp.iprox_(y, psi, T(g), T(d))
        ref = orc.iprox_box(name, xk, sj, g, d, l, u, 0.8, selected=sel_host)
        assert np.array_equal(N(y), ref, equal_nan=True), (name, np.flatnonzero(N(y) != ref)[:5])


@pytest.mark.parametrize("op", ["l0", "l1", "lhalf"])
def test_box_binf_golden_vectors(op):
    # runtests.jl:449-494: x = 1, Δ = 0.01, λ = 1, ν = 1/9.1e4
    h = {"l0": sp.NormL0(1.0), "l1": sp.NormL1(1.0), "lhalf": sp.RootNormLhalf(1.0)}[op]
    psi = sp.shifted(h, T(np.ones(5)), 0.01, sp.NormLinf(1.0))
    s = N(sp.prox(psi, T(Q5), NU))
    for a, b in zip(s, GOLD[op]):
        assert isapprox(a, b)
    assert np.max(np.abs(s)) <= 0.01
    # set_radius! -> ψ.l == -Δ2, ψ.u == Δ2 (runtests.jl:502-509)
    sp.set_radius_(psi, 0.02)
    assert psi.l == -0.02 and psi.u == 0.02


@pytest.mark.parametrize("op", ["l0", "l1", "lhalf"])
def test_box_prox_nine_cases(op):
    # testsbox.jl:1-99 (l = 0, u = 3, s = -1, σ = 1), atol 1e-2
    c = BOX9[op]
    H = {"l0": sp.NormL0, "l1": sp.NormL1, "lhalf": sp.RootNormLhalf}[op]
    for i in range(9):
        psi = sp.shifted(H(float(c["lam"][i])), T(np.array([float(c["x"][i])])), T(np.array([0.0])), T(np.array([3.0])))
        psi = sp.shifted(psi, T(np.array([-1.0])))
        y = N(sp.prox(psi, T(np.array([float(c["q"][i])])), 1.0))
        assert abs(y[0] - c["sol"][i]) <= 1e-2, (op, i, y[0])


@pytest.mark.parametrize("op", ["l0", "l1"])
def test_box_iprox_fourteen_cases_exact(op):
    # testsbox.jl:101-304 (l = -2, u = 1, s = -1): exact
    c = IPROX14[op]
    H = {"l0": sp.NormL0, "l1": sp.NormL1}[op]
    for i in range(14):
        psi = sp.shifted(H(float(c["lam"][i])), T(np.array([float(c["x"][i])])), T(np.array([-2.0])), T(np.array([1.0])))
        psi = sp.shifted(psi, T(np.array([-1.0])))
        y = N(sp.iprox(psi, T(np.array([float(c["g"][i])])), T(np.array([float(c["d"][i])]))))
        assert y[0] == c["sol"][i], (op, i, y[0])


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("selk", SELS)
def test_box_values(dt, selk):
    n = 10_007
    xk, sj, _ = inputs(n, dt)
    l, u = bounds(n, dt)
    sel_dev, sel_host = make_sel(selk, n)
    rtol = 1e-13 if dt == np.float64 else 2e-6
    # feasible y: sj + y inside [l, u]
    w = (l + (u - l) * orc.uniform(n, 11, dt)).astype(dt)
    y = (w - sj).astype(dt)
    for name, h in (("l1", sp.NormL1(1.1)), ("l0", sp.NormL0(1.1)), ("lhalf", sp.RootNormLhalf(1.1))):
        psi = sp.shifted(sp.shifted(h, T(xk), T(l), T(u), sel_dev), T(sj))
        ref = orc.value_box(name, xk, sj, y, l, u, 1.1, selected=sel_host)
        assert np.isfinite(ref)
        assert psi(T(y)) == pytest.approx(ref, rel=rtol)
        ybad = y.copy()
        ybad[n // 3] += dt(10.0)
        assert psi(T(ybad)) == np.inf
        assert orc.value_box(name, xk, sj, ybad, l, u, 1.1, selected=sel_host) == np.inf
        # fused value equals the stand-alone one (set semantics of `selected`)
        if selk != "list":
            q = orc.uniform(n, 2, dt, 4.0, -2.0)
            yy = torch.empty(n, dtype=T(q).dtype, device=DEV)
            _, vf = sp.prox_(yy, psi, T(q), 0.1, want_value=True)
            assert vf == pytest.approx(psi(yy), rel=rtol)


def test_box_constructor_rejects_l_gt_u():
    # shiftedNormL0Box.jl:33-35
    x = T(np.ones(5))
    with pytest.raises(ValueError):
        sp.shifted(sp.NormL0(1.0), x, T(np.ones(5)), T(np.zeros(5)))
    with pytest.raises(ValueError):
        sp.shifted(sp.NormL1(1.0), x, 1.0, 0.0)


def test_shift_and_set_bounds_semantics():
    # runtests.jl:183-191: shift! writes through into the aliased array
    x = T(np.ones(4)); s = T(np.zeros(4))
    psi = sp.shifted(sp.NormL1(1.0), x)
    assert psi.xk.data_ptr() == x.data_ptr() and float(psi.sj.abs().sum()) == 0.0
    sp.shift_(psi, T(np.full(4, 2.0)))
    assert np.array_equal(N(x), np.full(4, 2.0))
    phi = sp.shifted(psi, s)
    assert phi.shifted_twice and phi.xk.data_ptr() == x.data_ptr()
    sp.shift_(phi, T(np.full(4, 0.5)))
    assert np.array_equal(N(s), np.full(4, 0.5)) and np.array_equal(N(x), np.full(4, 2.0))
    l = T(np.zeros(4)); u = T(np.ones(4))
    box = sp.shifted(sp.NormL0(1.0), x, l, u)
    sp.set_bounds_(box, T(np.full(4, -1.0)), 3.0)
    assert np.array_equal(N(l), np.full(4, -1.0)) and np.array_equal(N(u), np.full(4, 3.0))


# ---------------------------------------------------------------------------- L1B2 ---
def test_l1b2_golden_vector():
    gold = [-0.006367076930786, 0.001288947922799, -0.001130889587543, -0.004285677352167, 0.006176811716709]
    psi = sp.shifted(sp.NormL1(1.0), T(np.ones(5)), 0.01, sp.NormL2(1.0))
    s = N(sp.prox(psi, T(Q5), NU))
    for a, b in zip(s, gold):
        assert isapprox(a, b)
    assert np.linalg.norm(s) <= 0.01 * (1 + 1e-12)
    y = 0.5 * 0.01 * np.ones(5) / np.sqrt(5)
    assert psi(T(y)) == pytest.approx(np.sum(np.abs(1 + y)), rel=1e-14)
    assert psi(T(3 * y)) == np.inf


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("n", [5, 1000, 262_147])
def test_l1b2_against_oracle(dt, n):
    xk, sj, q = inputs(n, dt)
    lam, sigma = 1.0, 0.1
    y0 = orc.prox_l1b2(xk, sj, q, lam, sigma, 1e30)  # ball inactive
    full = float(np.linalg.norm((y0 + sj).astype(np.float64)))
    for delta in (2.0 * full, 0.5 * full):
        psi = sp.shifted(sp.shifted(sp.NormL1(lam), T(xk), delta, sp.NormL2(1.0)), T(sj))
        y = torch.empty(n, dtype=T(q).dtype, device=DEV)
        _, val = sp.prox_(y, psi, T(q), sigma, want_value=True)
        ref = orc.prox_l1b2(xk, sj, q, lam, sigma, delta)
        got = N(y)
        # solver-limited: η agrees to a few ulp, y is linear in 1/η between kinks
        tol = (64 if dt == np.float64 else 16) * eps(dt) * (np.abs(ref) + np.abs(sj) + 1)
        assert np.all(np.abs(got.astype(np.float64) - ref.astype(np.float64)) <= tol)
        assert np.isfinite(val)
        assert val == pytest.approx(orc.value_l1b2(xk, sj, got, lam, delta), rel=1e-12 if dt == np.float64 else 1e-5)
        if delta < full:
            assert psi.last_passes <= (14 if dt == np.float64 else 9)
            nrm = np.linalg.norm((got + sj).astype(np.float64))
            assert nrm == pytest.approx(delta, rel=1e-10 if dt == np.float64 else 1e-5)


# -------------------------------------------------------------------------- groups ---
def ragged_offsets(ngroups, maxlen, seed=3):
    rng = np.random.default_rng(seed)
    sizes = np.floor(np.exp(rng.uniform(0, np.log(maxlen + 1), ngroups))).astype(np.int64).clip(1, maxlen)
    return np.concatenate([[0], np.cumsum(sizes)])


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("layout", ["one", "one_big", "two", "g64", "ragged"])
def test_group_l2_prox_and_value(dt, layout):
    # "one_big": a single group spanning a long vector (ShiftedGroupNormL2 from NormL2) -> grid-wide path
    offs = {"one": np.array([0, 777]), "one_big": np.array([0, 100_003]), "two": np.array([0, 3, 6]),
            "g64": np.arange(0, 64 * 501, 64), "ragged": ragged_offsets(300, 4096)}[layout]
    n = int(offs[-1]); ng = len(offs) - 1
    xk, sj, q = inputs(n, dt)
    lam_g = (dt(0.5) + orc.uniform(ng, 12, dt)).astype(dt)
    sigma = 0.3
    h = sp.GroupNormL2(T(lam_g), None, offsets=T(offs))
    psi = sp.shifted(sp.shifted(h, T(xk)), T(sj))
    y = torch.empty(n, dtype=T(q).dtype, device=DEV)
    _, val = sp.prox_(y, psi, T(q), sigma, want_value=True)
    ref = orc.prox_groupl2(xk, sj, q, offs, lam_g, sigma)
    got = N(y)
    # α = max(1 - σλ/‖sol‖, 0); the norm is within 1 ulp of exact on both sides (order unspecified
    # in the reference: BLAS nrm2).  Tolerance: 8 ulp of the un-shifted magnitude.
    tol = 8 * eps(dt) * (np.abs(xk) + np.abs(sj) + np.abs(q) + 1)
    assert np.all(np.abs(got.astype(np.float64) - ref.astype(np.float64)) <= tol)
    # support: whole groups thresholded to zero agree
    zr = ref == (np.zeros(n, dt) - (xk + sj)); zg = got == (np.zeros(n, dt) - (xk + sj))
    assert np.array_equal(zr, zg)
    rtol = 1e-13 if dt == np.float64 else 2e-6
    assert val == pytest.approx(orc.value_groupl2(xk, sj, got, offs, lam_g), rel=rtol)
    assert psi(y) == pytest.approx(val, rel=rtol)


@pytest.mark.parametrize("dt", DT)
def test_group_launch_census_is_only_a_hint(dt):
    """spx_group_validate_offsets records which group-size classes a layout holds and the entry points skip the
    CTA-per-group launches of an absent class.  (1) A layout of short groups only (no class launched) and one of short
    + 257..1024 groups only (one class) agree with the oracle.  (2) The census is a hint: when the offsets are
    rewritten in place after validation (same address, ngroups and n, now with groups of every class) prox! and ψ(y)
    of both group types still agree with the oracle."""
    rng = np.random.default_rng(11)
    G = 400
    short = rng.integers(1, 200, G)
    n = int(short.sum())
    mixed = np.full(G, 1, np.int64)
    mixed[rng.choice(G, 10, replace=False)] = [2000, 1500, 1100, 3000, 1025, 500, 700, 300, 1024, 257]
    mixed[mixed == 1] += np.diff(np.concatenate([[0], np.sort(rng.choice(n - mixed.sum() + 1, G - 11)), [n - mixed.sum()]]))
    assert int(mixed.sum()) == n and mixed.max() <= 4096
    midonly = short.copy()
    midonly[:3] = [600, 900, 300]
    layouts = {"short": np.concatenate([[0], np.cumsum(short)]), "mid": np.concatenate([[0], np.cumsum(midonly)])}
    lam_g = (dt(0.5) + orc.uniform(G, 12, dt)).astype(dt)
    sigma, delta = 0.3, 0.5
    rtol = 1e-13 if dt == np.float64 else 2e-6

    def check(psi2, psib, offs, xk, sj, q):
        m = int(offs[-1])
        tol = 8 * eps(dt) * (np.abs(xk) + np.abs(sj) + np.abs(q) + 1)
        y = torch.empty(m, dtype=T(q).dtype, device=DEV)
        for want in (False, True):
            r = sp.prox_(y, psi2, T(q), sigma, want_value=want)
            got = N(y)
            ref = orc.prox_groupl2(xk, sj, q, offs, lam_g, sigma)
            assert np.all(np.abs(got.astype(np.float64) - ref.astype(np.float64)) <= tol)
            if want:
                assert r[1] == pytest.approx(orc.value_groupl2(xk, sj, got, offs, lam_g), rel=rtol)
        assert psi2(y) == pytest.approx(orc.value_groupl2(xk, sj, N(y), offs, lam_g), rel=rtol)
        sp.prox_(y, psib, T(q), sigma)
        refb = orc.prox_groupl2binf(xk, sj, q, offs, lam_g, sigma, delta)
        scale = np.abs(xk) + np.abs(sj) + np.abs(q) + 1
        assert np.all(np.abs(N(y).astype(np.float64) - refb.astype(np.float64)) <= (1e-9 if dt == np.float64 else 2e-4) * scale)
        assert psib(y) == pytest.approx(orc.value_binf("groupl2", xk, sj, N(y), delta, offs=offs, lam_g=lam_g), rel=rtol)

    for name, offs in layouts.items():
        m = int(offs[-1])
        xk, sj, q = inputs(m, dt)
        t_offs = T(offs)
        h = sp.GroupNormL2(T(lam_g), None, offsets=t_offs)
        psi2 = sp.shifted(sp.shifted(h, T(xk)), T(sj))
        psib = sp.shifted(sp.shifted(h, T(xk), delta, sp.NormLinf(1.0)), T(sj))
        check(psi2, psib, offs, xk, sj, q)
        if name == "short":  # rewrite the validated layout in place: the recorded census (no class) is now wrong
            new = np.concatenate([[0], np.cumsum(mixed)])
            for psi in (psi2, psib):
                psi._offs.copy_(T(new))
            check(psi2, psib, new, xk, sj, q)


def test_group_l2_from_norml2_differential():
    # runtests.jl:244-251: ShiftedGroupNormL2 from NormL2 == NormL2 prox of (q + x) minus x
    rng = np.random.default_rng(5)
    x = rng.random(6); q = rng.random(6); lam, nu = 0.7, 0.4
    psi = sp.shifted(sp.NormL2(lam), T(x))
    y = N(sp.prox(psi, T(q), nu))
    v = q + x
    ytrue = max(1 - nu * lam / np.linalg.norm(v), 0) * v - x
    assert np.linalg.norm(y - ytrue) <= 1e-11


def test_group_l2binf_golden_vectors():
    # runtests.jl:587-606 (from NormL2, one group) and :648-705 (two groups)
    gold = [-0.010000000000000, 0.005862191941930, -0.005131948291800, -0.010000000000000, 0.010000000000000]
    psi = sp.shifted(sp.NormL2(1.0), T(np.ones(5)), 0.01, sp.NormLinf(1.0))
    s = N(sp.prox(psi, T(Q5), NU))
    for a, b in zip(s, gold):
        assert isapprox(a, b)
    lam = np.array([0.396767474230670, 0.538816734003357])
    q = np.array([-0.649013765191241, 1.181166041965532, -0.758453297283692, -1.109613038501522,
                  -0.845551240007797, -0.572664866457950])
    psi = sp.shifted(sp.GroupNormL2(lam.tolist(), [range(0, 3), range(3, 6)]), T(np.ones(6)), 0.01, sp.NormLinf(1.0))
    s = N(sp.prox(psi, T(q), 0.419194514403295))
    for a, b in zip(s, [-0.01, 0.01, -0.01, -0.01, -0.01, -0.01]):
        assert isapprox(a, b)
    y = 0.004 * np.ones(6)
    v = 1 + y
    assert psi(T(y)) == pytest.approx(lam[0] * np.linalg.norm(v[:3]) + lam[1] * np.linalg.norm(v[3:]), rel=1e-14)
    assert psi(T(3 * y)) == np.inf


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("layout", ["g64", "ragged", "ragged4096", "edges"])
def test_group_l2binf_against_oracle(dt, layout):
    # "ragged4096" / "edges": groups of 1025..4096 elements take a whole CTA (group_l2binf_big_kernel); the sizes
    # either side of both limits stay on the warp paths
    offs = {"g64": np.arange(0, 64 * 201, 64), "ragged": ragged_offsets(120, 1500),
            "ragged4096": ragged_offsets(90, 4096, seed=11),
            "edges": np.concatenate([[0], np.cumsum([1024, 1025, 7, 4096, 4097, 256, 257, 2049, 1, 3000])])}[layout]
    n = int(offs[-1]); ng = len(offs) - 1
    xk, sj, q = inputs(n, dt)
    lam_g = (dt(0.5) + orc.uniform(ng, 12, dt)).astype(dt)
    sigma, delta = 0.3, 0.5
    h = sp.GroupNormL2(T(lam_g), None, offsets=T(offs))
    psi = sp.shifted(sp.shifted(h, T(xk), delta, sp.NormLinf(1.0)), T(sj))
    y = torch.empty(n, dtype=T(q).dtype, device=DEV)
    sp.prox_(y, psi, T(q), sigma)
    # support identical; values within 8 ulp of the value scale + 2 ulps of root displacement through the group's
    # conditioning κ_g = σλ_g/(n* - σλ_g) (see check_groupl2binf); 99.9th percentile over κ_g <= 1 within 4 ulp
    check_groupl2binf(N(y), xk, sj, q, offs, lam_g, sigma, delta, label=f"{dt.__name__} {layout}")


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("delta_frac", [0.5, 1e9])
def test_l1b2_nan_in_xk_matches_the_oracle(dt, delta_frac):
    """A NaN in xk makes every norm of the search NaN (the Float32 passes clamp with min/max instructions, which would
    drop it: the packet's NaN flag puts it back): whatever the reference does with it -- the oracle restates it -- is
    what comes out, NaN positions included."""
    n = 50_001
    xk, sj, q = inputs(n, dt)
    y0 = orc.prox_l1b2(xk, sj, q, 1.0, 0.1, 1e30)
    full = float(np.linalg.norm((y0 + sj).astype(np.float64)))
    xk = xk.copy()
    xk[12345] = np.nan
    delta = delta_frac * full if delta_frac < 1e6 else 1e30
    psi = sp.shifted(sp.shifted(sp.NormL1(1.0), T(xk), delta, sp.NormL2(1.0)), T(sj))
    y = N(sp.prox(psi, T(q), 0.1))
    ref = orc.prox_l1b2(xk, sj, q, 1.0, 0.1, delta)
    assert np.array_equal(np.isnan(y), np.isnan(ref))
    ok = ~np.isnan(ref)
    assert np.all(np.abs(y[ok].astype(np.float64) - ref[ok].astype(np.float64)) <= 64 * eps(dt) * (np.abs(ref[ok]) + 1))


# --------------------------------------------------------------------------- top-r ---
@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("n,r", [(6, 2), (1000, 1), (1000, 999), (1000, 1000), (1000, 2000), (16_384, 100),
                                 (16_385, 1024), (65_536, 1024), (100_003, 5000), (131_072, 77), (300_001, 1234)])
@pytest.mark.parametrize("binf", [False, True])
def test_indballl0_bit_exact(dt, n, r, binf):
    xk, sj, q = inputs(n, dt)
    h = sp.IndBallL0(r)
    psi = sp.shifted(h, T(xk), 1.0, sp.NormLinf(1.0)) if binf else sp.shifted(h, T(xk))
    psi = sp.shifted(psi, T(sj))
    y = torch.empty(n, dtype=T(q).dtype, device=DEV)
    sp.prox_(y, psi, T(q), 1.0)
    ref = orc.prox_indballl0(xk, sj, q, r, delta=1.0 if binf else None)
    assert np.array_equal(N(y), ref), np.flatnonzero(N(y) != ref)[:8]


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("n,r", [(4096, 300), (70_001, 1024), (200_000, 4321), (3_000_017, 1_234_567)])
def test_indballl0_ties_keep_lowest_index(dt, n, r):
    # magnitudes quantised to 1/64: massive ties at the threshold (SURVEY.md §8d tie-stress)
    # (the last case: the multi-pass global path with thousands of threshold-equal entries spread over every
    # per-warp slice of its final pass)
    xk = np.zeros(n, dt); sj = np.zeros(n, dt)
    q = (np.round(orc.uniform(n, 2, dt, 4.0, -2.0) * 64) / 64).astype(dt)
    psi = sp.shifted(sp.IndBallL0(r), T(xk))
    y = N(sp.prox(psi, T(q), 1.0))
    ref = orc.prox_indballl0(xk, sj, q, r)
    assert np.array_equal(y, ref)
    assert np.count_nonzero(y) <= r


def test_indballl0_nan_and_source_text():
    z = np.array([1.0, np.nan, 3.0])
    psi = sp.shifted(sp.IndBallL0(1), T(np.zeros(3)))
    y = N(sp.prox(psi, T(z), 1.0))
    assert np.isnan(y[1]) and y[0] == 0 and y[2] == 0
    z = np.array([1.0, -2.0, 2.0, 0.5, -2.0, 2.0])
    psi = sp.shifted(sp.IndBallL0(3), T(np.zeros(6)))
    assert np.array_equal(N(sp.prox(psi, T(z), 1.0)), [0, -2.0, 2.0, 0, -2.0, 0])


@pytest.mark.parametrize("dt", DT)
def test_indballl0_batched(dt):
    nprob, n, r = 37, 8192 + 64, 100
    xk, sj, q = inputs(nprob * n, dt)
    psi = sp.shifted(sp.shifted(sp.IndBallL0(r), T(xk), 1.0, sp.NormLinf(1.0), nprob=nprob), T(sj))
    y = torch.empty(nprob * n, dtype=T(q).dtype, device=DEV)
    sp.prox_(y, psi, T(q), 1.0)
    got = N(y)
    for p in range(nprob):
        sl = slice(p * n, (p + 1) * n)
        assert np.array_equal(got[sl], orc.prox_indballl0(xk[sl], sj[sl], q[sl], r, delta=1.0)), p


@pytest.mark.parametrize("dt", DT)
def test_indballl0_batch_mixing_ordinary_and_flagged_problems(dt):
    """The linear-bin select flags what it cannot decide cheaply (crowded threshold bin, Inf/NaN, all zeros)
    and the radix kernel finishes exactly those problems: a batch mixing both kinds must match the oracle."""
    nprob, n, r = 9, 20_000, 700
    xk, sj, q = inputs(nprob * n, dt)
    xk = xk.copy(); sj = sj.copy(); q = q.copy()
    sl = lambda p: slice(p * n, (p + 1) * n)  # noqa: E731
    xk[sl(1)] = 0; sj[sl(1)] = 0; q[sl(1)] = dt(1.5)                      # all equal: 20000 ties in one bin
    xk[sl(3)] = 0; sj[sl(3)] = 0; q[sl(3)] = 0                             # all zeros
    q[3 * n + 17] = 0
    q[4 * n + 5] = np.inf                                                  # an infinity
    q[5 * n + 123] = np.nan                                                # a NaN (sorts largest)
    q[sl(7)] = (np.round(q[sl(7)] * 4) / 4).astype(dt); xk[sl(7)] = 0; sj[sl(7)] = 0   # ~20 distinct magnitudes
    psi = sp.shifted(sp.shifted(sp.IndBallL0(r), T(xk), nprob=nprob), T(sj))
    y = torch.empty(nprob * n, dtype=T(q).dtype, device=DEV)
    sp.prox_(y, psi, T(q), 1.0)
    got = N(y)
    for p in range(nprob):
        ref = orc.prox_indballl0(xk[sl(p)], sj[sl(p)], q[sl(p)], r)
        assert np.array_equal(got[sl(p)], ref, equal_nan=True), p


# ------------------------------------------- unshifted in-tree base functions ---
@pytest.mark.parametrize("dt", DT)
def test_unshifted_rootnormlhalf_and_groupnorml2_prox(dt):
    """prox!(y, h, x, γ) of RootNormLhalf (rootNormLhalf.jl:31-51) and GroupNormL2 (groupNormL2.jl:41-58):
    the shifted kernels with NULL shifts; the returned value is the reference's (h(y), resp. Σ λ_g ‖x_g‖)."""
    n = 40_003
    x = orc.uniform(n, 2, dt, 4.0, -2.0)
    y = torch.empty(n, dtype=T(x).dtype, device=DEV)
    _, v = sp.prox_(y, sp.RootNormLhalf(0.9), T(x), 0.2)
    ref, vref = orc.prox_rootlhalf_unshifted(x, 0.9, 0.2)
    got = N(y)
    assert np.array_equal(got == 0, ref == 0)
    assert np.all(np.abs(got.astype(np.float64) - ref.astype(np.float64)) <= 4 * eps(dt) * (np.abs(x) + 1))
    # the reference sums in R, sequentially (ysum += ...); the kernel sums in Float64
    assert v == pytest.approx(vref, rel=1e-12 if dt == np.float64 else 1e-4)
    offs = ragged_offsets(200, 700)
    n = int(offs[-1])
    x = orc.uniform(n, 2, dt, 4.0, -2.0)
    lam_g = (dt(0.5) + orc.uniform(len(offs) - 1, 12, dt)).astype(dt)
    y = torch.empty(n, dtype=T(x).dtype, device=DEV)
    _, v = sp.prox_(y, sp.GroupNormL2(T(lam_g), None, offsets=T(offs)), T(x), 0.3)
    ref, vref = orc.prox_groupl2_unshifted(x, offs, lam_g, 0.3)
    assert np.all(np.abs(N(y).astype(np.float64) - ref.astype(np.float64)) <= 8 * eps(dt) * (np.abs(x) + 1))
    assert v == pytest.approx(vref, rel=1e-12 if dt == np.float64 else 1e-4)


# ----------------------------------------------------------------- edge cases: aliasing, empty input ---
@pytest.mark.parametrize("dt", DT)
def test_prox_in_place_for_every_operator_family(dt):
    """prox!(y, ψ, y, σ) (test_allocs.jl:108 style): Box, L1B2 (the search must not use y as scratch then), groups
    (short and long), top-r -- the in-place result equals the out-of-place one."""
    n = 20_011
    xk, sj, q = inputs(n, dt)
    l, u = bounds(n, dt)
    offs = ragged_offsets(60, 2000)
    m = int(offs[-1])
    lam_g = (dt(0.5) + orc.uniform(len(offs) - 1, 12, dt)).astype(dt)
    y0 = orc.prox_l1b2(xk, sj, q, 1.0, 0.1, 1e30)
    full = float(np.linalg.norm((y0 + sj).astype(np.float64)))
    hg = sp.GroupNormL2(T(lam_g), None, offsets=T(offs))
    cases = [
        (sp.shifted(sp.shifted(sp.RootNormLhalf(1.0), T(xk), T(l), T(u)), T(sj)), n, 0.1),
        (sp.shifted(sp.shifted(sp.NormL1(1.0), T(xk), 0.5 * full, sp.NormL2(1.0)), T(sj)), n, 0.1),
        (sp.shifted(sp.shifted(hg, T(xk[:m])), T(sj[:m])), m, 0.3),
        (sp.shifted(sp.shifted(hg, T(xk[:m]), 0.5, sp.NormLinf(1.0)), T(sj[:m])), m, 0.3),
        (sp.shifted(sp.shifted(sp.IndBallL0(777), T(xk), 1.0, sp.NormLinf(1.0)), T(sj)), n, 1.0),
    ]
    for psi, k, sigma in cases:
        out = torch.empty(k, dtype=T(q).dtype, device=DEV)
        sp.prox_(out, psi, T(q[:k]), sigma)
        y = T(q[:k]).clone()
        sp.prox_(y, psi, y, sigma)
        assert torch.equal(y, out), type(psi).__name__


def test_empty_vectors_are_accepted_by_every_entry_point():
    """n = 0 (and no groups / no problems): every prox!/iprox!/ψ(y) entry returns SPX_OK without touching memory."""
    import ctypes as C
    from shiftedprox import _lib as L

    ctx = sp.context(DEV)
    z64, zd, nul = C.c_int64(0), C.c_double(1.0), None
    out = C.c_double(-1.0)
    for suf in ("f64", "f32"):
        for name in ("prox_l1", "prox_l0", "prox_lhalf"):
            L.call(f"spx_{name}_{suf}", ctx, z64, nul, nul, nul, nul, zd, zd, C.byref(out))
            assert out.value == 0.0
        for name in ("iprox_l1", "iprox_l0"):
            L.call(f"spx_{name}_{suf}", ctx, z64, nul, nul, nul, nul, nul, zd, nul, nul)
        b = L.Bound(None, 1.0)
        for name in ("prox_l1box", "prox_l0box", "prox_lhalfbox"):
            L.call(f"spx_{name}_{suf}", ctx, z64, nul, nul, nul, nul, C.byref(b), C.byref(b), nul, zd, zd, nul)
        L.call(f"spx_prox_l1b2_{suf}", ctx, z64, nul, nul, nul, nul, zd, zd, zd, zd, nul, nul)
        L.call(f"spx_prox_groupl2_{suf}", ctx, z64, nul, nul, nul, nul, z64, nul, nul, zd, nul)
        L.call(f"spx_prox_groupl2binf_{suf}", ctx, z64, nul, nul, nul, nul, z64, nul, nul, zd, zd, nul)
        L.call(f"spx_prox_indballl0_{suf}", ctx, z64, z64, nul, nul, nul, nul, C.c_int64(3), C.c_int32(0), zd)


# ---------------------------------------------------------------- strided SubArray shifts ---
@pytest.mark.parametrize("dt", DT)
def test_strided_view_shift_aliases_the_parent_array(dt):
    # runtests.jl:199-215: `y = rand(Float32, 10); x = view(y, 1:2:10); ψ = shifted(h, x); ψ(zeros(5)) == h(x)`;
    # and the aliasing of shift! (runtests.jl:183-191) through the view: ψ.xk .= v lands in the parent array
    n = 4099
    parent = orc.uniform(2 * n, 0, dt, 4.0, -2.0)
    tp = T(parent)
    x = tp[0::2]
    assert not x.is_contiguous()
    lam = 1.2
    for h, kind in ((sp.NormL1(lam), "l1"), (sp.NormL0(lam), "l0"), (sp.RootNormLhalf(lam), "lhalf")):
        psi = sp.shifted(h, x)
        z = np.zeros(n, dt)
        assert psi(T(z)) == pytest.approx(orc.value_plain(kind, parent[0::2], z, z, lam), rel=1e-6 if dt == np.float32 else 1e-13)
    psi = sp.shifted(sp.NormL1(lam), x)
    _, sj, q = inputs(n, dt)
    psi2 = sp.shifted(psi, T(sj))
    y = torch.empty(n, dtype=tp.dtype, device=DEV)
    sp.prox_(y, psi2, T(q), 0.1)
    assert np.array_equal(N(y), orc.prox_l1(parent[0::2].copy(), sj, q, lam, 0.1))
    # the caller writes the parent array: ψ sees it (the shadow is re-gathered before every call)
    tp[0::2] += 1.0
    sp.prox_(y, psi2, T(q), 0.1)
    assert np.array_equal(N(y), orc.prox_l1((parent[0::2] + dt(1.0)).astype(dt), sj, q, lam, 0.1))
    # shift!(ψ, v) writes through the view into the parent array, odd entries untouched
    v = orc.uniform(n, 9, dt, 2.0, -1.0)
    before_odd = N(tp[1::2]).copy()
    sp.shift_(psi, T(v))
    assert np.array_equal(N(tp[0::2]), v) and np.array_equal(N(tp[1::2]), before_odd)
    sp.prox_(y, psi2, T(q), 0.1)
    assert np.array_equal(N(y), orc.prox_l1(v, sj, q, lam, 0.1))


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("regime", ["big_lambda", "tiny_lambda", "tiny_delta", "huge_delta", "tiny_shift"])
def test_group_l2binf_search_regimes(dt, regime):
    # regimes that stress the root search: no sign change at all (y = 0), roots next to lmin (the pole of c(n): the
    # end-on-root rule of binf_solve) or next to lmax, an interval with lmax < lmin
    offs = np.concatenate([[0], np.cumsum([64] * 40 + [5, 1, 300, 17, 1100, 2])])
    n = int(offs[-1]); ng = len(offs) - 1
    xk, sj, q = inputs(n, dt)
    lam_g = (dt(0.5) + orc.uniform(ng, 12, dt)).astype(dt)
    sigma, delta = 0.3, 0.5
    if regime == "big_lambda":
        lam_g = (lam_g * dt(200)).astype(dt)
    elif regime == "tiny_lambda":
        lam_g = (lam_g * dt(1e-6)).astype(dt)
    elif regime == "tiny_delta":
        delta = 1e-4
    elif regime == "huge_delta":
        delta = 50.0
    elif regime == "tiny_shift":
        xk = (xk * dt(1e-3)).astype(dt); sj = (sj * dt(1e-3)).astype(dt); q = (q * dt(1e-3)).astype(dt)
        lam_g = (lam_g * dt(5)).astype(dt)
    h = sp.GroupNormL2(T(lam_g), None, offsets=T(offs))
    psi = sp.shifted(sp.shifted(h, T(xk), delta, sp.NormLinf(1.0)), T(sj))
    y = torch.empty(n, dtype=T(q).dtype, device=DEV)
    sp.prox_(y, psi, T(q), sigma)
    check_groupl2binf(N(y), xk, sj, q, offs, lam_g, sigma, delta, label=f"{dt.__name__} {regime}",
                      floor=1.0 if regime != "tiny_shift" else 1e-3)


# ------------------------------------------------- ψ(y) of the BInf forms (a20) ---
@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("n", [7, 4099, 300_001])
def test_indballl0binf_value_against_oracle(dt, n):
    """ShiftedIndBallL0BInf ψ(y) (shiftedIndBallL0BInf.jl:44-49): IndBallLinf(1.1Δ)(sj + y) + IndBallL0(r)(xk + sj + y),
    with the Float64 radius 1.1·Δ even for Float32 data (`1.1` is a Float64 literal)."""
    xk, sj, q = inputs(n, dt)
    delta = 0.7
    r = max(1, n // 10)
    psi = sp.shifted(sp.shifted(sp.IndBallL0(r), T(xk), delta, sp.NormLinf(1.0)), T(sj))
    # (1) at the prox outputs (once and twice shifted).  The reference clamps y (shiftedIndBallL0BInf.jl:91) AFTER the
    # top-r mask, so a dropped entry with |xk + sj| > Δ ends non-zero and the value can be Inf at the operator's own
    # prox output -- whatever it is, both sides must agree; want_value returns the same number
    psi1 = sp.shifted(sp.IndBallL0(r), T(xk), delta, sp.NormLinf(1.0))
    y1 = N(sp.prox(psi1, T(q), 1.0)).copy()
    z = np.zeros(n, dt)
    assert psi1(T(y1)) == orc.value_binf("indballl0", xk, z, y1, delta, r=r)
    yy = torch.empty(n, dtype=T(q).dtype, device=DEV)
    _, val = sp.prox_(yy, psi1, T(q), 1.0, want_value=True)
    assert val == orc.value_binf("indballl0", xk, z, N(yy), delta, r=r)
    y = N(sp.prox(psi, T(q), 1.0)).copy()
    assert psi(T(y)) == orc.value_binf("indballl0", xk, sj, y, delta, r=r)
    # with |xk| <= 0.3 Δ nothing is clamped: exactly min(r, n) non-zeros, inside the ball -> 0
    xs_ = (xk * dt(0.1)).astype(dt)
    psi2 = sp.shifted(sp.IndBallL0(r), T(xs_), delta, sp.NormLinf(1.0))
    qs_ = (q * dt(0.1)).astype(dt)
    y2 = N(sp.prox(psi2, T(qs_), 1.0)).copy()
    assert orc.value_binf("indballl0", xs_, z, y2, delta, r=r) == 0.0 == psi2(T(y2))
    # (2) feasible point (sj + y inside the ball; xk scaled so that the entries with v_i = 0, where w_i = -xk_i, are
    # inside as well): count <= r -> 0, r one short of the count -> Inf
    xk = (xk * dt(0.15)).astype(dt)
    keep = np.zeros(n, bool); keep[:: n // r + 1] = True  # at most r entries of v = xk + w stay non-zero
    sj = np.where(keep, sj, dt(0)).astype(dt)             # elsewhere sj = 0, y = -xk: w = -xk and v = 0 exactly
    psi = sp.shifted(sp.shifted(sp.IndBallL0(r), T(xk), delta, sp.NormLinf(1.0)), T(sj))
    w = (dt(delta) * orc.uniform(n, 21, dt, 2.0, -1.0)).astype(dt)
    y = np.where(keep, w - sj, -xk).astype(dt)
    ynz = np.count_nonzero((sj + y) + xk)
    assert 0 < ynz <= r and np.all(np.abs((sj + y).astype(np.float64)) <= 1.1 * float(dt(delta)))
    assert orc.value_binf("indballl0", xk, sj, y, delta, r=r) == 0.0 == psi(T(y))
    assert orc.value_binf("indballl0", xk, sj, y, delta, r=ynz) == 0.0
    psi_eq = sp.shifted(sp.shifted(sp.IndBallL0(ynz), T(xk), delta, sp.NormLinf(1.0)), T(sj))
    assert psi_eq(T(y)) == 0.0  # count == r
    if ynz > 1:
        psi_tight = sp.shifted(sp.shifted(sp.IndBallL0(ynz - 1), T(xk), delta, sp.NormLinf(1.0)), T(sj))
        assert orc.value_binf("indballl0", xk, sj, y, delta, r=ynz - 1) == np.inf == psi_tight(T(y))
    # (3) the edge of the ball: w = sj + y on either side of the Float64 radius 1.1Δ (strict test, Float64 compare)
    rad = 1.1 * float(dt(delta))
    inside = np.nextafter(dt(rad), dt(0)) if float(dt(rad)) > rad else dt(rad)  # largest R value <= rad
    outside = np.nextafter(inside, dt(np.inf))
    assert float(inside) <= rad < float(outside)
    for k, (wval, sign) in enumerate(((inside, 1), (inside, -1), (outside, 1), (outside, -1))):
        y3 = y.copy()
        i = (k * 97) % n
        y3[i] = dt(sign) * wval - sj[i]
        w = sj[i] + y3[i]
        want = orc.value_binf("indballl0", xk, sj, y3, delta, r=n)  # r = n: only the ball decides
        psin = sp.shifted(sp.shifted(sp.IndBallL0(n), T(xk), delta, sp.NormLinf(1.0)), T(sj))
        got = psin(T(y3))
        assert got == want, (k, float(w), rad, got, want)
        assert (want == np.inf) == (abs(float(w)) > rad)
    # (4) far outside
    y4 = y.copy(); y4[n // 2] += dt(10.0)
    assert psi(T(y4)) == np.inf == orc.value_binf("indballl0", xk, sj, y4, delta, r=r)


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("layout", ["g64", "ragged", "single"])
def test_groupl2binf_value_against_oracle(dt, layout):
    """ShiftedGroupNormL2Binf ψ(y) (shiftedGroupNormL2Binf.jl:34-39) at n >= 10^5 against orc.value_binf: inside the
    ball, on both sides of the Float64 radius 1.1Δ, and at the prox output (fused want_value)."""
    offs = {"g64": np.arange(0, 64 * 2001, 64), "ragged": ragged_offsets(600, 1500, seed=5),
            "single": np.array([0, 150_001])}[layout]
    n = int(offs[-1]); ng = len(offs) - 1
    xk, sj, q = inputs(n, dt)
    lam_g = (dt(0.5) + orc.uniform(ng, 12, dt)).astype(dt)
    sigma, delta = 0.3, 0.5
    h = sp.GroupNormL2(T(lam_g), None, offsets=T(offs))
    psi = sp.shifted(sp.shifted(h, T(xk), delta, sp.NormLinf(1.0)), T(sj))
    rtol = 1e-13 if dt == np.float64 else 3e-6
    # inside the ball: sj + y uniform in [-Δ, Δ]
    w = (dt(delta) * orc.uniform(n, 21, dt, 2.0, -1.0)).astype(dt)
    y = (w - sj).astype(dt)
    want = orc.value_binf("groupl2", xk, sj, y, delta, offs=offs, lam_g=lam_g)
    if np.isfinite(want):
        assert psi(T(y)) == pytest.approx(want, rel=rtol)
    else:  # rounding of w - sj + sj may leave the ball by an ulp: both sides must agree on that too
        assert psi(T(y)) == np.inf
    # shrink into the interior: finite on both sides, same value
    y = ((w * dt(0.9)) - sj).astype(dt)
    want = orc.value_binf("groupl2", xk, sj, y, delta, offs=offs, lam_g=lam_g)
    assert np.isfinite(want)
    assert psi(T(y)) == pytest.approx(want, rel=rtol)
    # the Float64 radius 1.1Δ, strict: largest R value inside, next one outside
    rad = 1.1 * float(dt(delta))
    inside = np.nextafter(dt(rad), dt(0)) if float(dt(rad)) > rad else dt(rad)
    outside = np.nextafter(inside, dt(np.inf))
    for wval, fin in ((inside, True), (-inside, True), (outside, False), (-outside, False)):
        y3 = y.copy()
        i = n - 3
        y3[i] = wval - sj[i]
        assert (abs(float(sj[i] + y3[i])) <= rad) == fin
        want = orc.value_binf("groupl2", xk, sj, y3, delta, offs=offs, lam_g=lam_g)
        got = psi(T(y3))
        assert np.isfinite(want) == fin
        assert got == pytest.approx(want, rel=rtol) if fin else got == np.inf
    # at the prox output, fused
    yy = torch.empty(n, dtype=T(q).dtype, device=DEV)
    _, val = sp.prox_(yy, psi, T(q), sigma, want_value=True)
    want = orc.value_binf("groupl2", xk, sj, N(yy), delta, offs=offs, lam_g=lam_g)
    assert np.isfinite(want) and val == pytest.approx(want, rel=rtol)


# ------------------------------------- uniform layouts: the fast path of GroupNormL2Binf ---
@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("m", [16, 32, 64, 128, 256])
@pytest.mark.parametrize("regime", ["base", "mid_lambda", "big_lambda", "tiny_delta", "huge_delta", "tiny_shift", "zeroed"])
def test_group_l2binf_uniform_layout(dt, m, regime):
    """n == ngroups * m with m = 8 L: spx_prox_groupl2binf_* takes the uniform-layout kernels (Float32 search, one
    Float64 evaluation + Halley step, the final pass as the acceptance test; failing rounds redone by the bracketing
    search).  ngroups is not a multiple of 32/L: the last round is partial."""
    ng = 1203
    n = ng * m
    offs = np.arange(0, n + 1, m)
    xk, sj, q = inputs(n, dt)
    lam_g = (dt(0.5) + orc.uniform(ng, 12, dt)).astype(dt)
    sigma, delta = 0.3, 0.5
    floor = 1.0
    if regime == "mid_lambda":
        lam_g = (lam_g * dt(20)).astype(dt)
    elif regime == "big_lambda":
        lam_g = (lam_g * dt(200)).astype(dt)
    elif regime == "tiny_delta":
        delta = 1e-4
    elif regime == "huge_delta":
        delta = 50.0
    elif regime == "tiny_shift":
        xk = (xk * dt(1e-3)).astype(dt); sj = (sj * dt(1e-3)).astype(dt); q = (q * dt(1e-3)).astype(dt)
        lam_g = (lam_g * dt(5)).astype(dt)
        floor = 1e-3
    elif regime == "zeroed":  # σλ_g > ||sol_g|| and |xk| <= Δ on half of the groups: fl*fm > 0 -> y_g = 0
        half = np.repeat(np.arange(ng) % 2 == 0, m)
        xk = np.where(half, xk * dt(0.2), xk).astype(dt)
        q = np.where(half, q * dt(0.05), q).astype(dt)
        sj = np.where(half, sj * dt(0.05), sj).astype(dt)
        lam_g = np.where(np.arange(ng) % 2 == 0, lam_g * dt(40), lam_g).astype(dt)
    h = sp.GroupNormL2(T(lam_g), None, offsets=T(offs))
    psi = sp.shifted(sp.shifted(h, T(xk), delta, sp.NormLinf(1.0)), T(sj))
    y = torch.empty(n, dtype=T(q).dtype, device=DEV)
    sp.prox_(y, psi, T(q), sigma)
    check_groupl2binf(N(y), xk, sj, q, offs, lam_g, sigma, delta, label=f"uniform m={m} {dt.__name__} {regime}", floor=floor)


@pytest.mark.parametrize("dt", DT)
def test_group_l2binf_uniform_candidates_that_are_not_uniform(dt):
    """n == ngroups * 64 but the groups are NOT all of 64 elements (the device-side check of offs must send the call
    to the generic kernels), and a uniform layout at a 16-byte-misaligned base (host-side test, generic kernels)."""
    m, ng = 64, 500
    sizes = np.full(ng, m)
    sizes[10] -= 3; sizes[11] += 3; sizes[300] += 20; sizes[301] -= 20
    offs = np.concatenate([[0], np.cumsum(sizes)])
    n = int(offs[-1])
    assert n == ng * m
    xk, sj, q = inputs(n, dt)
    lam_g = (dt(0.5) + orc.uniform(ng, 12, dt)).astype(dt)
    sigma, delta = 0.3, 0.5
    h = sp.GroupNormL2(T(lam_g), None, offsets=T(offs))
    psi = sp.shifted(sp.shifted(h, T(xk), delta, sp.NormLinf(1.0)), T(sj))
    y = torch.empty(n, dtype=T(q).dtype, device=DEV)
    sp.prox_(y, psi, T(q), sigma)
    check_groupl2binf(N(y), xk, sj, q, offs, lam_g, sigma, delta, label=f"almost uniform {dt.__name__}")
    # misaligned base: every vector starts one element into its allocation
    offs = np.arange(0, n + 1, m)
    pad = lambda a: T(np.concatenate([[0], a]).astype(dt))[1:]  # noqa: E731
    txk, tsj, tq = pad(xk), pad(sj), pad(q)
    ybuf = torch.empty(n + 1, dtype=tq.dtype, device=DEV)[1:]
    h = sp.GroupNormL2(T(lam_g), None, offsets=T(offs))
    psi = sp.shifted(sp.shifted(h, txk, delta, sp.NormLinf(1.0)), tsj)
    sp.prox_(ybuf, psi, tq, sigma)
    check_groupl2binf(N(ybuf), xk, sj, q, offs, lam_g, sigma, delta, label=f"uniform misaligned {dt.__name__}")
    # y aliases q (test/test_allocs.jl:108 calls prox!(y, ψ, y, σ))
    psi = sp.shifted(sp.shifted(h, T(xk), delta, sp.NormLinf(1.0)), T(sj))
    yq = T(q).clone()
    sp.prox_(yq, psi, yq, sigma)
    check_groupl2binf(N(yq), xk, sj, q, offs, lam_g, sigma, delta, label=f"uniform aliased {dt.__name__}")


# ------------------------------------------- batched top-r: the stream form (one CTA per problem) ---
@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("pn,r", [(4096, 100), (16_384, 1), (32_768, 700), (65_536, 1024), (65_536, 65_535), (131_072, 5000)])
@pytest.mark.parametrize("binf", [False, True])
def test_indballl0_batch_stream_form(dt, pn, r, binf):
    """Batches of >= 2 x 148 problems take topr_stream_kernel (16-bit keys in shared memory up to n = 65536, in the
    L2-resident scratch above): bit-exact against the oracle, including a tie-stress problem (magnitudes quantised to
    1/64), a constant problem, an all-zero problem and problems holding NaN / Inf (flagged for the radix kernel)."""
    nprob = 300
    n = nprob * pn
    xk, sj, q = inputs(n, dt)
    special = {3: "ties", 7: "const", 11: "zero", 13: "nan", 17: "inf", nprob - 1: "ties"}
    for p, kind in special.items():
        s = slice(p * pn, (p + 1) * pn)
        if kind == "ties":
            xk[s] = 0; sj[s] = 0
            q[s] = (np.round(q[s] * 64) / 64).astype(dt)
        elif kind == "const":
            xk[s] = 0; sj[s] = 0; q[s] = dt(1.25)
        elif kind == "zero":
            xk[s] = 0; sj[s] = 0; q[s] = 0
        elif kind == "nan":
            q[p * pn + 5] = np.nan; q[p * pn + pn - 2] = np.nan
        elif kind == "inf":
            q[p * pn + 9] = np.inf; q[p * pn + 77] = -np.inf
    h = sp.IndBallL0(r)
    psi = sp.shifted(h, T(xk), 1.0, sp.NormLinf(1.0), nprob=nprob) if binf else sp.shifted(h, T(xk), nprob=nprob)
    psi = sp.shifted(psi, T(sj))
    y = torch.empty(n, dtype=T(q).dtype, device=DEV)
    sp.prox_(y, psi, T(q), 1.0)
    got = N(y)
    check = sorted(set(special) | {0, 1, 2, nprob // 2, nprob - 2}) if pn > 20_000 else range(nprob)
    for p in check:
        s = slice(p * pn, (p + 1) * pn)
        ref = orc.prox_indballl0(xk[s], sj[s], q[s], r, delta=1.0 if binf else None)
        assert np.array_equal(got[s], ref, equal_nan=True), (p, special.get(p), np.flatnonzero(got[s] != ref)[:8])
    # every problem keeps exactly r entries (NaN problems excepted: NaN != NaN in the support test below)
    xs = (xk + sj).reshape(nprob, pn)
    kept = (got.reshape(nprob, pn) != np.clip(dt(0) - xs, -1.0, 1.0)) if binf else (got.reshape(nprob, pn) != dt(0) - xs)
    plain = [p for p in range(nprob) if p not in special]
    assert np.all(kept[plain].sum(1) <= r)
    if not binf:
        assert np.all(kept[plain].sum(1) == r)


def test_topr_digit_pick_with_counts_above_2p31():
    """The digit pick of the single-vector path scans 64-bit counts (a first-digit bin can hold every element of a
    vector of >= 2^31 elements, or of the histogram summed over the GPUs of a box)."""
    import ctypes as C
    from shiftedprox import _lib as L

    rng = np.random.default_rng(5)
    hist = np.zeros(2048, np.uint64)
    hist[2047] = 3_000_000_000  # > 2^31 in the top bin
    hist[2040] = 5_000_000_000
    hist[1000:1010] = rng.integers(1, 1 << 33, size=10).astype(np.uint64)
    hist[3] = 7
    cum = np.cumsum(hist[::-1].astype(object))  # counts from the top bin down
    for need in (1, 2_999_999_999, 3_000_000_000, 3_000_000_001, 8_000_000_000, int(cum[-1]) - 3, int(cum[-1])):
        k = int(np.searchsorted(np.array([int(c) for c in cum], dtype=object), need, side="left"))
        want_bin = 2047 - k
        want_above = int(cum[k - 1]) if k > 0 else 0
        b, a = C.c_int32(), C.c_int64()
        L.call("spx_selftest_topr_pick", sp.context(DEV), hist.ctypes.data_as(C.c_void_p), C.c_int64(need), C.byref(b),
               C.byref(a))
        assert (b.value, a.value) == (want_bin, want_above), (need, b.value, a.value, want_bin, want_above)


def test_group_offsets_are_validated_at_construction():
    """Device CSR offsets handed to GroupNormL2(offsets=...) are checked once when ψ is built (start at 0, monotone, end at n):
    the kernels trust them afterwards."""
    n = 1000
    x = T(np.zeros(n))
    good = T(np.array([0, 10, 10, 500, n], np.int64))
    sp.shifted(sp.GroupNormL2(T(np.ones(4)), None, offsets=good), x)
    for bad in ([1, 10, 500, n], [0, 600, 500, n], [0, 10, 500, n + 5], [0, 10, 500, n - 1]):
        with pytest.raises(ValueError, match="offsets"):
            sp.shifted(sp.GroupNormL2(T(np.ones(3)), None, offsets=T(np.array(bad, np.int64))), x)


# ------------------------------------------------------------- out-of-bounds writes (canaries) ---
@pytest.mark.parametrize("dt", DT)
def test_round2_kernels_write_only_inside_y(dt):
    """compute-sanitizer is not available on the GPU pool, so the kernels added in round 2 (uniform-layout and
    CTA-per-group GroupNormL2Binf, concurrent GroupNormL2 size classes, stream-form top-r, L1B2 with its stash in y) are
    run with y embedded in a larger buffer full of a sentinel: nothing outside y[0:n] may change."""
    pad = 256
    sentinel = dt(-12345.678)

    def embedded(n):
        buf = torch.full((n + 2 * pad,), float(sentinel), dtype=torch.float64 if dt == np.float64 else torch.float32, device=DEV)
        return buf, buf[pad:pad + n]

    def intact(buf, n):
        return bool((buf[:pad] == float(sentinel)).all()) and bool((buf[pad + n:] == float(sentinel)).all())

    # uniform layout, last round partial
    for m in (16, 64, 256):
        ng = 1203
        n = ng * m
        xk, sj, q = inputs(n, dt)
        lam_g = (dt(0.5) + orc.uniform(ng, 12, dt)).astype(dt)
        h = sp.GroupNormL2(T(lam_g), None, offsets=T(np.arange(0, n + 1, m)))
        buf, y = embedded(n)
        sp.prox_(y, sp.shifted(sp.shifted(h, T(xk), 0.5, sp.NormLinf(1.0)), T(sj)), T(q), 0.3)
        assert intact(buf, n), ("uniform Binf", m)
        buf, y = embedded(n)
        sp.prox_(y, sp.shifted(sp.shifted(h, T(xk)), T(sj)), T(q), 0.3)
        assert intact(buf, n), ("GroupL2", m)
    # ragged with groups up to 4096 and beyond (CTA-per-group, warp paths, concurrent size classes)
    offs = np.concatenate([[0], np.cumsum([1024, 1025, 7, 4096, 4097, 256, 257, 2049, 1, 3000, 5000, 64, 64])])
    n = int(offs[-1])
    xk, sj, q = inputs(n, dt)
    lam_g = (dt(0.5) + orc.uniform(len(offs) - 1, 12, dt)).astype(dt)
    h = sp.GroupNormL2(T(lam_g), None, offsets=T(offs))
    for psi in (sp.shifted(sp.shifted(h, T(xk), 0.5, sp.NormLinf(1.0)), T(sj)), sp.shifted(sp.shifted(h, T(xk)), T(sj))):
        buf, y = embedded(n)
        sp.prox_(y, psi, T(q), 0.3)
        assert intact(buf, n), type(psi).__name__
    # stream-form top-r: three shared-memory tiers and the L2-scratch tier
    for pn in (4096, 32_768, 65_536, 131_072):
        nprob = 300
        n = nprob * pn
        xk, sj, q = inputs(n, dt)
        buf, y = embedded(n)
        psi = sp.shifted(sp.shifted(sp.IndBallL0(97), T(xk), 1.0, sp.NormLinf(1.0), nprob=nprob), T(sj))
        sp.prox_(y, psi, T(q), 1.0)
        assert intact(buf, n), ("top-r stream", pn)
    # L1B2 (y is the stash of the search)
    n = 300_001
    xk, sj, q = inputs(n, dt)
    buf, y = embedded(n)
    psi = sp.shifted(sp.shifted(sp.NormL1(1.0), T(xk), 50.0, sp.NormL2(1.0)), T(sj))
    sp.prox_(y, psi, T(q), 0.1)
    assert intact(buf, n), "L1B2"

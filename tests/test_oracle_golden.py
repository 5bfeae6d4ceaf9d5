"""Pins the CPU oracle to every known-answer vector the reference's own tests hold
for the shifted prox path.  Citations are into /root/reference/test/.
The vectors are transcribed here (the reference tree does not exist on the GPU box)."""
import numpy as np
import pytest

from oracle import oracle as orc

NU = 1 / 9.1e4
Q5 = -NU * np.array([2631.441298528196, -533.9101219466443, 466.56156501426733,
                     1770.8953574224836, -2554.7769423950244])


def isapprox(a, b, rtol=None, atol=0.0):
    """Julia's isapprox for scalars (default rtol = sqrt(eps))."""
    rtol = np.sqrt(np.finfo(np.float64).eps) if rtol is None else rtol
    return abs(a - b) <= max(atol, rtol * max(abs(a), abs(b)))


# runtests.jl:113-126 -- unshifted RootNormLhalf prox KAT
def test_rootnormlhalf_kat():
    q = np.array([0.1097, 1.1287, -0.29, 1.2616])
    ytrue = np.array([0.0, 1.0893, -0.197463, 1.22444])
    y, _ = orc.prox_rootlhalf_unshifted(q, 0.7788, 0.1056)
    assert np.sum((y - ytrue) ** 2) <= 1e-11


# runtests.jl:127-157 -- GroupNormL2 prox/value vs per-group NormL2 prox
def test_groupnorml2_differential():
    rng = np.random.default_rng(0)
    x = rng.random(6); lam = rng.random(2); nu = rng.random()
    offs = [0, 3, 6]
    y, ysum = orc.prox_groupl2_unshifted(x, offs, lam, nu)
    ytrue = np.empty(6); ysumt = 0.0
    for g in range(2):
        xg = x[offs[g]:offs[g + 1]]
        nrm = np.linalg.norm(xg)
        ytrue[offs[g]:offs[g + 1]] = max(1 - nu * lam[g] / nrm, 0) * xg
        ysumt += lam[g] * nrm
    assert np.sum((y - ytrue) ** 2) <= 1e-11
    assert abs(ysum - ysumt) <= 1e-11


# runtests.jl:449-494 -- golden s_correct for the four trust-region operators (x=1, Δ=0.01, λ=1)
GOLD = {
    "l0": [-0.010000000000000, 0.005867144197216, -0.005127050164992, -0.010000000000000, 0.010000000000000],
    "l1": [-0.010000000000000, 0.005856155186227, -0.005138039175981, -0.010000000000000, 0.010000000000000],
    "lhalf": [-0.010000000000000, 0.005861665724748, -0.005132558825434, -0.010000000000000, 0.010000000000000],
}


@pytest.mark.parametrize("op", ["l0", "l1", "lhalf"])
def test_box_binf_golden(op):
    x = np.ones(5); s = np.zeros(5); delta = 0.01
    y = orc.prox_box(op, x, s, Q5, -delta, delta, 1.0, NU)
    for a, b in zip(y, GOLD[op]):
        assert isapprox(a, b)
    assert np.max(np.abs(y)) <= delta


def test_l1b2_golden():
    gold = [-0.006367076930786, 0.001288947922799, -0.001130889587543, -0.004285677352167, 0.006176811716709]
    x = np.ones(5); s = np.zeros(5); delta = 0.01
    y = orc.prox_l1b2(x, s, Q5, 1.0, NU, delta)
    for a, b in zip(y, gold):
        assert isapprox(a, b)
    assert np.linalg.norm(y) <= delta * (1 + 1e-12)


# runtests.jl:587-606 -- ShiftedGroupNormL2Binf built from NormL2 (one group)
def test_groupl2binf_single_group_golden():
    gold = [-0.010000000000000, 0.005862191941930, -0.005131948291800, -0.010000000000000, 0.010000000000000]
    x = np.ones(5); s = np.zeros(5)
    y = orc.prox_groupl2binf(x, s, Q5, [0, 5], np.array([1.0]), NU, 0.01)
    for a, b in zip(y, gold):
        assert isapprox(a, b)


# runtests.jl:648-705 -- two groups
def test_groupl2binf_two_groups_golden():
    lam = np.array([0.396767474230670, 0.538816734003357])
    q = np.array([-0.649013765191241, 1.181166041965532, -0.758453297283692, -1.109613038501522,
                  -0.845551240007797, -0.572664866457950])
    gold = [-0.01, 0.01, -0.01, -0.01, -0.01, -0.01]
    y = orc.prox_groupl2binf(np.ones(6), np.zeros(6), q, [0, 3, 6], lam, 0.419194514403295, 0.01)
    for a, b in zip(y, gold):
        assert isapprox(a, b)


# runtests.jl:814-843 -- L1Box vs clamp(prox_NormL1(xk+q), xk±Δ) - xk, once and twice shifted
def test_l1box_differential():
    rng = np.random.default_rng(1)
    for _ in range(50):
        n = 4; delta = 2 * rng.random(); q = 2 * (rng.random(n) - 0.5); nu = rng.random()
        xk = rng.random(n) - 0.5
        st = lambda v, t: np.sign(v) * np.maximum(np.abs(v) - t, 0)
        p1 = np.minimum(np.maximum(st(xk + q, nu), xk - delta), xk + delta) - xk
        p2 = orc.prox_box("l1", xk, np.zeros(n), q, -delta, delta, 1.0, nu)
        # reference as written clamps v=x+s+y to [x-Δ, x+Δ], i.e. s+y to ±Δ with s = 0
        assert np.allclose(p1, p2, rtol=1.5e-8, atol=0)
        sj = rng.random(n) - 0.5
        p1 = np.minimum(np.maximum(st(xk + sj + q, nu), xk - delta), xk + delta) - (xk + sj)
        p2 = orc.prox_box("l1", xk, sj, q, -delta, delta, 1.0, nu)
        assert np.allclose(p1, p2, rtol=1.5e-8, atol=1e-15)


# testsbox.jl:1-99 -- 9 branch-coverage cases per Box prox (l=0, u=3, s=-1, σ=1), atol 1e-2
BOX9 = {
    "l0": dict(q=[5, 5, 5, 0, 0, 0, 3, 3, 3], x=[1, -1, -1, 1, -1, -1, 1, -1, -1],
               lam=[1, 5, 3, 1, 2, 1, 1, 1, 0.1], sol=[4, 2, 4, 1, 2, 1, 3, 2, 3]),
    "l1": dict(q=[0.5, 5, 3, -2, 4, 1, 1, 7, 4], x=[1, -4, -2, -1, -5, -3, 3, -2, 1],
               lam=[1] * 9, sol=[1, 4, 3, 1, 4, 2, 1, 4, 3]),
    "lhalf": dict(q=[5, 5, 5, 2, 0, 1, 0, 3, 3], x=[1, -1, -1, 1, 1, -1, -1, -1, -1],
                  lam=[1, 10, 1, 1, 1, 1, 1, 0.5, 1], sol=[4, 2, 4, 1.6054, 1, 2, 1, 2.702, 2]),
}


@pytest.mark.parametrize("op", ["l0", "l1", "lhalf"])
def test_box_prox_nine_cases(op):
    c = BOX9[op]
    for i in range(9):
        y = orc.prox_box(op, np.array([float(c["x"][i])]), np.array([-1.0]), np.array([float(c["q"][i])]),
                         np.array([0.0]), np.array([3.0]), float(c["lam"][i]), 1.0)
        assert abs(y[0] - c["sol"][i]) <= 1e-2, (op, i, y[0], c["sol"][i])


# testsbox.jl:101-304 -- 14 exact cases per Box iprox (l=-2, u=1, s=-1)
IPROX14 = {
    "l0": dict(d=[0, 0, 0, 0, 0, 2, 2, 2, 2, 2, 2, -2, -2, -2],
               g=[0, 0, 2, 2, -2, 1, 0, 1, 10, -10, 4, -10, 10, -4],
               x=[0, -10, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0],
               lam=[1, 1, 1, 10, 1, 1, 0.1, 10, 1, 1, 10, 1, 1, 10],
               sol=[1, 0, -1, 1, 2, -0.5, 0, 1, -1, 2, 1, 2, -1, 1]),
    "l1": dict(d=[0, 0, 0, 0, 0, 2, 2, 2, 2, 2, 2, -2, -2, -2],
               g=[0.5, 0.5, 0.5, 2, -2, 0, 1, 1, -1, 1, 1, 0, 1, 1],
               x=[0, 4, -2, 0, 0, 4, -2, 1, 0.5, 0.5, 3, 1, 1, 1],
               lam=[1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 10, 1],
               sol=[1, -1, 2, -1, 2, -0.5, 0, 0, 0.5, 0, -1, 2, 0, -1]),
}


@pytest.mark.parametrize("op", ["l0", "l1"])
def test_box_iprox_fourteen_cases_exact(op):
    c = IPROX14[op]
    for i in range(14):
        y = orc.iprox_box(op, np.array([float(c["x"][i])]), np.array([-1.0]), np.array([float(c["g"][i])]),
                          np.array([float(c["d"][i])]), np.array([-2.0]), np.array([1.0]), float(c["lam"][i]))
        assert y[0] == c["sol"][i], (op, i, y[0], c["sol"][i])


# partial_prox.jl:1-74 -- `selected = 1:2:n`
@pytest.mark.parametrize("op", ["l0", "l1", "lhalf"])
def test_partial_prox(op):
    rng = np.random.default_rng(2)
    n = 5; lam = 3.14
    x = rng.random(n); s = rng.random(n); q = rng.random(n) - 0.5
    if op == "lhalf":
        l = -0.5 * np.ones(n); u = 0.5 * np.ones(n); lb, ub = -0.5, 0.5
    else:
        l = np.zeros(n); u = np.ones(n); lb, ub = l, u
    sel = np.arange(0, n, 2)
    y = orc.prox_box(op, x, s, q, lb, ub, lam, 1.0)
    z = orc.prox_box(op, x, s, q, lb, ub, lam, 1.0, selected=sel)
    p = np.minimum(np.maximum(q, l - s), u - s)
    for i in range(n):
        assert z[i] == (y[i] if i in sel else p[i])
    if op in ("l0", "l1"):
        for d in (np.ones(n), -np.ones(n), np.zeros(n)):
            y = orc.iprox_box(op, x, s, q, d, l, u, lam)
            z = orc.iprox_box(op, x, s, q, d, l, u, lam, selected=sel)
            p = [orc.iprox_zero(d[i], q[i], l[i] - s[i], u[i] - s[i]) for i in range(n)]
            for i in range(n):
                assert z[i] == (y[i] if i in sel else p[i])
        # unboxed iprox: assertion on d == 0, and iprox(ψ,q,d·1) == prox(ψ,q,σ=d) exactly (:58-72)
        fi = orc.iprox_l0 if op == "l0" else orc.iprox_l1
        fp = orc.prox_l0 if op == "l0" else orc.prox_l1
        with pytest.raises(AssertionError):
            fi(x, np.zeros(n), q, np.zeros(n), lam)
        for dv in (1.0, 2.0):
            yi = fi(x, np.zeros(n), q, dv * np.ones(n), lam)
            zp = fp(x, np.zeros(n), q, lam, dv)
            for i in sel:
                assert zp[i] == yi[i]


# value identities: runtests.jl:175-194, 443-447, 517-521
@pytest.mark.parametrize("kind", ["l0", "l1", "lhalf"])
def test_value_identities(kind):
    rng = np.random.default_rng(3)
    lam = 1.2; x = np.ones(3); y = rng.random(3)
    h = {"l1": lambda v: lam * np.sum(np.abs(v)), "l0": lambda v: lam * np.count_nonzero(v),
         "lhalf": lambda v: lam * np.sum(np.sqrt(np.abs(v)))}[kind]
    assert orc.value_plain(kind, x, np.zeros(3), np.zeros(3), lam) == pytest.approx(h(x), rel=1e-15)
    assert orc.value_plain(kind, x, np.zeros(3), y, lam) == pytest.approx(h(x + y), rel=1e-15)
    # Box: inside -> h(x+y), outside -> Inf
    n = 5; x = np.ones(n); delta = 0.01
    y = rng.random(n); y *= delta / np.max(np.abs(y)) / 2
    assert orc.value_box(kind, x, np.zeros(n), y, -delta, delta, 1.0) == pytest.approx(
        {"l1": np.sum(np.abs(x + y)), "l0": 5.0, "lhalf": np.sum(np.sqrt(np.abs(x + y)))}[kind], rel=1e-15)
    assert orc.value_box(kind, x, np.zeros(n), 3 * y, -delta, delta, 1.0) == np.inf


def test_value_l1b2_and_binf():
    rng = np.random.default_rng(4)
    n = 5; x = np.ones(n); delta = 0.01
    y = rng.random(n); y *= delta / np.linalg.norm(y) / 2
    assert orc.value_l1b2(x, np.zeros(n), y, 1.0, delta) == pytest.approx(np.sum(np.abs(x + y)), rel=1e-15)
    assert orc.value_l1b2(x, np.zeros(n), 3 * y, 1.0, delta) == np.inf
    y = rng.random(n); y *= delta / np.max(np.abs(y)) / 2
    assert orc.value_binf("indballl0", x, np.zeros(n), y, delta, r=5) == 0.0
    assert orc.value_binf("indballl0", x, np.zeros(n), y, delta, r=4) == np.inf
    assert orc.value_binf("indballl0", x, np.zeros(n), 3 * y, delta, r=5) == np.inf
    lam = np.array([0.3, 0.7]); x = np.ones(6); y = rng.random(6); y *= delta / np.max(np.abs(y)) / 2
    v = x + y
    expect = lam[0] * np.linalg.norm(v[:3]) + lam[1] * np.linalg.norm(v[3:])
    assert orc.value_binf("groupl2", x, np.zeros(6), y, delta, offs=[0, 3, 6], lam_g=lam) == pytest.approx(expect, rel=1e-15)
    assert orc.value_binf("groupl2", x, np.zeros(6), 3 * y, delta, offs=[0, 3, 6], lam_g=lam) == np.inf
    assert orc.value_groupl2(x, np.zeros(6), y, [0, 3, 6], lam) == pytest.approx(expect, rel=1e-15)


# runtests.jl:244-251, 318-329 -- ShiftedGroupNormL2 prox vs NormL2 prox of (q+x) minus x
def test_shifted_groupl2_differential():
    rng = np.random.default_rng(5)
    x = rng.random(6); s = rng.random(6); q = rng.random(6); lam = rng.random(2); nu = rng.random()
    offs = [0, 3, 6]
    y = orc.prox_groupl2(x, s, q, offs, lam, nu)
    ytrue = np.empty(6)
    for g in range(2):
        sl = slice(offs[g], offs[g + 1])
        v = q[sl] + x[sl] + s[sl]
        ytrue[sl] = max(1 - nu * lam[g] / np.linalg.norm(v), 0) * v - (x[sl] + s[sl])
    assert np.linalg.norm(y - ytrue) <= 1e-11


# Source-text checks for the operators the reference leaves "# TODO" (parity unpinned).
def test_unpinned_source_text_semantics():
    rng = np.random.default_rng(6)
    n = 1000
    x = 4 * rng.random(n) - 2; s = rng.random(n) - 0.5; q = 4 * rng.random(n) - 2
    lam, sig = 1.0, 0.1
    t = (-x) - s
    assert np.array_equal(orc.prox_l1(x, s, q, lam, sig), np.minimum(np.maximum(t, q - lam * sig), q + lam * sig))
    xps = x + s
    assert np.array_equal(orc.prox_l0(x, s, q, lam, sig), np.where(np.abs(xps + q) <= np.sqrt(2 * lam * sig), -xps, q))
    # top-r: ties resolved towards the lowest index; kept entries are (xs+q)-xs, not q
    z = np.array([1.0, -2.0, 2.0, 0.5, -2.0, 2.0])
    y = orc.prox_indballl0(np.zeros(6), np.zeros(6), z, 2)
    assert np.array_equal(y, [0, -2.0, 2.0, 0, 0, 0])
    y = orc.prox_indballl0(np.zeros(6), np.zeros(6), z, 3)
    assert np.array_equal(y, [0, -2.0, 2.0, 0, -2.0, 0])
    xs = x[:6] + s[:6]
    y = orc.prox_indballl0(x[:6], s[:6], q[:6], 2)
    zz = xs + q[:6]
    keep = np.argsort(-np.abs(zz), kind="stable")[:2]
    exp = -xs.copy(); exp[keep] = zz[keep] - xs[keep]
    assert np.array_equal(y, exp)
    yb = orc.prox_indballl0(x[:6], s[:6], q[:6], 2, delta=0.3)
    assert np.array_equal(yb, np.clip(exp, -0.3, 0.3))
    # NaN sorts as the largest magnitude (isless)
    z = np.array([1.0, np.nan, 3.0])
    y = orc.prox_indballl0(np.zeros(3), np.zeros(3), z, 1)
    assert np.isnan(y[1]) and y[0] == 0 and y[2] == 0


def test_julia_minmax_semantics():
    assert np.signbit(orc.prox_zero(-0.0, 0.0, 1.0)) == False  # max(-0.0, 0.0) = 0.0
    assert np.signbit(orc.prox_zero(0.0, -1.0, -0.0)) == True  # min(0.0, -0.0) = -0.0
    assert np.isnan(orc.prox_zero(np.nan, 0.0, 1.0))
    assert orc.iprox_zero(0.0, 0.0, -1.0, 1.0) == 0.0
    assert orc.iprox_zero(0.0, 2.0, -1.0, 1.0) == -1.0
    assert orc.iprox_zero(-2.0, 1.0, -1.0, 2.0) == 2.0
    assert orc.iprox_zero(2.0, 1.0, -1.0, 2.0) == -0.5


def test_float32_lhalf_promotions():
    # Float32 data: power/acos/cos run in Float64, result rounded to Float32 on store
    x = np.array([0.3, -0.7, 1.1], dtype=np.float32); s = np.zeros(3, np.float32)
    q = np.array([1.0, -2.0, 0.01], dtype=np.float32)
    y = orc.prox_lhalf(x, s, q, 0.7, 0.2)
    assert y.dtype == np.float32
    z = (q + x).astype(np.float64); nl = np.float32(np.float32(0.2) * np.float32(0.7))
    p = 54 ** (1 / 3) * (2 * float(nl)) ** (2 / 3) / 4
    for i in range(3):
        if abs(z[i]) <= p:
            exp = np.float32(0) - x[i]
        else:
            phi = np.arccos(float(nl / np.float32(4)) * (float(np.float32(abs(z[i])) / np.float32(3))) ** -1.5)
            coef = np.float32(2) * np.float32(np.sign(z[i])) / np.float32(3) * np.float32(abs(z[i]))
            exp = np.float32(float(coef) * (1 + np.cos(2 * np.pi / 3 - 2 * phi / 3))) - x[i]
        assert abs(float(y[i]) - float(exp)) <= 2 * np.spacing(np.float32(abs(exp)))


def test_spectral_stage_matches_the_reference_identities():
    """runtests.jl:945-946,963-964 (Rank: the thresholded singular values are NormL0's prox of S),
    :1163-1164,1181-1182 (Nuclearnorm: NormL1's prox of S), :1064-1066 (Cappedl1, θ = 1, λ = 10 on a diagonal with
    entries in [0, 1.75): NormL0's prox as well).  The oracle's stage scales column i of U by exactly that value."""
    rng = np.random.default_rng(0)
    for lam, gamma in ((10.0, 10.0), (1.0, 5.0)):
        S = np.sort(rng.random(11) * (20.0 if lam == 10.0 else 8.0))[::-1].copy()
        U = np.eye(11)
        hard = np.where(np.abs(S) > np.sqrt(2 * lam * gamma), S, 0.0)  # prox of λ‖·‖₀ with step γ
        soft = np.sign(S) * np.maximum(np.abs(S) - lam * gamma, 0.0)  # prox of λ‖·‖₁
        Ur, Sr = orc.spectral_threshold("rank", U, S, lam, gamma)
        assert np.array_equal(np.diag(Ur), hard) and np.array_equal(Sr, S)
        Un, Sn = orc.spectral_threshold("nuclear", U, S, lam, gamma)
        assert np.array_equal(np.diag(Un), soft) and np.array_equal(Sn, soft)
    st1 = rng.random(10)
    S = st1 + st1 ** 2 + st1 / 2  # the diagonal of runtests.jl:1056-1066, all below sqrt(2·10·10)
    Uc, Sc = orc.spectral_threshold("cappedl1", np.eye(10), S, 10.0, 10.0, theta=1.0)
    assert np.array_equal(Sc, np.where(np.abs(S) > np.sqrt(2 * 10.0 * 10.0), S, 0.0))
    assert np.array_equal(np.diag(Uc), Sc)

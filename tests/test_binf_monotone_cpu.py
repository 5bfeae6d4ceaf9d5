"""Properties of froot (shiftedGroupNormL2Binf.jl:87-93) the GroupNormL2Binf root search leans on, pinned on the
reference's own formula (numpy restatement, CPU only): froot is increasing in n with slope >= 1 (the norm it
subtracts falls with n), so a bracket end whose residual is within k ulps of n lies within k ulps of the root --
the rule by which the kernel steps over such an end instead of bisecting towards the far one (csrc/spx_group.cu,
binf_solve) -- and the reference's "no root" test f(lmin) f(lmax) > 0 (:102) means f(lmin) > 0 or f(lmax) < 0."""
import numpy as np
import pytest

from oracle import oracle as orc


def froot(n, sol, xk, lam, sigma, delta):
    c = n / (sigma * (n - sigma * lam))
    t = sol / sigma - c * xk
    w = sigma * np.sign(t) * np.maximum(0.0, np.abs(t) - delta * c) - sol
    return n - np.linalg.norm(w)


@pytest.mark.parametrize("m", [1, 5, 64, 700])
@pytest.mark.parametrize("delta", [0.01, 0.5, 3.0])
def test_froot_is_increasing_between_lmin_and_lmax(m, delta):
    rng = np.random.default_rng(m)
    sigma = 0.3
    for _ in range(20):
        xk = 4 * rng.random(m) - 2
        sol = (4 * rng.random(m) - 2) + xk + (rng.random(m) - 0.5)
        lam = 0.5 + rng.random()
        sl = sigma * lam
        lmin = sl * (1 + np.finfo(float).eps)
        lmax = np.linalg.norm(sol) + sigma * (np.linalg.norm(sol / sigma) + lam * np.linalg.norm(xk)) + 10.0
        # geometric grid from just above lmin (where c(n) ~ 1/eps) to lmax
        grid = sl + (lmin - sl) * np.geomspace(1.0, (lmax - sl) / (lmin - sl), 400)
        f = np.array([froot(n, sol, xk, lam, sigma, delta) for n in grid])
        scale = np.abs(grid) + np.abs(grid - f)
        assert np.all(np.diff(f) >= np.diff(grid) - 1e-12 * scale[1:])  # slope >= 1 up to the rounding of the norm


def test_oracle_prox_is_zero_exactly_when_there_is_no_sign_change():
    # a shift far inside the threshold: f(lmin) > 0, the reference returns y = -(xk + sj) (:103-104 with y = 0)
    m = 16
    xk = np.full(m, 1e-3); sj = np.zeros(m); q = np.full(m, 1e-3)
    offs = np.array([0, m]); lam = np.array([5.0])
    y = orc.prox_groupl2binf(xk, sj, q, offs, lam, 0.3, 0.5)
    assert np.array_equal(y, -(xk + sj))
    sol = (q + xk) + sj
    lmin = 0.3 * 5.0 * (1 + np.finfo(float).eps)
    assert froot(lmin, sol, xk, 5.0, 0.3, 0.5) > 0

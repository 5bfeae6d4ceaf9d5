"""the thresholding stage of ShiftedRank / ShiftedNuclearnorm / ShiftedCappedl1 given an SVD
(shiftedRank.jl:72-82, shiftedNuclearnorm.jl:72-78, shiftedCappedl1.jl:71-83) -- bit-exact against the oracle's
restatement of the three loops, and the whole prox! (library SVD + library GEMM around the stage) against numpy."""
import numpy as np
import pytest
import torch

from gpu_util import DEV, N, T, orc, sp
from shiftedprox import spectral

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("kind", ["rank", "nuclear", "cappedl1"])
@pytest.mark.parametrize("m,k", [(1, 1), (7, 3), (300, 40), (1025, 129)])
def test_spectral_threshold_stage_bit_exact(dt, kind, m, k):
    rng = np.random.default_rng(3)
    U = rng.standard_normal((m, k)).astype(dt)
    S = np.sort(np.abs(rng.standard_normal(k)) * 2.0)[::-1].astype(dt)
    lam, sigma, theta = 0.9, 0.4, 1.1
    S[k // 2] = dt(np.sqrt(dt(2) * dt(lam) * dt(sigma)))  # exactly on the Rank threshold: `<=` zeroes the column
    Uref, Sref = orc.spectral_threshold(kind, U, S, lam, sigma, theta)
    tU = T(np.asfortranarray(U).T.copy()).T  # column-major on the device
    tS = T(S.copy())
    spectral.spectral_threshold_(tU, tS, kind, lam, sigma, theta)
    assert np.array_equal(N(tU), Uref)
    assert np.array_equal(N(tS), Sref)


@pytest.mark.parametrize("kind", ["rank", "nuclear", "cappedl1"])
def test_spectral_prox_against_numpy(kind):
    """Whole prox! on a 60 x 40 matrix: numpy SVD + the reference formula in Float64 vs library SVD/GEMM + our stages."""
    m, n = 60, 40
    rng = np.random.default_rng(11)
    xk, sj, q = (rng.standard_normal(m * n) * s for s in (1.0, 0.3, 1.0))
    lam, sigma, theta = 0.8, 0.5, 1.5
    y = torch.empty(m * n, dtype=torch.float64, device=DEV)
    spectral.prox_spectral_(y, kind, (m, n), T(xk), T(sj), T(q), lam, sigma, theta)
    sol = (q + xk) + sj
    A = sol.reshape(n, m).T
    U, S, Vt = np.linalg.svd(A, full_matrices=False)
    U2, _ = orc.spectral_threshold(kind, U, S, lam, sigma, theta)
    ref = (U2 @ Vt).T.reshape(-1) - (xk + sj)
    assert np.max(np.abs(N(y) - ref)) <= 1e-11 * max(1.0, np.max(np.abs(ref)))


def test_spectral_sol_and_finish_are_bit_exact():
    n = 100_003
    rng = np.random.default_rng(1)
    xk, sj, q = (rng.standard_normal(n) for _ in range(3))
    a = torch.empty(n, dtype=torch.float64, device=DEV)
    spectral.spectral_sol_(a, T(xk), T(sj), T(q))
    assert np.array_equal(N(a), (q + xk) + sj)
    y = torch.empty_like(a)
    spectral.spectral_finish_(y, a, T(xk), T(sj))
    assert np.array_equal(N(y), ((q + xk) + sj) - (xk + sj))

"""Committed fixtures (tests/golden/oracle_r1.npz, made by tools/make_golden.py): the oracle must still
reproduce them bit for bit (CPU), and the CUDA path must agree with them on the GPU with the same criteria as
the live-oracle parity tests.  The fixtures are oracle outputs (the Julia reference cannot run here); the oracle
itself is pinned to the reference's known-answer tests in tests/test_oracle_golden.py."""
import importlib.util
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "..", "tools", "make_golden.py"))
mg = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mg)
GOLD = np.load(os.path.join(HERE, "golden", "oracle_r1.npz"))


def test_oracle_reproduces_the_committed_fixtures():
    now = mg.build()
    assert sorted(now) == sorted(GOLD.files)
    for k in GOLD.files:
        assert np.array_equal(now[k], GOLD[k], equal_nan=True), k


@pytest.mark.gpu
@pytest.mark.parametrize("suf", ["f64", "f32"])
def test_cuda_path_against_the_committed_fixtures(suf):
    import torch
    from gpu_util import DEV, N, T, sp

    dt = np.float64 if suf == "f64" else np.float32
    eps = np.finfo(dt).eps
    xk, sj, q, l, ub, d, offs, lam_g = mg.inputs(dt)
    dpos = np.abs(d) + dt(0.25)
    lam, sigma, delta = mg.LAM, mg.SIGMA, mg.DELTA
    scale = np.abs(xk) + np.abs(sj) + np.abs(q) + 1.0
    sh = lambda psi: sp.shifted(psi, T(sj))  # noqa: E731
    exact = lambda got, key: np.array_equal(N(got), GOLD[f"{key}_{suf}"], equal_nan=True)  # noqa: E731
    close = lambda got, key, tol: np.all(np.abs(N(got).astype(np.float64) - GOLD[f"{key}_{suf}"]) <= tol * scale)  # noqa: E731
    assert exact(sp.prox(sh(sp.shifted(sp.NormL1(lam), T(xk))), T(q), sigma), "prox_l1")
    assert exact(sp.prox(sh(sp.shifted(sp.NormL0(lam), T(xk))), T(q), sigma), "prox_l0")
    assert close(sp.prox(sh(sp.shifted(sp.RootNormLhalf(lam), T(xk))), T(q), sigma), "prox_lhalf", 4 * eps)
    assert exact(sp.iprox(sh(sp.shifted(sp.NormL1(lam), T(xk))), T(q), T(dpos)), "iprox_l1")
    assert exact(sp.iprox(sh(sp.shifted(sp.NormL0(lam), T(xk))), T(q), T(dpos)), "iprox_l0")
    for name, h in (("l1", sp.NormL1(lam)), ("l0", sp.NormL0(lam))):
        psi = sh(sp.shifted(h, T(xk), T(l), T(ub)))
        assert exact(sp.prox(psi, T(q), sigma), f"prox_{name}box")
        assert exact(sp.iprox(psi, T(q), T(d)), f"iprox_{name}box")
    psi = sh(sp.shifted(sp.RootNormLhalf(lam), T(xk), T(l), T(ub)))
    assert close(sp.prox(psi, T(q), sigma), "prox_lhalfbox", 4 * eps)
    psi = sh(sp.shifted(sp.NormL1(lam), T(xk), float(GOLD[f"l1b2_delta_{suf}"][0]), sp.NormL2(1.0)))
    assert close(sp.prox(psi, T(q), sigma), "prox_l1b2", 64 * eps)
    hg = sp.GroupNormL2(T(lam_g), None, offsets=T(offs))
    assert close(sp.prox(sh(sp.shifted(hg, T(xk))), T(q), 0.3), "prox_groupl2", 8 * eps)
    assert close(sp.prox(sh(sp.shifted(hg, T(xk), delta, sp.NormLinf(1.0))), T(q), 0.3), "prox_groupl2binf",
                 1e-9 if suf == "f64" else 2e-4)
    # fused solver step: s and xsy like the prox! of the same h, the three scalars to the sums' accuracy
    for name, h in (("l1", sp.NormL1(lam)), ("l0", sp.NormL0(lam)), ("lhalf", sp.RootNormLhalf(lam))):
        for tag, psi in (("step", sh(sp.shifted(h, T(xk)))), ("stepbox", sh(sp.shifted(h, T(xk), T(l), T(ub))))):
            s_ = torch.empty(len(q), dtype=T(q).dtype, device=DEV)
            xsy = torch.empty_like(s_)
            _, res = sp.step_(s_, psi, T(q), mg.NU, xsy=xsy)
            if name == "lhalf":
                assert close(s_, f"{tag}_{name}_s", 4 * eps) and close(xsy, f"{tag}_{name}_xsy", 4 * eps)
            else:
                assert exact(s_, f"{tag}_{name}_s") and exact(xsy, f"{tag}_{name}_xsy")
            gpsi, gsn, ggd = GOLD[f"{tag}_{name}_scalars_{suf}"]
            rel = (1e-12 if suf == "f64" else 1e-5) if name != "lhalf" else (1e-9 if suf == "f64" else 1e-4)
            assert res.psi == pytest.approx(gpsi, rel=rel) and res.snorm == pytest.approx(gsn, rel=rel)
            assert abs(res.gdots - ggd) <= rel * (abs(ggd) + float(np.sum(np.abs(q.astype(np.float64)) ** 2)))
    assert exact(sp.prox(sh(sp.shifted(sp.IndBallL0(31), T(xk))), T(q), 1.0), "prox_indballl0")
    assert exact(sp.prox(sh(sp.shifted(sp.IndBallL0(31), T(xk), 1.0, sp.NormLinf(1.0))), T(q), 1.0), "prox_indballl0binf")
    torch.cuda.synchronize()

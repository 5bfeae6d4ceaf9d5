"""CPU check of the oracle's solver-step composition (oracle.solver_step) on a hand-computed case, and of its
consistency with the oracle's own prox! / ψ(y) (the step has no reference fixture: its caller lives in
RegularizedOptimization.jl, outside /root/reference)."""
import numpy as np
import pytest

from oracle import oracle as orc


def test_solver_step_hand_computed_l1():
    xk = np.array([1.0, -2.0, 0.5]); sj = np.zeros(3); grad = np.array([1.0, 1.0, -4.0])
    s, xsy, psi, snorm, gdots = orc.solver_step("l1", xk, sj, grad, 1.0, 0.5)
    # q = -0.5 grad = [-0.5, -0.5, 2];  s = min(max(-xk, q - 0.5), q + 0.5)   (shiftedNormL1.jl:46-51)
    assert np.array_equal(s, [-1.0, 0.0, 1.5])
    assert np.array_equal(xsy, [0.0, -2.0, 2.0])
    assert psi == 4.0 and snorm == pytest.approx(np.sqrt(3.25)) and gdots == -7.0


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("op", ["l1", "l0", "lhalf"])
def test_solver_step_is_prox_plus_value(dt, op):
    n = 1237
    xk = orc.uniform(n, 0, dt, 4.0, -2.0); sj = orc.uniform(n, 1, dt, 1.0, -0.5); grad = orc.uniform(n, 2, dt, 4.0, -2.0)
    l = -(dt(0.25) + orc.uniform(n, 3, dt)); u = dt(0.25) + orc.uniform(n, 4, dt)
    lam, nu = 0.9, 0.3
    q = (dt(-dt(nu)) * grad).astype(dt)
    s, xsy, psi, _, _ = orc.solver_step(op, xk, sj, grad, lam, nu)
    ref = {"l1": orc.prox_l1, "l0": orc.prox_l0, "lhalf": orc.prox_lhalf}[op](xk, sj, q, lam, nu)
    assert np.array_equal(s, ref) and psi == orc.value_plain(op, xk, sj, ref, lam)
    sb, _, psib, _, _ = orc.solver_step(op, xk, sj, grad, lam, nu, l, u, np.arange(0, n, 2))
    refb = orc.prox_box(op, xk, sj, q, l, u, lam, nu, np.arange(0, n, 2))
    assert np.array_equal(sb, refb) and np.isfinite(psib)
    assert np.all(sj + sb >= l - 1e-3) and np.all(sj + sb <= u + 1e-3)

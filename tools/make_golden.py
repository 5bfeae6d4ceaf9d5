#!/usr/bin/env python
"""Writes tests/golden/oracle_r1.npz: outputs of the CPU oracle (oracle/) on small seeded inputs, one entry per
operator and element type.  The reference itself (Julia) cannot run in this image, so these are NOT reference
outputs: they freeze the oracle (which tests/test_oracle_golden.py pins to the reference's own known-answer
tests) so that a later edit of the oracle or of a kernel shows up as a diff against committed numbers.

    python tools/make_golden.py          # regenerate (commit the result together with the oracle change)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402

N = 257
LAM, SIGMA, DELTA = 0.8, 0.1, 0.5
NU = 0.35  # step length of the fused solver step


def inputs(dt):
    u = lambda k, **kw: orc.uniform(N, k, dt, **kw)  # noqa: E731
    xk, sj, q = u(0, scale=4.0, shift=-2.0), u(1, shift=-0.5), u(2, scale=4.0, shift=-2.0)
    l = (-(dt(0.25) + u(3))).astype(dt)
    ub = (dt(0.25) + u(4)).astype(dt)
    d = dt(0.5) + u(5)
    b = orc.uniform(N, 6, np.float64)
    d = np.where(b < 0.1, -d, d)
    d = np.where((b >= 0.1) & (b < 0.2), dt(0), d).astype(dt)
    offs = np.array([0, 1, 4, 12, 76, 77, 140, 257], dtype=np.int64)
    lam_g = (dt(0.5) + orc.uniform(len(offs) - 1, 12, dt)).astype(dt)
    return xk, sj, q, l, ub, d, offs, lam_g


def build():
    out = {}
    for dt, suf in ((np.float64, "f64"), (np.float32, "f32")):
        xk, sj, q, l, ub, d, offs, lam_g = inputs(dt)
        dpos = np.abs(d) + dt(0.25)
        out[f"prox_l1_{suf}"] = orc.prox_l1(xk, sj, q, LAM, SIGMA)
        out[f"prox_l0_{suf}"] = orc.prox_l0(xk, sj, q, LAM, SIGMA)
        out[f"prox_lhalf_{suf}"] = orc.prox_lhalf(xk, sj, q, LAM, SIGMA)
        out[f"iprox_l1_{suf}"] = orc.iprox_l1(xk, sj, q, dpos, LAM)
        out[f"iprox_l0_{suf}"] = orc.iprox_l0(xk, sj, q, dpos, LAM)
        for h in ("l1", "l0", "lhalf"):
            out[f"prox_{h}box_{suf}"] = orc.prox_box(h, xk, sj, q, l, ub, LAM, SIGMA)
        for h in ("l1", "l0"):
            out[f"iprox_{h}box_{suf}"] = orc.iprox_box(h, xk, sj, q, d, l, ub, LAM)
        y0 = orc.prox_l1b2(xk, sj, q, LAM, SIGMA, 1e30)
        full = float(np.linalg.norm((y0 + sj).astype(np.float64)))
        out[f"prox_l1b2_{suf}"] = orc.prox_l1b2(xk, sj, q, LAM, SIGMA, 0.5 * full)
        out[f"l1b2_delta_{suf}"] = np.array([0.5 * full])
        out[f"prox_groupl2_{suf}"] = orc.prox_groupl2(xk, sj, q, offs, lam_g, 0.3)
        out[f"prox_groupl2binf_{suf}"] = orc.prox_groupl2binf(xk, sj, q, offs, lam_g, 0.3, DELTA)
        # fused solver step (oracle.solver_step: the composition -ν∇f, prox!, xk+sj+s, ψ(s), ‖s‖, ∇f's), ∇f = q
        for h in ("l1", "l0", "lhalf"):
            for tag, bounds in (("step", ()), ("stepbox", (l, ub))):
                s_, xsy, psi, sn, gd = orc.solver_step(h, xk, sj, q, LAM, NU, *bounds)
                out[f"{tag}_{h}_s_{suf}"] = s_
                out[f"{tag}_{h}_xsy_{suf}"] = xsy
                out[f"{tag}_{h}_scalars_{suf}"] = np.array([psi, sn, gd])
        out[f"prox_indballl0_{suf}"] = orc.prox_indballl0(xk, sj, q, 31)
        out[f"prox_indballl0binf_{suf}"] = orc.prox_indballl0(xk, sj, q, 31, delta=1.0)
    return out


if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "oracle_r1.npz"), **build())
    print("wrote tests/golden/oracle_r1.npz")

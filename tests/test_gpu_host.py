"""Host-buffer entry points (spx_box_host_*, spx_box_multi_host_*): chunked H2D -> kernel -> D2H must give
the oracle's results for any chunking (several chunks, ragged last chunk), with scalar or vector bounds."""
import numpy as np
import pytest

from gpu_util import DEV, bounds, check_lhalfbox, diag, inputs, orc, sp
from shiftedprox import hostpath as hp

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("n,chunk", [(1, 4), (1003, 256), (65_539, 20_000), (5000, 1 << 22)])
@pytest.mark.parametrize("vecb", [True, False])
def test_box_host_single_and_multi(dt, n, chunk, vecb):
    xk, sj, q = inputs(n, dt)
    d = diag(n, dt)
    l, u = bounds(n, dt) if vecb else (-1.0, 1.0)
    lo, uo = (l, u) if vecb else (dt(l), dt(u))
    lam, sigma = 0.8, 0.1
    ctx = sp.context(DEV)
    ref_p = orc.prox_box("l0", xk, sj, q, lo, uo, lam, sigma)
    ref_p1 = orc.prox_box("l1", xk, sj, q, lo, uo, lam, sigma)
    ref_i = orc.iprox_box("l0", xk, sj, q, d, lo, uo, lam)
    # one operation per call
    y = np.empty(n, dt)
    v = hp.box_host(ctx, "l0", y, xk, sj, q, l, u, lam, sigma, chunk=chunk, want_value=True)
    assert np.array_equal(y, ref_p)
    assert v == pytest.approx(orc.value_box("l0", xk, sj, ref_p, lo, uo, lam), rel=1e-6 if dt == np.float32 else 1e-12)
    hp.box_host(ctx, "l0", y, xk, sj, q, l, u, lam, d=d, chunk=chunk)
    assert np.array_equal(y, ref_i, equal_nan=True)
    # several operations, one pass
    y0, y1, y2, y3 = (np.empty(n, dt) for _ in range(4))
    jobs = [dict(op="l0", y=y0, q=q, lam=lam, sigma=sigma), dict(op="lhalf", y=y1, q=q, lam=lam, sigma=sigma),
            dict(op="l0", y=y2, q=q, d=d, lam=lam), dict(op="l1", y=y3, q=q, lam=lam, sigma=sigma)]
    vals = hp.box_multi_host(ctx, jobs, xk, sj, l, u, chunk=chunk, want_value=True)
    assert np.array_equal(y0, ref_p)
    assert np.array_equal(y2, ref_i, equal_nan=True)
    assert np.array_equal(y3, ref_p1)
    check_lhalfbox(y1, xk, sj, q, lo, uo, lam, sigma, label=f"host {dt.__name__} n={n}")
    rel = 1e-6 if dt == np.float32 else 1e-12
    assert vals[0] == pytest.approx(orc.value_box("l0", xk, sj, ref_p, lo, uo, lam), rel=rel)
    assert vals[3] == pytest.approx(orc.value_box("l1", xk, sj, ref_p1, lo, uo, lam), rel=rel)


def test_box_multi_host_rejects_bad_jobs():
    from shiftedprox import _lib as L
    xk = np.zeros(8)
    y = np.zeros(8)
    with pytest.raises(L.SpxError):
        hp.box_multi_host(sp.context(DEV), [dict(op="lhalf", y=y, q=xk, d=xk, lam=1.0)], xk, xk, -1.0, 1.0)
    with pytest.raises(L.SpxError):
        hp.box_multi_host(sp.context(DEV), [], xk, xk, -1.0, 1.0)

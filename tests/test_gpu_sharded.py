"""GPU checks of the sharded entry points on ONE device (two shards processed one after the other; the
collective is emulated by adding the partials).  The real 2..8-rank NCCL path is exercised by
tools/check_sharded_nccl.py under torchrun (gpurun --gpus N)."""
import ctypes as C

import numpy as np
import pytest
import torch

from gpu_util import DEV, N, T, bounds, inputs, orc, sp
from shiftedprox import _lib as L
from shiftedprox import sharded

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_value_partials_add_up(dt):
    n = 100_003
    xk, sj, _ = inputs(n, dt)
    l, u = bounds(n, dt)
    w = (l + (u - l) * orc.uniform(n, 11, dt)).astype(dt)
    y = (w - sj).astype(dt)
    rtol = 1e-13 if dt == np.float64 else 2e-6
    for kind, h in (("l1", sp.NormL1(1.7)), ("l0", sp.NormL0(1.7)), ("lhalf", sp.RootNormLhalf(1.7))):
        tot, bad = 0.0, False
        for r in range(2):
            lo, hi = sharded.shard_bounds(n, 2, r)
            psi = sp.shifted(sp.shifted(h, T(xk[lo:hi]), T(l[lo:hi]), T(u[lo:hi])), T(sj[lo:hi]))
            out = (C.c_double * 2)()
            lb, ub = psi._bounds()
            yl = T(y[lo:hi])
            psi._call("value_partial", C.c_int32(psi._H_KIND), C.c_int64(psi.n), C.c_void_p(psi.xk.data_ptr()),
                      C.c_void_p(psi.sj.data_ptr()), C.c_void_p(yl.data_ptr()), C.byref(lb), C.byref(ub),
                      psi._sel.ref(), C.c_int32(1), out)
            tot += out[0]
            bad = bad or out[1] > 0
        val = sharded.combine_value(kind, tot, 1.0 if bad else 0.0, 1.7, torch.float64 if dt == np.float64 else torch.float32)
        assert val == pytest.approx(orc.value_box(kind, xk, sj, y, l, u, 1.7), rel=rtol)
        # un-initialised process group: value_sharded degenerates to the local value
        psi = sp.shifted(sp.shifted(h, T(xk), T(l), T(u)), T(sj))
        assert sharded.value_sharded(psi, T(y)) == pytest.approx(psi(T(y)), rel=rtol)


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_l1b2_sharded_callback_matches_single_device(dt):
    n = 200_001
    xk, sj, q = inputs(n, dt)
    y0 = orc.prox_l1b2(xk, sj, q, 1.0, 0.1, 1e30)
    delta = 0.5 * float(np.linalg.norm((y0 + sj).astype(np.float64)))
    psi = sp.shifted(sp.shifted(sp.NormL1(1.0), T(xk), delta, sp.NormL2(1.0)), T(sj))
    y1 = torch.empty(n, dtype=T(q).dtype, device=DEV)
    sp.prox_(y1, psi, T(q), 0.1)
    calls = []

    def ident(_u, vals, count):  # world size 1: the all-reduce is the identity
        calls.append(count)
        return 0

    cb = L.ALLREDUCE_FN(ident)
    y2 = torch.empty_like(y1)
    passes = C.c_int32()
    val = C.c_double()
    tq = T(q)
    psi._call("prox_l1b2_sharded", C.c_int64(n), C.c_void_p(y2.data_ptr()), C.c_void_p(psi.xk.data_ptr()),
              C.c_void_p(psi.sj.data_ptr()), C.c_void_p(tq.data_ptr()), C.c_double(1.0), C.c_double(0.1),
              C.c_double(delta), C.c_double(1.0), cb, None, C.byref(passes), C.byref(val))
    assert np.array_equal(N(y1), N(y2))
    assert len(calls) == passes.value  # one all-reduce per norm pass + one for the ψ sums
    assert val.value == pytest.approx(psi(y2), rel=1e-12 if dt == np.float64 else 1e-5)


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("n,r,quant", [(300_001, 1234, False), (200_000, 50_000, True), (70_000, 69_999, False)])
@pytest.mark.parametrize("binf", [False, True])
def test_topr_single_vector_over_two_shards(dt, n, r, quant, binf):
    """One vector cut into two shards (two contexts with their own streams, two host threads standing in for two
    ranks; the all-reduce callback meets at a barrier): histogram exchange per radix digit + cross-shard tie
    rule give exactly the single-device / oracle result, ties included (magnitudes quantised to 1/64)."""
    import ctypes as C
    import threading

    from shiftedprox import _lib as L

    xk, sj, q = inputs(n, dt)
    if quant:
        xk, sj, q = (np.round(a * 64) / 64 for a in (xk, sj, q))
        xk, sj, q = xk.astype(dt), sj.astype(dt), q.astype(dt)
    delta = 1.0
    ref = orc.prox_indballl0(xk, sj, q, r, delta=delta if binf else None)
    world = 2
    cut = (n // 2 + 3) & ~3
    bounds_ = [(0, cut), (cut, n)]
    suf = "f64" if dt == np.float64 else "f32"
    barrier = threading.Barrier(world)
    slots = [None] * world
    outs = [None] * world
    errs = []

    def worker(rank):
        try:
            lo, hi = bounds_[rank]
            ctx = C.c_void_p()
            L.call("spx_ctx_create", C.byref(ctx), C.c_int32(0), C.c_void_p(0), C.c_int32(1))
            txk, tsj, tq = T(xk[lo:hi]), T(sj[lo:hi]), T(q[lo:hi])
            y = torch.empty_like(tq)
            torch.cuda.synchronize()

            def reduce(_user, vals, count):
                slots[rank] = [vals[i] for i in range(count)]
                barrier.wait()
                tot = [sum(s[i] for s in slots) for i in range(count)]
                barrier.wait()
                for i in range(count):
                    vals[i] = tot[i]
                return 0

            cb = L.ALLREDUCE_FN(reduce)
            L.call(f"spx_prox_indballl0_sharded_{suf}", ctx, C.c_int64(hi - lo), C.c_int64(n), C.c_void_p(y.data_ptr()),
                   C.c_void_p(txk.data_ptr()), C.c_void_p(tsj.data_ptr()), C.c_void_p(tq.data_ptr()), C.c_int64(r),
                   C.c_int32(1 if binf else 0), C.c_double(delta), C.c_int32(rank), C.c_int32(world), cb, None)
            L.call("spx_ctx_synchronize", ctx)
            outs[rank] = N(y)
            L.call("spx_ctx_destroy", ctx)
        except Exception as e:  # pragma: no cover
            errs.append(e)
            barrier.abort()

    th = [threading.Thread(target=worker, args=(k,)) for k in range(world)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    assert np.array_equal(np.concatenate(outs), ref)


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_step_shards_add_up(dt):
    # fused solver step on two contiguous shards: s is the concatenation, the three scalars add up
    n = 100_003
    xk, sj, grad = inputs(n, dt)
    l, u = bounds(n, dt)
    nu = 0.15
    for h, boxed in ((sp.NormL1(1.2), False), (sp.NormL0(1.2), True), (sp.RootNormLhalf(1.2), True)):
        mk = (lambda a, b: sp.shifted(sp.shifted(h, T(xk[a:b]), T(l[a:b]), T(u[a:b])), T(sj[a:b]))) if boxed else \
             (lambda a, b: sp.shifted(sp.shifted(h, T(xk[a:b])), T(sj[a:b])))
        whole = mk(0, n)
        s = torch.empty(n, dtype=T(grad).dtype, device=DEV)
        _, ref = sp.step_(s, whole, T(grad), nu)
        psi_sum, ss, gd, parts = 0.0, 0.0, 0.0, []
        for r in range(2):
            lo, hi = sharded.shard_bounds(n, 2, r)
            sl = torch.empty(hi - lo, dtype=s.dtype, device=DEV)
            _, res = sp.step_(sl, mk(lo, hi), T(grad[lo:hi]), nu)
            parts.append(sl)
            psi_sum += res.psi; ss += res.snorm ** 2; gd += res.gdots
        assert torch.equal(torch.cat(parts), s)
        got = sharded.allreduce_step(psi_sum, ss, gd)  # un-initialised process group: the identity
        rel = 1e-12 if dt == np.float64 else 1e-5
        assert got[0] == pytest.approx(ref.psi, rel=rel) and got[1] == pytest.approx(ref.snorm, rel=1e-12)
        assert abs(got[2] - ref.gdots) <= 1e-12 * float(np.sum(np.abs(grad.astype(np.float64) * N(s).astype(np.float64))))
        # step_sharded_ without a process group degenerates to the local step
        s2 = torch.empty_like(s)
        _, r2 = sharded.step_sharded_(s2, whole, T(grad), nu)
        assert torch.equal(s2, s) and r2.psi == ref.psi and r2.snorm == pytest.approx(ref.snorm, rel=1e-15)

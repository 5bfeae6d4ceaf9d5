"""world_size-2 `gloo` tests (CPU) of the N>1 host logic: shard partitioning and the ψ(y) scalar all-reduce.
The per-shard numbers come from the oracle here (no GPU in this container); on the GPU box the same
`allreduce_value` is fed by spx_value_partial_* (tests/test_gpu_sharded.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle as orc
from shiftedprox import sharded


def test_shard_bounds_cover_and_align():
    for n in (0, 1, 7, 1000, 2 ** 20 + 3):
        for world in (1, 2, 3, 8):
            spans = [sharded.shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and a <= b
            for a, b in spans[:-1]:
                assert b % 4 == 0 or b == n


def test_shard_groups_split_on_boundaries():
    rng = np.random.default_rng(0)
    sizes = rng.integers(1, 4097, size=500)
    offs = np.concatenate([[0], np.cumsum(sizes)])
    for world in (1, 2, 4, 8):
        runs = sharded.shard_groups(offs, world)
        assert runs[0][0] == 0 and runs[-1][1] == 500
        for (a, b), (c, d) in zip(runs, runs[1:]):
            assert b == c
        elems = [offs[b] - offs[a] for a, b in runs]
        assert max(elems) <= offs[-1] / world + 4096  # balanced to within one group


def test_shard_problems():
    assert [sharded.shard_problems(4096, 8, r) for r in (0, 7)] == [(0, 512), (3584, 4096)]
    assert sharded.shard_problems(5, 4, 3) == (5, 5)


def test_combine_value_semantics():
    assert sharded.combine_value("l1", 3.0, 0.0, 2.0, torch.float64) == 6.0
    assert sharded.combine_value("l1", 3.0, 1.0, 2.0, torch.float64) == np.inf
    assert sharded.combine_value("indballl0", 5.0, 0.0, 0.0, torch.float64, r=5) == 0.0
    assert sharded.combine_value("indballl0", 6.0, 0.0, 0.0, torch.float64, r=5) == np.inf
    # λ·Σ is rounded in R like the reference's value functor
    assert sharded.combine_value("l1", 1.0 / 3.0, 0.0, 3.0, torch.float32) == float(np.float32(3) * np.float32(1 / 3))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        xk = orc.uniform(n, 0, np.float64, 4.0, -2.0)
        sj = orc.uniform(n, 1, np.float64, 1.0, -0.5)
        l = -(0.25 + orc.uniform(n, 3))
        u = 0.25 + orc.uniform(n, 4)
        w = l + (u - l) * orc.uniform(n, 11)
        y = w - sj
        lo, hi = sharded.shard_bounds(n, world, rank)
        sl = slice(lo, hi)
        res = {}
        for kind in ("l1", "l0", "lhalf"):
            # per-shard partial = un-scaled Σ (λ = 1) and the infeasibility flag
            part = orc.value_box(kind, xk[sl], sj[sl], y[sl], l[sl], u[sl], 1.0)
            res[kind] = sharded.allreduce_value(kind, part, False, 1.7, torch.float64)
            ybad = y.copy()
            ybad[n - 1] += 10.0  # infeasible on the last rank only
            pb = orc.value_box(kind, xk[sl], sj[sl], ybad[sl], l[sl], u[sl], 1.0)
            res[kind + "_bad"] = sharded.allreduce_value(kind, 0.0 if np.isinf(pb) else pb, bool(np.isinf(pb)), 1.7,
                                                         torch.float64)
        # top-r feasibility count
        cnt = float(np.count_nonzero((xk[sl] + sj[sl]) + y[sl]))
        res["indball"] = sharded.allreduce_value("indballl0", cnt, False, 0.0, torch.float64, r=n)
        res["indball_small_r"] = sharded.allreduce_value("indballl0", cnt, False, 0.0, torch.float64, r=n // 2)
        if rank == 0:
            full = {k: orc.value_box(k, xk, sj, y, l, u, 1.7) for k in ("l1", "l0", "lhalf")}
            out.put((res, full))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_value_allreduce_gloo_world2():
    world, n = 2, 10_007
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, out)) for r in range(world)]
    for p in procs:
        p.start()
    res, full = out.get(timeout=100)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for kind in ("l1", "l0", "lhalf"):
        assert res[kind] == pytest.approx(full[kind], rel=1e-13)
        assert res[kind + "_bad"] == np.inf
    assert res["indball"] == 0.0
    assert res["indball_small_r"] == np.inf


def _step_worker(rank, world, port, n, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        xk = orc.uniform(n, 0, np.float64, 4.0, -2.0)
        sj = orc.uniform(n, 1, np.float64, 1.0, -0.5)
        grad = orc.uniform(n, 2, np.float64, 4.0, -2.0)
        l = -(0.25 + orc.uniform(n, 3))
        u = 0.25 + orc.uniform(n, 4)
        lo, hi = sharded.shard_bounds(n, world, rank)
        sl = slice(lo, hi)
        res = {}
        for op in ("l1", "l0", "lhalf"):
            # the per-shard scalars come from the oracle here; on the GPU box from spx_step_* (step_sharded_)
            _, _, psi, sn, gd = orc.solver_step(op, xk[sl], sj[sl], grad[sl], 1.3, 0.2)
            res[op] = sharded.allreduce_step(psi, sn * sn, gd)
            _, _, psi, sn, gd = orc.solver_step(op, xk[sl], sj[sl], grad[sl], 1.3, 0.2, l[sl], u[sl])
            res[op + "_box"] = sharded.allreduce_step(psi, sn * sn, gd)
        # an infeasible shard (inverted box on the last rank only) makes the whole value Inf
        lbad = l.copy()
        lbad[n - 1] = u[n - 1] + 1.0
        _, _, psi, sn, gd = orc.solver_step("lhalf", xk[sl], sj[sl], grad[sl], 1.3, 0.2, lbad[sl], u[sl])
        res["bad"] = sharded.allreduce_step(psi, sn * sn, gd)
        if rank == 0:
            full = {}
            for op in ("l1", "l0", "lhalf"):
                full[op] = orc.solver_step(op, xk, sj, grad, 1.3, 0.2)[2:]
                full[op + "_box"] = orc.solver_step(op, xk, sj, grad, 1.3, 0.2, l, u)[2:]
            out.put((res, full))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_step_allreduce_gloo_world2():
    world, n = 2, 10_007
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_step_worker, args=(r, world, port, n, out)) for r in range(world)]
    for p in procs:
        p.start()
    res, full = out.get(timeout=100)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for k, ref in full.items():
        assert res[k] == pytest.approx(ref, rel=1e-12)
    assert res["bad"][0] == np.inf and np.isfinite(res["bad"][1])

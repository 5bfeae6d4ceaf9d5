#!/usr/bin/env python
"""Multi-rank (NCCL) check of the sharded paths; run under torchrun on N GPUs of one box:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29533 tools/check_sharded_nccl.py

Every rank holds a contiguous shard of the same synthetic vectors.  Checks that (1) the all-reduced ψ(y)
equals the single-device value, (2) the sharded ShiftedNormL1B2 prox! (K partial sums all-reduced per pass,
root search replicated) reproduces the single-device result bit for bit on every shard, (3) the top-r projection
of one vector spread over the ranks (histogram all-reduce per radix digit) is identical to the single-device one,
(4) the same scalars through the collectives INSIDE libshiftedprox (spx_comm_init + spx_comm_reduce_scalars:
ncclAllReduce on the device slots): ψ(y), fused ψ, an infeasible shard -> Inf on every rank, the IndBallL0 count,
the fused solver step's three scalars, ShiftedNormL1B2 prox! with its per-pass sums, a single ℓ2 group spanning the
sharded vector."""
import ctypes as C
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "shiftedproximaloperators.jl_b200"))
import shiftedprox as sp  # noqa: E402
from shiftedprox import _lib as L, sharded  # noqa: E402

SEED = 20261018


def uniform(dev, n, stream, scale=1.0, shift=0.0, i0=0, dtype=torch.float64):
    t = torch.empty(n, dtype=dtype, device=dev)
    suf, ct = ("f64", C.c_double) if dtype == torch.float64 else ("f32", C.c_float)
    L.call(f"spx_fill_uniform_{suf}", sp.context(dev), C.c_void_p(t.data_ptr()), C.c_int64(n), C.c_int64(i0),
           C.c_uint64(SEED), C.c_uint64(stream), ct(scale), ct(shift))
    return t


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n = 1 << 24
    ok = True
    for dtype in (torch.float64, torch.float32):
        full = [uniform(dev, n, 0, 4.0, -2.0, dtype=dtype), uniform(dev, n, 1, 1.0, -0.5, dtype=dtype),
                uniform(dev, n, 2, 4.0, -2.0, dtype=dtype)]
        lo, hi = sharded.shard_bounds(n, world, rank)
        xk, sj, q = (t[lo:hi].clone() for t in full)
        # (1) ψ(y) all-reduce
        psi_f = sp.shifted(sp.shifted(sp.NormL1(1.3), full[0]), full[1])
        psi_s = sp.shifted(sp.shifted(sp.NormL1(1.3), xk), sj)
        yf = torch.empty_like(full[2])
        sp.prox_(yf, psi_f, full[2], 0.1)
        v_full = psi_f(yf)
        v_sh = sharded.value_sharded(psi_s, yf[lo:hi].clone())
        rel = abs(v_full - v_sh) / abs(v_full)
        ok &= rel < (1e-13 if dtype == torch.float64 else 1e-6)
        # (2) L1B2 sharded
        b0 = sp.shifted(sp.shifted(sp.NormL1(1.0), full[0], 1e30, sp.NormL2(1.0)), full[1])
        sp.prox_(yf, b0, full[2], 0.1)
        delta = 0.5 * float(torch.linalg.vector_norm((yf + full[1]).double()))
        bf = sp.shifted(sp.shifted(sp.NormL1(1.0), full[0], delta, sp.NormL2(1.0)), full[1])
        _, vf = sp.prox_(yf, bf, full[2], 0.1, want_value=True)
        bs = sp.shifted(sp.shifted(sp.NormL1(1.0), xk, delta, sp.NormL2(1.0)), sj)
        ys = torch.empty_like(q)
        _, vs = sharded.prox_l1b2_sharded_(ys, bs, q, 0.1, want_value=True)
        # the K sums differ from the single-device fold only in summation order: η may move by an ulp
        tol = (64 if dtype == torch.float64 else 16) * torch.finfo(dtype).eps
        diff = float((ys - yf[lo:hi]).abs().max())
        ok &= diff <= tol * 4.0
        ok &= abs(vs - vf) <= 1e-9 * abs(vf) if dtype == torch.float64 else abs(vs - vf) <= 1e-4 * abs(vf)
        # (3) one vector's top-r across the ranks: histogram all-reduce per radix digit, cross-shard tie rule
        r_top = 100_003
        qq = (full[2] * 64).round() / 64  # quantised: many ties at the threshold
        pf = sp.shifted(sp.shifted(sp.IndBallL0(r_top), full[0], 1.0, sp.NormLinf(1.0)), full[1])
        sp.prox_(yf, pf, qq, 1.0)
        ps = sp.shifted(sp.shifted(sp.IndBallL0(r_top), xk, 1.0, sp.NormLinf(1.0)), sj)
        yt = torch.empty_like(q)
        sharded.prox_indballl0_sharded_(yt, ps, qq[lo:hi].clone(), n)
        top_ok = bool(torch.equal(yt, yf[lo:hi]))
        ok &= top_ok
        if rank == 0:
            print(f"[{dtype}] world={world} sharded top-r (r={r_top}, quantised ties) identical to single device: {top_ok}",
                  flush=True)
        # (4) collectives inside the library
        rtol = 1e-13 if dtype == torch.float64 else 1e-6
        if dtype == torch.float64 and not getattr(main, "_comm", False):
            sharded.comm_init(dev)
            main._comm = True
        sharded.reduce_scalars(True, dev)
        c0 = sharded.comm_info(dev)[2]
        ysh = torch.empty_like(q)
        # ψ(y) stand-alone and fused
        sp.prox_(yf, psi_f, full[2], 0.1)
        sharded.reduce_scalars(False, dev)
        v_full = psi_f(yf)
        sharded.reduce_scalars(True, dev)
        v_lib = psi_s(yf[lo:hi].clone())
        _, v_fused = sp.prox_(ysh, psi_s, q, 0.1, want_value=True)
        lib_ok = abs(v_lib - v_full) <= rtol * abs(v_full) and abs(v_fused - v_full) <= rtol * abs(v_full)
        lib_ok &= bool(torch.equal(ysh, yf[lo:hi]))
        # Box ψ: infeasible on the LAST rank only -> Inf everywhere
        bx = sp.shifted(sp.shifted(sp.NormL0(1.1), xk, -3.0, 3.0), sj)
        ybad = (yf[lo:hi] * 0).clone()
        fin = bx(ybad)
        if rank == world - 1:
            ybad[-1] = 100.0
        inf = bx(ybad)
        lib_ok &= (fin < float("inf")) and (inf == float("inf"))
        # IndBallL0 count over all shards: r = global count -> 0, r one short -> Inf
        cnt = int(torch.count_nonzero((full[0] + full[1]) + yf).item())
        ib = sp.shifted(sp.shifted(sp.IndBallL0(max(cnt, 1)), xk), sj)
        ib1 = sp.shifted(sp.shifted(sp.IndBallL0(max(cnt - 1, 1)), xk), sj)
        lib_ok &= ib(yf[lo:hi].clone()) == 0.0 and (cnt < 2 or ib1(yf[lo:hi].clone()) == float("inf"))
        # fused solver step: the three scalars over the whole vector
        sharded.reduce_scalars(False, dev)
        once_f = sp.shifted(sp.NormL1(1.3), full[0])
        sfull = torch.empty_like(full[2])
        _, rf = sp.step_(sfull, once_f, full[2], 0.2)
        sharded.reduce_scalars(True, dev)
        once_s = sp.shifted(sp.NormL1(1.3), xk)
        _, rs = sp.step_(ysh, once_s, q, 0.2)
        lib_ok &= all(abs(a - b) <= rtol * max(1.0, abs(b)) for a, b in zip(rs, rf)) and bool(torch.equal(ysh, sfull[lo:hi]))
        # L1B2: the plain prox! entry point, its per-pass sums all-reduced on the device
        bs2 = sp.shifted(sp.shifted(sp.NormL1(1.0), xk, delta, sp.NormL2(1.0)), sj)
        _, vs2 = sp.prox_(ysh, bs2, q, 0.1, want_value=True)
        sharded.reduce_scalars(False, dev)
        sp.prox_(yf, bf, full[2], 0.1)
        diff2 = float((ysh - yf[lo:hi]).abs().max())
        lib_ok &= diff2 <= tol * 4.0 and (abs(vs2 - vf) <= (1e-9 if dtype == torch.float64 else 1e-4) * abs(vf))
        # one ℓ2 group spanning the whole (sharded) vector: ShiftedGroupNormL2 from NormL2
        gf = sp.shifted(sp.shifted(sp.NormL2(0.7), full[0]), full[1])
        sp.prox_(yf, gf, full[2], 0.3)
        sharded.reduce_scalars(True, dev)
        gs = sp.shifted(sp.shifted(sp.NormL2(0.7), xk), sj)
        sp.prox_(ysh, gs, q, 0.3)
        gdiff = float((ysh - yf[lo:hi]).abs().max())
        lib_ok &= gdiff <= 8 * torch.finfo(dtype).eps * 8.0
        # single-vector top-r: histograms and tie counts all-reduced by the library's communicator (no callback)
        sp.prox_(yf, pf, qq, 1.0)
        yt2 = torch.empty_like(q)
        sharded.prox_indballl0_sharded_(yt2, ps, qq[lo:hi].clone(), n)
        lib_ok &= bool(torch.equal(yt2, yf[lo:hi]))
        ncoll = sharded.comm_info(dev)[2] - c0
        sharded.reduce_scalars(False, dev)
        ok &= lib_ok
        if rank == 0:
            print(f"[{dtype}] world={world} in-library NCCL: psi {v_lib:.12g} / fused {v_fused:.12g} vs {v_full:.12g}; "
                  f"l1b2 max diff {diff2:.3e} passes {bs2.last_passes}; one-group max diff {gdiff:.3e}; "
                  f"{ncoll} collectives; ok={lib_ok}", flush=True)
        if rank == 0:
            print(f"[{dtype}] world={world} psi rel err {rel:.2e}; l1b2 passes {bs.last_passes} (single {bf.last_passes}), "
                  f"max |y_sharded - y_single| {diff:.3e}, psi {vs:.12g} vs {vf:.12g}", flush=True)
    if getattr(main, "_comm", False):
        # latency of one all-reduced scalar (ψ of a 4096-element shard, host gets the value back every call), and
        # whether every rank got the same bits
        import time
        small = torch.randn(4096, dtype=torch.float64, device=dev)
        hs = sp.shifted(sp.NormL1(1.0), small)
        sharded.reduce_scalars(True, dev)
        for _ in range(20):
            v = hs(small)
        torch.cuda.synchronize(dev)
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(500):
            v = hs(small)
        torch.cuda.synchronize(dev)
        us = (time.perf_counter() - t0) / 500 * 1e6
        sharded.reduce_scalars(False, dev)
        v_loc = hs(small)
        mine = torch.tensor([v, us], dtype=torch.float64, device=dev)
        every = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(every, mine)
        same = all(float(e[0]) == float(every[0][0]) for e in every)
        ok &= same
        if rank == 0:
            print(f"all-reduced scalar: peer-memory path {sharded.comm_peer_active(dev)}; "
                  f"{max(float(e[1]) for e in every):.1f} us per call (local, no collective: measured below); "
                  f"identical bits on every rank: {same}", flush=True)
        torch.cuda.synchronize(dev)
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(500):
            v_loc = hs(small)
        us0 = (time.perf_counter() - t0) / 500 * 1e6
        if rank == 0:
            print(f"same call without the collective: {us0:.1f} us", flush=True)
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("SHARDED CHECK", "OK" if flag.item() == 1.0 else "FAILED", flush=True)
    dist.barrier()
    if getattr(main, "_comm", False):
        sharded.comm_destroy(dev)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()

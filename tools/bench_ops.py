#!/usr/bin/env python
"""Per-operator device-resident timing of every in-scope prox!/iprox!/ψ(y).

    python tools/bench_ops.py [--log2n 28] [--dtype f64] [--reps 10] [--only REGEX] [--json out.json]

Prints one line per operator: ms, elements/s, algorithmic GB/s, fraction of the measured HBM peak.
CUDA events on the launching stream, 3 warm-ups, operands far larger than L2.
"""
import argparse
import json
import os
import re
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "shiftedproximaloperators.jl_b200"))
import ctypes as C  # noqa: E402

import shiftedprox as sp  # noqa: E402
from shiftedprox import _lib as L  # noqa: E402

SEED = 20261018


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n", type=int, default=28)
    ap.add_argument("--dtype", default="f64")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--only", default=".*")
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    tdt = torch.float64 if args.dtype == "f64" else torch.float32
    R = 8 if args.dtype == "f64" else 4
    n = 1 << args.log2n
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0

    def uni(stream, scale=1.0, shift=0.0, m=n):
        t = torch.empty(m, dtype=tdt, device=dev)
        f = f"spx_fill_uniform_{args.dtype}"
        ct = C.c_double if args.dtype == "f64" else C.c_float
        L.call(f, sp.context(dev), C.c_void_p(t.data_ptr()), C.c_int64(m), C.c_int64(0), C.c_uint64(SEED),
               C.c_uint64(stream), ct(scale), ct(shift))
        return t

    xk, sj, q = uni(0, 4.0, -2.0), uni(1, 1.0, -0.5), uni(2, 4.0, -2.0)
    l = uni(3).add_(0.25).neg_()
    u = uni(4).add_(0.25)
    d = uni(5).add_(0.5)
    b = uni(6)
    dpos = d.clone()
    d = torch.where(b < 0.1, -d, d)
    d = torch.where((b >= 0.1) & (b < 0.2), torch.zeros_like(d), d)
    del b
    y = torch.empty(n, dtype=tdt, device=dev)
    lam, sigma = 1.0, 0.1
    results = []

    def timeit(name, fn, alg_bytes_per_elt, elems=n, note=""):
        if not re.search(args.only, name):
            return
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.reps)]
        for a, b_ in ev:
            a.record()
            fn()
            b_.record()
        torch.cuda.synchronize()
        ts = sorted(a.elapsed_time(b_) for a, b_ in ev)
        ms = ts[len(ts) // 2]
        gbs = alg_bytes_per_elt * elems / (ms * 1e-3) / 1e9
        r = {"op": name, "dtype": args.dtype, "n": elems, "ms_median": ms, "ms_best": ts[0],
             "elements_per_s": elems / (ms * 1e-3), "alg_bytes_per_elt": alg_bytes_per_elt, "GBps": gbs,
             "frac_measured_peak": gbs / peak, "frac_nominal_8000": gbs / 8000.0, "note": note}
        results.append(r)
        print(f"{name:34s} {ms:9.3f} ms  {r['elements_per_s']:.3e} el/s  {gbs:8.1f} GB/s  "
              f"{100 * gbs / peak:5.1f}% of measured {peak:.0f}  {note}", flush=True)

    def two(h, *a, **k):
        return sp.shifted(sp.shifted(h, xk, *a, **k), sj)

    # separable
    for nm, h in (("l1", sp.NormL1(lam)), ("l0", sp.NormL0(lam)), ("lhalf", sp.RootNormLhalf(lam))):
        psi = two(h)
        timeit(f"prox_{nm}", lambda psi=psi: sp.prox_(y, psi, q, sigma), 4 * R)
        timeit(f"prox_{nm}+psi", lambda psi=psi: sp.prox_(y, psi, q, sigma, want_value=True), 4 * R, note="fused ψ(y)")
        timeit(f"value_{nm}", lambda psi=psi: psi(y), 3 * R)
        if nm != "lhalf":
            timeit(f"iprox_{nm}", lambda psi=psi: sp.iprox_(y, psi, q, dpos), 5 * R)
    # Box, vector bounds and scalar bounds
    for nm, h in (("l1box", sp.NormL1(lam)), ("l0box", sp.NormL0(lam)), ("lhalfbox", sp.RootNormLhalf(lam))):
        psi = two(h, l, u)
        timeit(f"prox_{nm}_vec", lambda psi=psi: sp.prox_(y, psi, q, sigma), 6 * R)
        timeit(f"prox_{nm}_vec+psi", lambda psi=psi: sp.prox_(y, psi, q, sigma, want_value=True), 6 * R, note="fused ψ(y)")
        psis = two(h, -1.0, 1.0)
        timeit(f"prox_{nm}_scalar", lambda psi=psis: sp.prox_(y, psi, q, sigma), 4 * R)
        psir = two(h, l, u, range(0, n, 2))
        timeit(f"prox_{nm}_vec_sel1:2:n", lambda psi=psir: sp.prox_(y, psi, q, sigma), 6 * R)
        timeit(f"value_{nm}_vec", lambda psi=psi: psi(y), 5 * R)
        if nm != "lhalfbox":
            timeit(f"iprox_{nm}_vec", lambda psi=psi: sp.iprox_(y, psi, q, d), 7 * R)
            timeit(f"iprox_{nm}_scalar", lambda psi=psis: sp.iprox_(y, psi, q, d), 5 * R)
    # fused solver step: q = -ν∇f, prox!, ψ(s), xk+sj+s, ‖s‖, ∇f's in one pass
    step_names = [f"step_{nm}{sfx}" for nm in ("l1", "l0", "lhalf") for sfx in ("_once", "", "box_vec", "box_scalar")]
    if any(re.search(args.only, nm) for nm in step_names + ["step_l1_once_unfused"]):
        xsy = torch.empty(n, dtype=tdt, device=dev)
        for nm, h in (("l1", sp.NormL1(lam)), ("l0", sp.NormL0(lam)), ("lhalf", sp.RootNormLhalf(lam))):
            once = sp.shifted(h, xk)
            timeit(f"step_{nm}_once", lambda psi=once: sp.step_(y, psi, q, sigma, xsy=xsy), 4 * R,
                   note="ψ shifted once (R2): 2 reads + 2 writes")
            timeit(f"step_{nm}", lambda psi=two(h): sp.step_(y, psi, q, sigma, xsy=xsy), 5 * R)
            timeit(f"step_{nm}box_vec", lambda psi=two(h, l, u): sp.step_(y, psi, q, sigma, xsy=xsy), 7 * R)
            timeit(f"step_{nm}box_scalar", lambda psi=two(h, -1.0, 1.0): sp.step_(y, psi, q, sigma, xsy=xsy), 5 * R)

        def unfused(psi=sp.shifted(sp.NormL1(lam), xk)):  # the same step as separate sweeps (torch for the BLAS-1 parts)
            mq = q * (-sigma)
            sp.prox_(y, psi, mq, sigma)
            v = psi(y)
            torch.add(xk, y, out=xsy)
            return v, float(torch.linalg.vector_norm(y)), float(torch.dot(q, y))
        timeit("step_l1_once_unfused", unfused, 4 * R, note="prox! + ψ(s) + 4 BLAS-1 sweeps, same result")
    # L1B2 (ball active: Δ = half the unconstrained norm)
    if re.search(args.only, "prox_l1b2"):
        psi0 = two(sp.NormL1(lam), 1e30, sp.NormL2(1.0))
        sp.prox_(y, psi0, q, sigma)
        full = float(torch.linalg.vector_norm((y + sj).double()))
        psi = two(sp.NormL1(lam), 0.5 * full, sp.NormL2(1.0))
        sp.prox_(y, psi, q, sigma)
        passes = psi.last_passes
        # bytes actually moved: decision pass 3R, first search pass 3R + 1W (stashes sj + q in y), later search
        # passes 2R, finish 3R + 1W
        nsearch = max(passes - 2, 0)
        timeit("prox_l1b2", lambda: sp.prox_(y, psi, q, sigma), 3 * R + (4 * R if nsearch else 0) + 2 * R * max(nsearch - 1, 0) + 4 * R,
               note=f"{passes - 1} norm passes (3R, 3R+1W, then 2R each) + 1 finish (4R)")
        timeit("prox_l1b2_inactive", lambda: sp.prox_(y, psi0, q, sigma), 7 * R, note="1 norm pass + finish")
    # groups of 64 (and ragged)
    for gname in ("g64", "ragged"):
        names = [f"prox_groupl2_{gname}", f"value_groupl2_{gname}", f"prox_groupl2binf_{gname}",
                 "prox_groupl2binf_g64_biglambda", "step_groupl2_g64"]
        if not any(re.search(args.only, nm) for nm in names):
            continue
        if gname == "g64":
            ng = n // 64
            offs = torch.arange(0, n + 1, 64, dtype=torch.int64, device=dev)
        else:
            rng = np.random.default_rng(3)
            sizes = np.floor(np.exp(rng.uniform(0, np.log(4097), n // 400))).astype(np.int64).clip(1, 4096)
            cs = np.concatenate([[0], np.cumsum(sizes)])
            cs = cs[cs <= n]
            if cs[-1] != n:
                cs = np.concatenate([cs, [n]])
            offs = torch.from_numpy(cs).to(dev)
            ng = offs.numel() - 1
        lam_g = uni(12, 1.0, 0.5, m=ng)
        h = sp.GroupNormL2(lam_g, None, offsets=offs)
        psi = sp.shifted(sp.shifted(h, xk), sj)
        timeit(f"prox_groupl2_{gname}", lambda psi=psi: sp.prox_(y, psi, q, 0.3), 4 * R, note=f"{ng} groups")
        timeit(f"value_groupl2_{gname}", lambda psi=psi: psi(y), 3 * R)
        if gname == "g64" and re.search(args.only, "step_groupl2_g64"):
            xsy_g = torch.empty_like(y)
            timeit("step_groupl2_g64", lambda psi=psi: sp.step_(y, psi, q, 0.3, xsy=xsy_g), 5 * R,
                   note="spx_step_groupl2: one pass on this layout (groups <= 256); alg. bytes = 3 reads + 2 writes")
            del xsy_g
        psib = sp.shifted(sp.shifted(h, xk, 0.5, sp.NormLinf(1.0)), sj)
        timeit(f"prox_groupl2binf_{gname}", lambda psi=psib: sp.prox_(y, psi, q, 0.3), 4 * R, note=f"{ng} groups")
        if gname == "g64":  # σλ_g >> ||sol_g||: the regime of a sparse solution (root next to the pole of c(n) at σλ)
            hb = sp.GroupNormL2(lam_g * 200.0, None, offsets=offs)
            psil = sp.shifted(sp.shifted(hb, xk, 0.5, sp.NormLinf(1.0)), sj)
            timeit("prox_groupl2binf_g64_biglambda", lambda psi=psil: sp.prox_(y, psi, q, 0.3), 4 * R,
                   note="lambda_g x 200")
    # top-r batch: problems of 65536, r = 1024
    if any(re.search(args.only, nm) for nm in ("prox_indballl0_batch", "prox_indballl0binf_batch", "prox_indballl0_single")):
        pn = 65536
        nprob = n // pn
        for binf in (False, True):
            h = sp.IndBallL0(1024)
            psi = (sp.shifted(h, xk, 1.0, sp.NormLinf(1.0), nprob=nprob) if binf else sp.shifted(h, xk, nprob=nprob))
            psi = sp.shifted(psi, sj)
            timeit(f"prox_indballl0{'binf' if binf else ''}_batch", lambda psi=psi: sp.prox_(y, psi, q, 1.0), 4 * R,
                   note=f"{nprob} problems x {pn}, r=1024")
        psi1 = sp.shifted(sp.shifted(sp.IndBallL0(1 << 20), xk), sj)
        timeit("prox_indballl0_single", lambda: sp.prox_(y, psi1, q, 1.0), 4 * R, note="one vector, r=2^20, global path")
    if args.json:
        os.makedirs(os.path.dirname(os.path.abspath(args.json)), exist_ok=True)
        json.dump(results, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""froot evaluations per group of the GroupNormL2Binf root search (needs a -DSPX_GROUP_STATS build:
SPX_LIB=.../var/lib_gstats.so python tools/group_stats.py)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "shiftedproximaloperators.jl_b200"))
import shiftedprox as sp  # noqa: E402
from shiftedprox import _lib as L  # noqa: E402

dev = torch.device("cuda:0")
n = 1 << 22
g = torch.Generator(device="cpu").manual_seed(1)
xk = (torch.rand(n, generator=g, dtype=torch.float64) * 4 - 2).to(dev)
sj = (torch.rand(n, generator=g, dtype=torch.float64) - 0.5).to(dev)
q = (torch.rand(n, generator=g, dtype=torch.float64) * 4 - 2).to(dev)
y = torch.empty_like(q)
for name in ("g64", "ragged"):
    if name == "g64":
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
    lam = (torch.rand(ng, generator=g, dtype=torch.float64) + 0.5).to(dev)
    psi = sp.shifted(sp.shifted(sp.GroupNormL2(lam, None, offsets=offs), xk, 0.5, sp.NormLinf(1.0)), sj)
    out = (C.c_ulonglong * 5)()
    L.lib().spx_debug_group_stats(out, 1)
    sp.prox_(y, psi, q, 0.3)
    torch.cuda.synchronize()
    L.lib().spx_debug_group_stats(out, 1)
    print(f"{name}: groups {out[1]} (of {ng}), froot evaluations (warp-level) {out[0]}, per group-round {out[0] / max(1, out[1]):.2f}")
    print(f"   CTA-per-group kernel: {out[3]} accepted, {out[4]} left to the bracketing search, "
          f"{out[2] / max(1, out[3] + out[4]):.2f} froot evaluations per group (plus the norms pass and the final pass)")

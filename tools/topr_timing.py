#!/usr/bin/env python
"""Where one cluster spends its time per top-r problem (needs a -DSPX_TR_TIMING build of spx_topr.cu:
SPX_LIB=.../var/lib_trtime.so python tools/topr_timing.py).  clock64 deltas of thread 0 of CTA 0."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "shiftedproximaloperators.jl_b200"))
import shiftedprox as sp  # noqa: E402
from shiftedprox import _lib as L  # noqa: E402

dev = torch.device("cuda:0")
nprob, pn, r = 4096, 65536, 1024
n = nprob * pn
g = torch.Generator(device=dev).manual_seed(1)
xk = torch.rand(n, generator=g, dtype=torch.float64, device=dev) * 4 - 2
sj = torch.rand(n, generator=g, dtype=torch.float64, device=dev) - 0.5
q = torch.rand(n, generator=g, dtype=torch.float64, device=dev) * 4 - 2
y = torch.empty_like(q)
ctx = sp.context(dev)
args = (ctx, C.c_int64(nprob), C.c_int64(pn), C.c_void_p(y.data_ptr()), C.c_void_p(xk.data_ptr()),
        C.c_void_p(sj.data_ptr()), C.c_void_p(q.data_ptr()), C.c_int64(r), C.c_int32(0), C.c_double(0.0))
L.call("spx_prox_indballl0_f64", *args)
torch.cuda.synchronize()
out = (C.c_ulonglong * 16)()
L.lib().spx_debug_topr_timing(out, 1)
L.call("spx_prox_indballl0_f64", *args)
torch.cuda.synchronize()
L.lib().spx_debug_topr_timing(out, 1)
names = {0: "radix: load", 1: "radix: local histogram", 2: "radix: cluster sync (hist ready)", 3: "radix: remote sum + sync",
         4: "radix: scan + pick", 5: "radix: keep masks / ties", 6: "radix: write + final sync",
         8: "lin: load", 9: "lin: kmax exchange", 10: "lin: histogram", 11: "lin: sync + remote sum + sync",
         12: "lin: scan + pick", 13: "lin: gather + sync", 14: "lin: ranking + sync", 15: "lin: keep + write + sync"}
tot = sum(out[i] for i in range(16))
for i, nm in names.items():
    if out[i]:
        print(f"{nm:34s} {out[i]:12d} cycles  {100.0 * out[i] / tot:5.1f} %")
print("total cycles of CTA 0:", tot)

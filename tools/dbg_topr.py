import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/shiftedproximaloperators.jl_b200"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
from gpu_util import *
n, r = 4096, 300
dt = np.float64
xk = np.zeros(n, dt); sj = np.zeros(n, dt)
q = (np.round(orc.uniform(n, 2, dt, 4.0, -2.0) * 64) / 64).astype(dt)
psi = sp.shifted(sp.IndBallL0(r), T(xk))
psi.sol.fill_(777.0)
y = N(sp.prox(psi, T(q), 1.0))
torch.cuda.synchronize()
ref = orc.prox_indballl0(xk, sj, q, r)
print("unwritten:", (y == 777.0).sum(), "mismatch:", (y != ref).sum(), "nnz", np.count_nonzero(y), np.count_nonzero(ref))
bad = np.flatnonzero(y != ref)[:10]
print(bad, y[bad], ref[bad], q[bad])

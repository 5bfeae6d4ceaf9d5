#!/usr/bin/env python
"""Error of the stationary-point Newton (tools/proto/lhalf_newton.c) against 200-bit mpmath, by t bin,
next to the error of the reference's own Float64 acos/cos chain (numpy).  CPU only; development aid."""
import ctypes as C
import os
import subprocess
import sys

import mpmath as mp
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
so = "/tmp/lhalf_newton.so"
subprocess.check_call(["gcc", "-O2", "-mfma", "-ffp-contract=off", "-shared", "-fPIC", "-o", so,
                       os.path.join(HERE, "lhalf_newton.c"), "-lm"])
lib = C.CDLL(so)
lib.lhalf_root_fast_v.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_void_p, C.c_void_p]

mp.mp.prec = 200
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
rng = np.random.default_rng(1)
c4 = 0.1 * 1.0 / 4
t = np.concatenate([rng.uniform(1e-4, 1.0, n), 1 - 10 ** rng.uniform(-6, -1, n // 4)])
az = 3.0 * (c4 / t) ** (2.0 / 3.0)
exact = []
tt = []
for a in az:
    A = mp.mpf(float(a))
    T = mp.mpf(c4) * (A / 3) ** mp.mpf(-1.5)
    tt.append(float(T))
    exact.append(mp.mpf(2) / 3 * A * (1 + mp.cos(2 * mp.pi / 3 - 2 * mp.acos(T) / 3)))
tt = np.array(tt)
ok = tt <= 1.0
ex = np.array([float(e) for e in exact])


def ulps(v):
    return np.array([abs(float((mp.mpf(float(x)) - e) / mp.mpf(float(np.spacing(abs(float(e))))))) for x, e in zip(v, exact)])


tnp = c4 * (az / 3) ** (-1.5)
ref = 2.0 / 3.0 * az * (1 + np.cos(2 * np.pi / 3 - 2 * np.arccos(np.minimum(tnp, 1.0)) / 3))
bins = [0, 0.3, 0.7072, 0.8, 0.9, 0.95, 0.98, 0.99, 0.999, 1.0]
print("bins", bins)
rows = {"ref(numpy)": ulps(ref)}
for variant in (0, 1, 4, 5):
    out = np.empty_like(az)
    t32 = np.empty(len(az), np.float32)
    lib.lhalf_root_fast_v(az.ctypes.data, c4, variant, len(az), out.ctypes.data, t32.ctypes.data)
    rows[f"newton v{variant}"] = ulps(out)
for k, u in rows.items():
    line = []
    for lo, hi in zip(bins[:-1], bins[1:]):
        m = ok & (tt > lo) & (tt <= hi)
        line.append(f"{u[m].max():8.2f}" if m.any() else "     n/a")
    print(f"{k:14s} max ulp per bin: " + " ".join(line))

#!/usr/bin/env python
"""Per-SASS-instruction listing of `ncu --page source --csv --print-source sass`: address, warp-instructions executed,
share of the kernel's total, share of the stall samples, instruction text (for finding which part of a long kernel the instructions go to)."""
import csv, sys
rows = csv.reader(open(sys.argv[1]))
hdr = None; tot = 0; agg = []
for r in rows:
    if hdr is None:
        if "Source" in r and "Instructions Executed" in r:
            hdr = r; idx = {h: i for i, h in enumerate(hdr)}
        continue
    try:
        n = int(r[idx["Instructions Executed"]])
    except Exception:
        continue
    try:
        smp = int(r[idx["# Samples"]])
    except Exception:
        smp = 0
    agg.append((r[idx["Address"]] if "Address" in idx else "", n, smp, r[idx["Source"]]))
    tot += n
stot = sum(a[2] for a in agg)
print("total", tot, "samples", stot)
for a, n, smp, sx in agg:
    print(a, n, "%.2f%%" % (100 * n / max(tot, 1)), "smp %.2f%%" % (100 * smp / max(stot, 1)), sx)

#!/usr/bin/env python
"""Per-SASS-instruction listing of `ncu --page source --csv --print-source sass`: address, warp-instructions executed,
share of the kernel's total, instruction text (for finding which part of a long kernel the instructions go to)."""
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
    agg.append((r[idx["Address"]] if "Address" in idx else "", n, r[idx["Source"]]))
    tot += n
print("total", tot)
for a, n, sx in agg:
    print(a, n, "%.2f%%" % (100 * n / max(tot, 1)), sx)

#!/usr/bin/env python
"""Print the SASS listing of an `ncu --page source --csv --print-source sass` export with, per instruction,
executed warp-instructions (in units of the hottest instruction), average active threads and stall samples.
    python tools/ncu_sass.py file_src.csv [min_fraction]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
minf = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
hdr = next(r for r in rows if r and r[0] == "Address")
idx = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows if r and r[0].startswith("0x")]
mx = max(int(r[idx["Instructions Executed"]]) for r in body)
tot = sum(int(r[idx["Instructions Executed"]]) for r in body)
print(f"# hottest instruction executed {mx} times; total {tot} = {tot/mx:.1f} x hottest")
for n, r in enumerate(body):
    ex = int(r[idx["Instructions Executed"]])
    if ex < minf * mx: continue
    print(f"{n:5d} {ex/mx:6.3f} thr={float(r[idx['Avg. Threads Executed']]):4.1f} smp={r[idx['# Samples']]:>6s}  {r[idx['Source']].strip()}")

#!/usr/bin/env python
"""Aggregate `ncu --page source --csv --print-source sass` by source line and by SASS opcode."""
import collections, csv, re, sys
rows = csv.reader(open(sys.argv[1]))
hdr = None
by_line = collections.Counter(); by_op = collections.Counter(); samples = collections.Counter(); src = {}
tot = 0
for row in rows:
    if not row: continue
    if row[0] in ("File Path", "Function Name"): continue
    if row[0] == "Address" or row[0] == "#" or "Source" in row[:3] and hdr is None:
        hdr = row; idx = {h: i for i, h in enumerate(hdr)}; continue
    if hdr is None: continue
    try:
        inst = int(row[idx["Instructions Executed"]])
    except Exception:
        continue
    sass = row[idx["Source"]]
    m = re.match(r'\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)?)', sass)
    op = m.group(1) if m else "?"
    by_op[op.split('.')[0] if not op.startswith("MUFU") else op] += inst
    tot += inst
    try: samples[op.split('.')[0]] += int(row[idx["# Samples"]])
    except Exception: pass
print("total warp-instructions", tot)
for op, c in by_op.most_common(40):
    print(f"{100*c/tot:5.1f}%  {op}")

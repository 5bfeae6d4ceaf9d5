#!/usr/bin/env python
"""Pipe utilisation / issue / stall percentages of one `ncu --page raw --csv` export."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h = rows[0]; v = rows[-1]
for a, b in zip(h, v):
    if any(k in a for k in ("inst_executed_pipe", "pipe_", "issue_active")) and "pct" in a:
        print(a, b)

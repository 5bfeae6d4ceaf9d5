#!/bin/bash
# usage: tools/launch_times.sh <bench_ops --only regex> [extra bench_ops args]: per-kernel mean times (ncu launch list,
# serialised and cold-cache) of the library's kernels behind one bench_ops line
ONLY="$1"; shift
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/_lt.csv \
    python tools/bench_ops.py --only "$ONLY" --reps 2 "$@" > /dev/null 2>&1
python - <<PY
import csv,collections
rows=list(csv.reader(open("gpurun_out/_lt.csv")))
hdr=None; agg=collections.defaultdict(list)
for r in rows:
    if len(r)>5 and r[0]=="ID": hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r))
        if d.get("Metric Name")=="gpu__time_duration.sum" and "spx" in d["Kernel Name"]:
            v=float(d["Metric Value"].replace(",","")); u=d["Metric Unit"]
            v = v/1e6 if u=="ns" else v/1e3 if u=="us" else v
            agg[d["Kernel Name"][:70]+" grid"+d["Grid Size"]].append(v)
for k,v in agg.items(): print(len(v), round(sum(v)/len(v),3), "ms", k)
PY
rm -f gpurun_out/_lt.csv

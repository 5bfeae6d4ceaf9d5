#!/bin/bash
# usage: tools/ncu_one.sh <kernel-regex> <bench_ops --only regex> <out-stem> [extra bench_ops args]
# One `ncu --set full` capture (1 launch, after warm-up launches are skipped) of a kernel driven by tools/bench_ops.py,
# then the raw / source pages exported as CSV next to the .ncu-rep under gpurun_out/.
set -e
K="$1"; ONLY="$2"; OUT="gpurun_out/$3"; shift 3
python tools/bench_ops.py --only "$ONLY" --reps 1 "$@" > "$OUT.plain.log" 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$K" -s 3 -c 1 -f -o "$OUT" \
    python tools/bench_ops.py --only "$ONLY" --reps 1 "$@" > "$OUT.ncu.log" 2>&1
ncu -i "$OUT.ncu-rep" --page raw --csv > "${OUT}_raw.csv" 2>/dev/null
ncu -i "$OUT.ncu-rep" --page source --csv --print-source sass > "${OUT}_src.csv" 2>/dev/null

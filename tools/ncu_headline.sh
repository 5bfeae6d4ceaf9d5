#!/bin/bash
# The headline step's part of tools/ncu_round2.sh (launch list + one `ncu --set full` capture of each of its three
# kernels) and the per-operator tables: what has to be redone when a source file outside those kernels changed
# (profiles/traffic.json is keyed on the hash of all CUDA sources).
set -e
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-configs"
$B > gpurun_out/r02_plain_bench.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -c 400 --csv \
    --log-file gpurun_out/r02_bench_launches.csv $B > gpurun_out/r02_ncu_launches.log 2>&1
for K in ProxL0Box ProxLhalfBox IproxL0Box; do
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$K" -s 3 -c 1 -f \
      -o gpurun_out/r02_bench_$K $B > gpurun_out/r02_ncu_$K.log 2>&1
  ncu -i gpurun_out/r02_bench_$K.ncu-rep --page raw --csv > gpurun_out/r02_bench_${K}_raw.csv 2>/dev/null
  python tools/ncu_summary.py gpurun_out/r02_bench_${K}_raw.csv > gpurun_out/r02_bench_$K.ncu_full_summary.txt
done
rm -f gpurun_out/*.ncu-rep gpurun_out/*_src.csv
python tools/bench_ops.py --json gpurun_out/r02_ops_f64_n2p28.json > gpurun_out/r02_ops_f64_n2p28.txt 2>&1
python tools/bench_ops.py --dtype f32 --log2n 29 --json gpurun_out/r02_ops_f32_n2p29.json > gpurun_out/r02_ops_f32_n2p29.txt 2>&1

#!/bin/bash
# Round-2 profiler evidence, one gpurun call: the launch list of the headline step and one `ncu --set full` capture of
# each kernel the round is judged on.  Outputs under gpurun_out/ (summaries are copied to profiles/ by hand).
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
tools/ncu_one.sh "uniform_kernel<double, .int.8, .bool.1>" "prox_groupl2binf_g64$" r02_group_l2binf_uniform_g64 --log2n 26
python tools/ncu_summary.py gpurun_out/r02_group_l2binf_uniform_g64_raw.csv > gpurun_out/r02_group_l2binf_uniform_g64_n2p26.ncu_full_summary.txt
tools/ncu_one.sh "topr_stream_kernel<double, .bool.1" "prox_indballl0binf_batch" r02_topr_stream_batch --log2n 27
python tools/ncu_summary.py gpurun_out/r02_topr_stream_batch_raw.csv > gpurun_out/r02_topr_stream_batch_n2p27.ncu_full_summary.txt
tools/ncu_one.sh "group_l2binf_big_kernel<double, .int.256" "prox_groupl2binf_ragged" r02_group_l2binf_big_ragged --log2n 26
python tools/ncu_summary.py gpurun_out/r02_group_l2binf_big_ragged_raw.csv > gpurun_out/r02_group_l2binf_big_ragged_n2p26.ncu_full_summary.txt
tools/ncu_one.sh "group_l2_big_kernel<double, .bool.0, .bool.1, .int.256" "^prox_groupl2_ragged" r02_group_l2_big_ragged --log2n 26
python tools/ncu_summary.py gpurun_out/r02_group_l2_big_ragged_raw.csv > gpurun_out/r02_group_l2_big_ragged_n2p26.ncu_full_summary.txt
# keep the call's output under gpurun's 64 MiB limit: the reports and per-instruction pages stay on the box
rm -f gpurun_out/*.ncu-rep gpurun_out/*_src.csv
python tools/bench_ops.py --json gpurun_out/r02_ops_f64_n2p28.json > gpurun_out/r02_ops_f64_n2p28.txt 2>&1
python tools/bench_ops.py --dtype f32 --log2n 29 --json gpurun_out/r02_ops_f32_n2p29.json > gpurun_out/r02_ops_f32_n2p29.txt 2>&1
tools/launch_times.sh "^prox_groupl2_ragged" > gpurun_out/r02_ragged_launch_times.txt 2>&1
tools/launch_times.sh "^prox_groupl2binf_ragged" >> gpurun_out/r02_ragged_launch_times.txt 2>&1

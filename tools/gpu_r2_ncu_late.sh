#!/bin/bash
# Round-2 late evidence after the issuer / pipeline rewrites: launch list of the bench step and ncu --set full of the
# rewritten kernels (exported to CSV on the box; the report is too big to travel).
set -u
mkdir -p gpurun_out
BENCH_CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --chamfer-steps 1 --train-steps 2 --sampling-steps 2 --batched-scans 0"
timeout 300 $BENCH_CMD > gpurun_out/plain_bench.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv \
    --log-file gpurun_out/launches_r02_late.csv $BENCH_CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?"; tail -c 300 gpurun_out/ncu_launches.log
timeout 300 python tools/ncu_once_r2.py > gpurun_out/plain_once.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on \
    -k regex:'sa_mlp_tc_kernel|noise_mlp_kernel|train_gemm_kernel|train_wgrad_kernel|vox_runs_kernel|vox_scatter_kernel' \
    -s 0 -c 60 -o gpurun_out/prof_r02_late python tools/ncu_once_r2.py > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"; tail -3 gpurun_out/ncu_full.log
ncu -i gpurun_out/prof_r02_late.ncu-rep --page raw --csv > gpurun_out/prof_r02_late_raw.csv 2> /dev/null
for k in noise_mlp_kernel sa_mlp_tc_kernel; do
  ncu -i gpurun_out/prof_r02_late.ncu-rep --page source --csv -k regex:$k > gpurun_out/prof_r02_late_source_$k.csv 2> /dev/null
done
rm -f gpurun_out/prof_r02_late.ncu-rep
ls -la gpurun_out | tail -8

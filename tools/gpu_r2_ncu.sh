#!/bin/bash
# Round-2 evidence: launch list of the bench step (gpu__time_duration) and ncu --set full of the round's kernels.
set -u
mkdir -p gpurun_out
if [ "${NCU_LAUNCHES:-1}" = "1" ]; then
BENCH_CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --chamfer-steps 1 --train-steps 2 --sampling-steps 2 --batched-scans 0"
timeout 300 $BENCH_CMD > gpurun_out/plain_bench.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
    --log-file gpurun_out/launches_r02.csv $BENCH_CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?"; tail -c 300 gpurun_out/ncu_launches.log
fi
timeout 300 python tools/ncu_once_r2.py > gpurun_out/plain_once.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on \
    -k regex:'fps_kernel|fps2_kernel|sa_mlp_tc_kernel|noise_mlp_kernel|train_gemm_kernel|train_wgrad_kernel|col_sums_kernel|pool_argmax_kernel|grid_query_kernel|chamfer_shard|nn_min_pair_kernel|bq_mask_kernel|bq_emit_kernel' \
    -s ${NCU_SKIP:-60} -c ${NCU_COUNT:-70} -o gpurun_out/prof_r02 python tools/ncu_once_r2.py > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"; tail -3 gpurun_out/ncu_full.log
# the report itself is too big to travel (64 MiB cap on gpurun_out): export what is read here, then drop it
ncu -i gpurun_out/prof_r02.ncu-rep --page raw --csv > gpurun_out/prof_r02_raw.csv 2> /dev/null
for k in noise_mlp_kernel train_gemm_kernel train_wgrad_kernel; do
  ncu -i gpurun_out/prof_r02.ncu-rep --page source --csv -k regex:$k > gpurun_out/prof_r02_source_$k.csv 2> /dev/null
done
rm -f gpurun_out/prof_r02.ncu-rep gpurun_out/fps2.ncu-rep
ls -la gpurun_out | tail -12

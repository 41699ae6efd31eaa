set -u
mkdir -p gpurun_out
timeout 300 python tools/ncu_batched_mlp.py > gpurun_out/plain_batched_mlp.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'sa_mlp_tc_kernel' -s 6 -c 3 -o gpurun_out/prof_batched_mlp python tools/ncu_batched_mlp.py > gpurun_out/ncu_bm.log 2>&1
echo "ncu exit $?"; cat gpurun_out/plain_batched_mlp.log | tail -5

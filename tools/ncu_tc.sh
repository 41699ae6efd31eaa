set -u
mkdir -p gpurun_out
timeout 300 python tools/ncu_once.py > gpurun_out/plain_once.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'sa_mlp_tc_kernel|fps_kernel' -s 5 -c 5 -o gpurun_out/prof_tc python tools/ncu_once.py > gpurun_out/ncu_tc.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_tc.log

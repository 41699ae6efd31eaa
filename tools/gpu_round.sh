#!/bin/bash
# One gpurun call: GPU parity tests, smoke, a short bench, kernel timings, then (NCU=1) ncu captures.
# Everything is written under gpurun_out/ so it comes back to the build container.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q --timeout=150 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" | tee -a gpurun_out/smoke.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"
tail -3 gpurun_out/bench.err
cat gpurun_out/bench.json
if [ "${REFARM:-0}" = "1" ]; then
  timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
  echo "bench reference exit $?"; cat gpurun_out/bench_ref.json
fi
timeout 300 python tools/prof_kernels.py ${PROF_ARGS:-} > gpurun_out/kernels.log 2>&1
echo "prof exit $?"; cat gpurun_out/kernels.log
if [ "${NCU:-0}" = "1" ]; then
  if [ "${NCU_LAUNCHES:-1}" = "1" ]; then
  BENCH_CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --chamfer-steps 1"
  timeout 300 $BENCH_CMD > gpurun_out/plain_bench.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
      --log-file gpurun_out/launches.csv $BENCH_CMD > gpurun_out/ncu_launches.log 2>&1
  echo "ncu launches exit $?"; tail -2 gpurun_out/ncu_launches.log
  fi
  timeout 300 python tools/ncu_once.py > gpurun_out/plain_once.log 2>&1 &&
  timeout 1200 ncu --set full --clock-control none --import-source on \
      -k regex:'fps_kernel|nn_min_kernel|nn_min_pair_kernel|nn_min_pair_arg_kernel|bq_mask_kernel|bq_emit_kernel|bq_small_kernel|sa_mlp_tc_kernel|mlp_layer_kernel|knn_kernel|vox_scatter_kernel|vox_hash_kernel|minmax_kernel' \
      -s ${NCU_SKIP:-30} -c ${NCU_COUNT:-30} -o gpurun_out/prof_full python tools/ncu_once.py > gpurun_out/ncu_full.log 2>&1
  echo "ncu full exit $?"; tail -5 gpurun_out/ncu_full.log
fi

#!/bin/bash
# One gpurun call: GPU parity tests, smoke, a short bench, kernel timings, then ncu captures.
# Everything is written under gpurun_out/ so it comes back to the build container.
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" | tee -a gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"
tail -3 gpurun_out/bench.err
cat gpurun_out/bench.json
timeout 600 python bench.py --steps 20 --warmup 5 --mlp-precision 1 --no-cpu-baseline > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err
echo "bench bf16 exit $?"; cat gpurun_out/bench_bf16.json
timeout 300 python tools/prof_kernels.py > gpurun_out/kernels.log 2>&1
echo "prof exit $?"; cat gpurun_out/kernels.log
if [ "${NCU:-0}" = "1" ]; then
  timeout 300 python tools/prof_kernels.py fps nn > gpurun_out/plain.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'fps_kernel|nn_min_kernel' -c 6 \
      -o gpurun_out/prof_fps_nn python tools/prof_kernels.py fps nn > gpurun_out/ncu.log 2>&1
  echo "ncu exit $?"; tail -5 gpurun_out/ncu.log
fi

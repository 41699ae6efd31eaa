#!/bin/bash
# Round-2 call 1: parity tests (incl. the new pins), smoke, bench (both arms), batched-MLP A/B (80-register narrow build), FPS probes.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q --timeout=200 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" | tee -a gpurun_out/smoke.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 300 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "bench reference exit $?"; cat gpurun_out/bench_ref.json
for t in "" "sa_mlp.regs=168" "sa_mlp.regs=80"; do
  echo "PCST_TUNE=$t"
  PCST_TUNE="$t" REPS=7 timeout 100 python tools/ncu_batched_mlp.py 2>&1 | grep -E "^SA[123]|Error|error" | cut -c1-100
done > gpurun_out/mlp_ab.log 2>&1
cat gpurun_out/mlp_ab.log
timeout 200 python tools/prof_kernels.py fpsprobe ball > gpurun_out/kernels.log 2>&1
cat gpurun_out/kernels.log

set -x
timeout 600 python -m pytest tests -m gpu -x -q -k "mlp or encoder or bench or set_abstraction or sa_" 2>&1 | tail -4
REPS=7 timeout 100 python tools/ncu_batched_mlp.py 2>&1 | grep -E "^SA[123]|Error|error" | cut -c1-100 | tee gpurun_out/mlp_batched_converged.log
timeout 300 python bench.py --steps 20 --warmup 3 --train-steps 0 --sampling-steps 0 2>gpurun_out/bench_err.log | tee gpurun_out/bench_mlp.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['ms_per_step'], d['mlp_precisions'])
for k in d['kernels'][:3]: print(k['kernel'], k['ms'])
"

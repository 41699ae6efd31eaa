timeout 600 python -m pytest tests -m gpu -x -q -k "mlp or encoder or bench or set_abstraction or sa_ or smoke" 2>&1 | tail -2
REPS=7 timeout 100 python tools/ncu_batched_mlp.py 2>&1 | grep -E "^SA[123]|Error|error" | cut -c1-100
timeout 200 python tools/mlp_tile_timeline.py 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 3 --train-steps 0 --sampling-steps 0 --no-cpu-baseline --chamfer-steps 1 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['batched']['value'])
for k in d['kernels'][:3]: print(k['kernel'][:50], k.get('ms'))
"

set -x
timeout 600 python -m pytest tests -m gpu -x -q -k "training_step" 2>&1 | tail -15
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --chamfer-steps 1 --train-steps 20 --sampling-steps 0 --batched-scans 0 2>gpurun_out/bench_err.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(json.dumps(d['train_c4'], indent=1)[:1500])
"
tail -5 gpurun_out/bench_err.log

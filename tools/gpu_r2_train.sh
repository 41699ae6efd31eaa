set -x
timeout 900 python -m pytest tests -m gpu -x -q -k "train" 2>&1 | tail -4
timeout 300 python bench.py --steps 10 --warmup 3 --train-steps 10 --sampling-steps 10 2>gpurun_out/bench_err.log | tee gpurun_out/bench_after_issuer.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['ms_per_step'], d['e2e'])
print(d['train_c4']['ms_per_step'], d['train_c4']['op_ms_eager'])
print(d['denoiser']); print(d['sampling_loop'])
"

set -x
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29547 bench.py --gpus 8 --steps 10 --warmup 3 --train-steps 10 --sampling-steps 0 2>gpurun_out/bench_n8_err.log | tail -1 > gpurun_out/bench_n8_late.json
echo "exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n8_late.json'))
print(d['value'], d['ms_per_step'], d['n_gpus'])
print(json.dumps(d.get('train_c4'))[:600])
print(json.dumps(d.get('chamfer_query_sharded'))[:900])
PY
tail -5 gpurun_out/bench_n8_err.log

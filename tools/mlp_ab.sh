# tcgen05 MLP at the throughput shape (32 scans x 16 384 points) + its per-tile timeline + the headline step, once per set
# of tuning knobs given as arguments (A/B measurements):  bash tools/mlp_ab.sh "" "sa_mlp.persistent=2" ...
[ $# -eq 0 ] && set -- ""
for t in "$@"; do
  echo "PCST_TUNE=$t"
  PCST_TUNE="$t" REPS=7 timeout 100 python tools/ncu_batched_mlp.py 2>&1 | grep -E "^SA[123]|Error|error" | cut -c1-90
  PCST_TUNE="$t,sa_mlp.regs=168" timeout 120 python tools/mlp_tile_timeline.py 2>&1 | tail -6
  PCST_TUNE="$t" timeout 120 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --batched-scans 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('step', d['ms_per_step'], d['e2e']['ms_per_step'])"
done

timeout 600 python -m pytest tests -m gpu -x -q -k "mlp or encoder or bench or set_abstraction or sa_ or smoke" 2>&1 | tail -2
for t in "" "sa_mlp.bf16_feats=2"; do
  echo "PCST_TUNE=$t"
  PCST_TUNE="$t" REPS=7 timeout 100 python tools/ncu_batched_mlp.py 2>&1 | grep -E "^SA[123]|Error|error" | cut -c1-100
done
timeout 200 python tools/mlp_tile_timeline.py 2>&1 | tail -3

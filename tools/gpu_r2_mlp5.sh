for t in "" "sa_mlp.reuse_h=2" "sa_mlp.early_gather=2" "" "sa_mlp.reuse_h=2"; do
  echo "PCST_TUNE=$t"
  PCST_TUNE="$t" REPS=7 timeout 100 python tools/ncu_batched_mlp.py 2>&1 | grep -E "^SA[12]|Error|error" | cut -c1-100
done

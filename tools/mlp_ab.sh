# tcgen05 MLP at the throughput shape: timings and the per-tile timeline, optionally for another build (PCST_LIB)
REPS=7 timeout 100 python tools/ncu_batched_mlp.py 2>&1 | grep -E "^SA[12]|Error|error" | cut -c1-90
TUNE=sa_mlp.regs=168 timeout 120 python tools/mlp_tile_timeline.py 2>&1 | tail -6

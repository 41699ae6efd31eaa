"""Chamfer 120k x 120k as bench.py times it (L2 flushed between calls), repeated, with several candidate-split counts."""
import os, sys, statistics
sys.path.insert(0, os.getcwd())
import torch
from pointcloud_style_transfer_b200 import _lib, ops, synthetic as S
from pointcloud_style_transfer_b200.models.losses import chamfer_distance_chunked_optimized
dev = torch.device("cuda:0")
x, y = S.lidar_scan(0).to(dev), S.lidar_scan(100).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timed(fn, steps):
    out = []
    for _ in range(steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        out.append(a.elapsed_time(b))
    return out
with torch.no_grad():
    for splits in (0, 5, 10, 15, 20, 0):
        _lib.set_tuning("nn_min.splits", splits)
        for _ in range(2): chamfer_distance_chunked_optimized(x, y)
        for rep in range(3):
            ms = timed(lambda: chamfer_distance_chunked_optimized(x, y), 5)
            ms2 = timed(lambda: ops.nn_min_pair(x, y, 0), 5)
            print(f"splits={splits} rep {rep}: chamfer call {statistics.mean(ms):.3f} ms (min {min(ms):.3f} max {max(ms):.3f}); nn_min_pair alone {statistics.mean(ms2):.3f}", flush=True)
_lib.set_tuning("nn_min.splits", 0)

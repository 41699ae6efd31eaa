"""Grid kNN against the brute-force sweep on a few cloud pairs (LiDAR scans, Gaussian noise, a half-noised scan):
CUDA-event ms per call of the default dispatch, the grid forced on, the sweep forced."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointcloud_style_transfer_b200 import _lib, ops, synthetic as S  # noqa: E402

dev = torch.device("cuda:0")
x, y = S.lidar_scan(0).to(dev), S.lidar_scan(100).to(dev)
perm = torch.randperm(120000, generator=torch.Generator().manual_seed(0))
r30 = x[:, perm[:30000].sort().values.to(dev)].contiguous()
q90 = x[:, perm[30000:].sort().values.to(dev)].contiguous()
g = torch.randn(1, 120000, 3, device=dev)
g30 = g[:, perm[:30000].sort().values.to(dev)].contiguous()
noisy = (0.5 * x + 0.866 * g).contiguous()
n30 = noisy[:, perm[:30000].sort().values.to(dev)].contiguous()
cases = {"lidar 90k x 30k k3": (q90, r30, 3), "lidar self 120k k9": (x, x, 9), "two scans 120k k3": (x, y, 3),
         "gaussian 120k x 30k k3": (g, g30, 3), "half-noised scan 120k x 30k k3": (noisy, n30, 3)}


def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for name, (q, r, k) in cases.items():
    t_auto = timeit(lambda: ops.knn(q, r, k))
    _lib.set_tuning("knn.grid", 1); t_grid = timeit(lambda: ops.knn(q, r, k))
    _lib.set_tuning("knn.grid", 2); t_sweep = timeit(lambda: ops.knn(q, r, k))
    _lib.set_tuning("knn.grid", 0)
    print(f"{name:32s} ms: default {t_auto:.3f}  grid {t_grid:.3f}  sweep {t_sweep:.3f}")

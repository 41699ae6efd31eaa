import os, sys, torch
sys.path.insert(0, os.getcwd())
from pointcloud_style_transfer_b200 import _lib, ops
dev = torch.device("cuda:0")
perm = torch.randperm(120000, generator=torch.Generator().manual_seed(0))
g = torch.randn(1, 120000, 3, device=dev)
g30 = g[:, perm[:30000].sort().values.to(dev)].contiguous()
_lib.set_tuning("knn.grid", 1)
for _ in range(2):
    ops.knn(g, g30, 3)
torch.cuda.synchronize()

"""One launch of every hot kernel at the BASELINE shapes (the command `ncu --set full` wraps).

    python tools/ncu_once.py            # warm-up pass + one measured pass of each kernel
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointcloud_style_transfer_b200 import ops, synthetic as S  # noqa: E402
from pointcloud_style_transfer_b200.models.pointnet2_encoder import PointNet2Encoder  # noqa: E402

dev = torch.device("cuda:0")
x = S.lidar_scan(0).to(dev)
y = S.lidar_scan(100).to(dev)
start = torch.tensor([1234], device=dev)
torch.manual_seed(42)
encs = [PointNet2Encoder(feature_dim=256, mlp_precision=p).eval().to(dev) for p in (0, 1)]
q, r = S.uniform_cloud(1, 1, 90000).to(dev), S.uniform_cloud(2, 1, 30000).to(dev)

for rep in range(2):  # pass 0 warms up (module load, function attributes), pass 1 is the one to read
    with torch.no_grad():
        for enc in encs:
            torch.manual_seed(1234)
            enc(x)                      # fps x2, ball query x2, sa_mlp x3 (fp32 then tcgen05)
        ops.nn_min(x, y, 0, False)      # Chamfer direction, loss form
        ops.nn_min(x, y, 0, True)       # with argmin (training)
        ops.nn_min_pair(x, y, 0)        # both Chamfer directions in one sweep
        ops.nn_min_pair_arg(x, y)       # ... with both argmins (training)
        ops.knn(q, r, 3)                # upsample_knn search
        ops.voxel_representatives(x, ops.minmax(x)[:, :3].contiguous(), torch.tensor([0.05], device=dev))
    torch.cuda.synchronize()
print("ncu_once ok")

"""Small shapes through every entry point (for `compute-sanitizer --tool memcheck`, one tool per gpurun call)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointcloud_style_transfer_b200 import ops, synthetic as S  # noqa: E402
from pointcloud_style_transfer_b200.models.diffusion_model import HierarchicalProcessor  # noqa: E402
from pointcloud_style_transfer_b200.models.losses import chamfer_distance_chunked_optimized  # noqa: E402
from pointcloud_style_transfer_b200.models.pointnet2_encoder import PointNet2Encoder  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
for prec in (0, 1):
    enc = PointNet2Encoder(feature_dim=256, mlp_precision=prec).eval().to(dev)
    with torch.no_grad():
        for B, N in ((2, 4099), (1, 20011)):
            print("encoder", prec, B, N, enc(S.uniform_cloud(1, B, N).to(dev)).shape, flush=True)
x, y = S.uniform_cloud(2, 2, 3001).to(dev), S.uniform_cloud(3, 2, 2500).to(dev)
print("chamfer", chamfer_distance_chunked_optimized(x, y), flush=True)
xg = x.clone().requires_grad_(True)
chamfer_distance_chunked_optimized(xg, y).sum().backward()
print("chamfer grad", xg.grad.abs().sum().item(), flush=True)
print("nn_min", ops.nn_min(x, y, 1, True)[0].sum().item(), ops.nn_min_pair(x, y, 1)[1].sum().item(), flush=True)
d, i = ops.knn(x, y, 9)
print("knn", d.sum().item(), ops.knn_interpolate(y, i[..., :3].contiguous(), d[..., :3].contiguous()).sum().item(), flush=True)
hp = HierarchicalProcessor(3001, 700)
down, idx = hp.downsample(x)
print("downsample", down.shape, hp.upsample_knn(down, x, idx).shape, flush=True)
torch.cuda.synchronize()
print("sanitize_small ok")

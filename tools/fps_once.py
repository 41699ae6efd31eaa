"""One look-ahead FPS launch on the 120k LiDAR scan (for ncu captures): PCST_TUNE selects the variant."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointcloud_style_transfer_b200 import ops, synthetic as S  # noqa: E402

dev = torch.device("cuda:0")
x = S.lidar_scan(0).to(dev)
start = torch.tensor([1234], device=dev)
for _ in range(3):
    ops.fps(x, 512, start)
torch.cuda.synchronize()

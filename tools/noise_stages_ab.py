"""Fused denoiser: time per call against the depth of the weight ring (tuning key noise.stages) and cluster size."""
import os
import sys

sys.path.insert(0, os.getcwd())
import torch

from pointcloud_style_transfer_b200 import _lib
from pointcloud_style_transfer_b200.config import Config
from pointcloud_style_transfer_b200.models import diffusion_model as DM

dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False
torch.manual_seed(3)
net = DM.PointCloudDiffusionModel(Config(), mlp_precision=1).to(dev).eval().noise_predictor
tt = torch.tensor([500, 37], device=dev)
st = torch.randn(2, 256, device=dev)
xc = torch.randn(2, 30000, 3, device=dev)


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


with torch.no_grad():
    net.fused_inference = False
    ref = net(xc, tt, st)
    net.fused_inference = True
    for stages in (3, 4, 5):
        for c in (1, 2, 4):
            _lib.set_tuning("noise.stages", stages)
            _lib.set_tuning("noise.cluster", c)
            out = net(xc, tt, st)
            err = ((out - ref).abs().max() / ref.abs().max()).item()
            print(f"stages={stages} cluster={c}: {timeit(lambda: net(xc, tt, st)) * 1e3:.1f} us  err {err:.2e}", flush=True)

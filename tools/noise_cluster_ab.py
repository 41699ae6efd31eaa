"""A/B of the fused denoiser's weight-stage sharing: cluster of 1 / 2 / 4 CTAs (tuning key noise.cluster), with the
parity of each against the torch fp32 module.  Run on a B200: python tools/noise_cluster_ab.py"""
import os
import sys

sys.path.insert(0, os.getcwd())
import torch

from pointcloud_style_transfer_b200 import _lib
from pointcloud_style_transfer_b200.config import Config
from pointcloud_style_transfer_b200.models import diffusion_model as DM

dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
torch.manual_seed(3)
net = DM.PointCloudDiffusionModel(Config(), mlp_precision=1).to(dev).eval().noise_predictor
tt = torch.tensor([500, 37], device=dev)
st = torch.randn(2, 256, device=dev)


def timeit(fn, reps=10):
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
    for n in (30000, 129, 1000, 16384):
        xc = torch.randn(2, n, 3, device=dev)
        net.fused_inference = False
        ref = net(xc, tt, st)
        net.fused_inference = True
        for c in (1, 2, 4):
            _lib.set_tuning("noise.cluster", c)
            out = net(xc, tt, st)
            torch.cuda.synchronize()
            err = ((out - ref).abs().max() / ref.abs().max()).item()
            print(f"N=2x{n} cluster={c}: {timeit(lambda: net(xc, tt, st)) * 1e3:.1f} us  max err / max |ref| = {err:.2e}", flush=True)
_lib.set_tuning("noise.cluster", 0)

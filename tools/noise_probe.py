"""Where the fused denoiser's time goes: the kernel with parts switched off (tuning key noise.probe; results are garbage,
only the time is meaningful).  Times the C-ABI call alone (prep kernel + chain) with CUDA events."""
import os
import sys

sys.path.insert(0, os.getcwd())
import torch

from pointcloud_style_transfer_b200 import _lib, ops
from pointcloud_style_transfer_b200.config import Config
from pointcloud_style_transfer_b200.models import diffusion_model as DM

dev = torch.device("cuda:0")
torch.manual_seed(3)
net = DM.PointCloudDiffusionModel(Config(), mlp_precision=1).to(dev).eval().noise_predictor
tt = torch.tensor([500, 37], device=dev)
st = torch.randn(2, 256, device=dev)
packed = net._packed_params()
F, T, nb = net.style_proj.out_features, net.time_proj.in_features, len(net.layers)


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


names = {0: "full", 1: "no weight streaming", 2: "empty epilogues", 4: "no MMAs", 3: "no weights, empty epilogues",
         5: "no weights, no MMAs", 6: "empty epilogues, no MMAs", 7: "barriers only", 15: "barriers only, no fence.proxy.async",
         23: "barriers only, arrive instead of tcgen05.commit", 31: "barriers only, neither", 135: "barriers only, no tcgen05.fence", 256: "full, test_wait spin", 263: "barriers only, test_wait spin"}
with torch.no_grad():
    for n in (64, 30000):
        xc = torch.randn(2, n, 3, device=dev)
        for probe in (0,):
            _lib.set_tuning("noise.probe", probe)
            t = timeit(lambda: ops.noise_predictor(xc, tt, st, packed, F, T, nb))
            print(f"N=2x{n} ({2 * ((n + 127) // 128)} tiles) probe={probe} ({names[probe]}): {t * 1e3:.1f} us", flush=True)
xc = torch.randn(2, 30000, 3, device=dev)
for probe in (64,):
    print("clock stamps, probe", probe, flush=True)
    _lib.set_tuning("noise.probe", probe)
    with torch.no_grad():
        ops.noise_predictor(xc, tt, st, packed, F, T, nb)
    torch.cuda.synchronize()
_lib.set_tuning("noise.probe", 0)

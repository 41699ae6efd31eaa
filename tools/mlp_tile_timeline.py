"""Per-tile timeline of the tcgen05 shared-MLP kernel at the throughput shape (32 scans x 16 384 points), from the SM-clock
stamps pcst_sa_mlp_set_probe records: mean cycles per phase (gather, wait for the accumulator, epilogue, ...) per stage.

    [PCST_TUNE=key=value,...] python tools/mlp_tile_timeline.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointcloud_style_transfer_b200 import _lib, synthetic as S  # noqa: E402
from pointcloud_style_transfer_b200.models.pointnet2_encoder import PointNet2Encoder  # noqa: E402

dev = torch.device("cuda:0")
B, N = int(os.environ.get("SCANS", "32")), 16384
x = torch.cat([S.lidar_scan(i, N) for i in range(B)], 0).to(dev)
torch.manual_seed(42)
enc = PointNet2Encoder(feature_dim=256, mlp_precision=1).eval().to(dev)
lib = _lib.load()
TILES = 4096
with torch.no_grad():
    torch.manual_seed(1)
    enc(x)  # warm
    stage_inputs = {}
    torch.manual_seed(1)
    l1_xyz, l1_pts = enc.sa1(x, None)
    for name, fn in (("SA1", lambda: enc.sa1(x, None)), ("SA2", lambda: enc.sa2(l1_xyz, l1_pts.permute(0, 2, 1)))):
        buf = torch.zeros(TILES, 16, dtype=torch.int64, device=dev)
        torch.manual_seed(1)
        fn()
        torch.cuda.synchronize()
        lib.pcst_sa_mlp_set_probe(buf.data_ptr(), TILES)
        torch.manual_seed(1)
        fn()
        torch.cuda.synchronize()
        lib.pcst_sa_mlp_set_probe(None, 0)
        t = buf.cpu().numpy()
        t = t[t[:, 0] != 0]
        if len(t) == 0:
            print(f"{name}: no stamps (the one-tile-per-CTA build of the kernel does not record them)")
            continue
        nst = int((t[0, :15] != 0).sum())
        d = np.diff(t[:, :nst], axis=1).astype(np.float64)
        # stamps: tile start, then per epilogue (accumulator ready, epilogue done)
        names = [f"{'gather+mma' if i == 0 else ('mma' if i % 2 == 0 else 'epilogue')}{i // 2}" for i in range(nst - 1)]
        span = t[:, nst - 1].max() - t[:, 0].min()
        print(f"{name}: {len(t)} tiles, {nst} stamps; kernel span {span} cycles; mean tile {d.sum(1).mean():.0f} cycles")
        print("   " + "  ".join(f"{n} {v:.0f}" for n, v in zip(names, d.mean(0))))
        # tiles per SM in flight: back-to-back gap between consecutive tiles of the same CTA cannot be seen here; report
        # the spread of tile starts instead
        sm = t[:, 15]
        per_sm = np.bincount(sm.astype(np.int64), minlength=148)
        print(f"   tiles per SM: min {per_sm.min()} max {per_sm.max()}")

"""The tcgen05 shared-MLP kernel at a throughput shape (BASELINE config 4 per-GPU slice x 8: 32 scans of 16 384 points):
SA1 = 524 288 rows (4096 tiles), SA2 = 262 144 rows (2048 tiles), SA3 = 4096 rows (32 tiles).  Used under
`ncu --set full -k regex:sa_mlp_tc_kernel` for the tensor-pipe utilisation the kernel reaches when the machine is full,
and stand-alone for CUDA-event timings."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointcloud_style_transfer_b200 import ops, synthetic as S  # noqa: E402
from pointcloud_style_transfer_b200.models.pointnet2_encoder import PointNet2Encoder  # noqa: E402

dev = torch.device("cuda:0")
B, N = 32, 16384
x = torch.cat([S.lidar_scan(i, N) for i in range(B)], 0).to(dev)
torch.manual_seed(42)
enc = PointNet2Encoder(feature_dim=256, mlp_precision=1).eval().to(dev)
with torch.no_grad():
    logs = []
    for rep in range(int(os.environ.get("REPS", "3"))):
        ops.start_event_log()
        torch.manual_seed(1)
        enc(x)
        logs.append(ops.stop_event_log())
# median over the passes after the first (one pass under ncu: REPS=1)
keep = logs[1:] or logs
log = {k: [sorted(l[k][i] for l in keep)[len(keep) // 2] for i in range(len(v))] for k, v in logs[-1].items()}
print("per-op us (median pass):", {k: [round(v * 1e3, 1) for v in vs] for k, vs in log.items()})
rows = [B * 512 * 32, B * 128 * 64, B * 128]
macs = [3 * 64 + 64 * 64 + 64 * 128, 131 * 128 + 128 * 128 + 128 * 256, 259 * 256 + 256 * 512 + 512 * 256]
for name, r, m, t in zip(("SA1", "SA2", "SA3"), rows, macs, log["pcst_sa_mlp_max_f32"]):
    print(f"{name}: {r} rows, {2 * r * m / 1e9:.2f} GFLOP, {t * 1e3:.1f} us -> {2 * r * m / (t * 1e-3) / 1e12:.1f} TFLOP/s (algorithmic, bf16)")

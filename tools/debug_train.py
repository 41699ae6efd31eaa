"""Debug aid: one train-mode SetAbstraction stage on the native kernels, every intermediate of the saved blob against a
torch fp32 evaluation of the same formulas."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointcloud_style_transfer_b200 import _lib, ops, synthetic as S  # noqa: E402
from pointcloud_style_transfer_b200.models.pointnet2_encoder import SetAbstraction, query_ball_point  # noqa: E402

dev = torch.device("cuda:0")
B, N, Sq, K, D, mlp = 2, 600, 32, 16, 4, [32, 32, 64]
if len(sys.argv) > 1:
    B, N, Sq, K, D = [int(v) for v in sys.argv[1:6]]
    mlp = [int(v) for v in sys.argv[6:9]]
torch.manual_seed(3)
GA = Sq == 0
sa = SetAbstraction(Sq or None, 0.4, K, in_channel=D, mlp=mlp, group_all=GA).to(dev).train()
x = S.uniform_cloud(5, B, N).to(dev)
f = torch.randn(B, N, D, device=dev) if D else None
start = torch.zeros(B, dtype=torch.long, device=dev)
if GA:
    new_xyz = idx = None
    grouped = torch.cat([x, f], -1) if D else x
    Sq, K = 1, N
else:
    _, new_xyz = ops.fps(x, Sq, start)
    idx = query_ball_point(0.4, K, x, new_xyz)
    grouped = ops.group(x, f, new_xyz, idx)          # [B,S,K,3+D]
rows = grouped.reshape(-1, 3 + D)
R = rows.shape[0]

out = ops.sa_mlp_train(x, f, new_xyz, idx, list(sa.mlp_convs), list(sa.mlp_bns))
torch.cuda.synchronize()
# recover the saved blob through autograd's graph
saved = out.grad_fn.saved_tensors[4]
off = 0
Z = []
for c in mlp:
    Z.append(saved[off: off + R * c * 4].view(torch.float32).reshape(R, c))
    off += (R * c * 4 + 255) // 256 * 256
stats = []
for c in mlp:
    stats.append(saved[off: off + 4 * c * 4].view(torch.float32).reshape(4, c))
    off += (4 * c * 4 + 255) // 256 * 256

xin = rows
for l, (conv, bn) in enumerate(zip(sa.mlp_convs, sa.mlp_bns)):
    w = conv.weight.reshape(conv.out_channels, -1)
    z_ref = xin.to(torch.bfloat16).float() @ w.to(torch.bfloat16).float().t() + conv.bias
    print(f"layer {l}: Z rel err {((Z[l] - z_ref).norm() / z_ref.norm()).item():.3e}  max abs {(Z[l] - z_ref).abs().max().item():.3e}")
    mean, var = Z[l].mean(0), Z[l].var(0, unbiased=False)
    print(f"   mean err {(stats[l][0] - mean).abs().max().item():.3e}  invstd rel err "
          f"{((stats[l][1] - 1 / torch.sqrt(var + bn.eps)).abs() / stats[l][1].abs()).max().item():.3e}")
    a = bn.weight * stats[l][1]
    b = bn.bias - stats[l][0] * a
    print(f"   a err {(stats[l][2] - a).abs().max().item():.3e} b err {(stats[l][3] - b).abs().max().item():.3e}")
    xin = torch.relu(Z[l] * a + b)   # continue from the kernel's own Z so that errors do not compound in the check
pooled = xin.reshape(B * Sq, K, -1).max(1)[0]
print("pooled rel err", ((out.reshape(B * Sq, -1) - pooled).norm() / pooled.norm()).item())

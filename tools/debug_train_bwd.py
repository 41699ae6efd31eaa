"""Debug aid: one train-mode SetAbstraction stage, native (precision 1) vs the fp32 torch composition: per-tensor relative
L2 error of the output and of every gradient."""
import copy
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointcloud_style_transfer_b200 import synthetic as S  # noqa: E402
from pointcloud_style_transfer_b200.models.pointnet2_encoder import SetAbstraction  # noqa: E402

dev = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
B, N, Sq, K, D = [int(v) for v in sys.argv[1:6]]
mlp = [int(v) for v in sys.argv[6:9]]
GA = Sq == 0
torch.manual_seed(3)
sa = SetAbstraction(Sq or None, 0.4, K or None, in_channel=D, mlp=mlp, group_all=GA).to(dev).train()
with torch.no_grad():
    for bn in sa.mlp_bns:
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.normal_(0, 0.2)
sb = copy.deepcopy(sa)
sa.mlp_precision, sb.mlp_precision = int(os.environ.get('PREC', '1')), 0
sb.train_backend = 'torch'
x = S.uniform_cloud(5, B, N).to(dev)
f = torch.randn(B, N, D, device=dev).requires_grad_(True) if D else None
f2 = f.detach().clone().requires_grad_(True) if D else None


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


outs = []
for mod, feats in ((sa, f), (sb, f2)):
    torch.manual_seed(11)
    _, out = mod(x, feats)
    w = torch.randn(out.shape, generator=torch.Generator().manual_seed(6)).to(dev)
    (out * w).sum().backward()
    outs.append(out.detach())
print("out rel", rel(outs[0], outs[1]), "max abs", (outs[0] - outs[1]).abs().max().item(), "frac differing > 0.05:",
      ((outs[0] - outs[1]).abs() > 0.05).float().mean().item())
if D:
    print("feat grad rel", rel(f.grad, f2.grad))
for (name, pa), (_, pb) in zip(sa.named_parameters(), sb.named_parameters()):
    print(f"{name:24s} rel {rel(pa.grad, pb.grad):.3e}  |ref| {pb.grad.norm().item():.3e}")

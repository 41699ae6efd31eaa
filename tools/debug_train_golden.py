"""Debug aid: the train-mode encoder step against tests/golden/train_encoder.npz (the reference's own CPU fp32 result):
relative L2 error of the feature, every gradient and the BatchNorm buffers.  PREC=0|1 selects the operand mode."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointcloud_style_transfer_b200.models.pointnet2_encoder import PointNet2Encoder  # noqa: E402

g = dict(np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "train_encoder.npz")))
dev = torch.device("cuda:0")
enc = PointNet2Encoder(feature_dim=128, mlp_precision=int(os.environ.get("PREC", "0")))
enc.load_state_dict({k[4:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("sd0.")})
enc = enc.to(dev).train()
x, coef = torch.from_numpy(g["x"]).to(dev), torch.from_numpy(g["coef"]).to(dev)
torch.manual_seed(1234)
feat = enc(x)
(feat * coef).sum().backward()
rel = lambda a, b: float(np.linalg.norm(a.astype(np.float64) - b) / max(np.linalg.norm(b.astype(np.float64)), 1e-30))
f = feat.detach().cpu().numpy()
print("feature rel", rel(f, g["feature"]), "max abs", np.abs(f - g["feature"]).max())
for name, prm in enc.named_parameters():
    if name.endswith("convs.0.bias") or name.endswith("convs.1.bias") or name.endswith("convs.2.bias"):
        continue
    print(f"{name:28s} rel {rel(prm.grad.cpu().numpy(), g['grad.' + name]):.3e}")
for k, v in g.items():
    if k.startswith("sd1.") and "num_batches" not in k:
        print(f"{k:36s} rel {rel(enc.state_dict()[k[4:]].cpu().numpy(), v):.3e}")

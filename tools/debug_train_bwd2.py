"""Debug aid: backward of one train-mode stage through the C ABI with a caller-visible workspace; every intermediate
(dY_1, dY_0, per-layer vec = g|m1|m2|mean|invstd, weight gradients) against torch formulas evaluated on the kernel's own Z."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointcloud_style_transfer_b200 import _lib, ops, synthetic as S  # noqa: E402
from pointcloud_style_transfer_b200.models.pointnet2_encoder import SetAbstraction, query_ball_point  # noqa: E402

dev = torch.device("cuda:0")
B, N, Sq, K, D = [int(v) for v in sys.argv[1:6]]
mlp = [int(v) for v in sys.argv[6:9]]
torch.manual_seed(3)
sa = SetAbstraction(Sq, 0.4, K, in_channel=D, mlp=mlp).to(dev).train()
with torch.no_grad():
    for bn in sa.mlp_bns:
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.normal_(0, 0.2)
x = S.uniform_cloud(5, B, N).to(dev)
f = torch.randn(B, N, D, device=dev) if D else None
_, new_xyz = ops.fps(x, Sq, torch.zeros(B, dtype=torch.long, device=dev))
idx = query_ball_point(0.4, K, x, new_xyz)
rows = ops.group(x, f, new_xyz, idx).reshape(-1, 3 + D)
R, G = rows.shape[0], B * Sq
lib = _lib.load()
al = lambda v: (v + 255) // 256 * 256
c3 = (ctypes.c_int * 3)(*mlp)
m = _lib.Mlp3Train()
ps = []
for l, (conv, bn) in enumerate(zip(sa.mlp_convs, sa.mlp_bns)):
    w = conv.weight.detach().reshape(conv.out_channels, -1).contiguous()
    ps.append((w, conv.bias.detach(), bn.weight.detach(), bn.bias.detach()))
    m.w[l], m.bias[l], m.gamma[l], m.beta[l] = [t.data_ptr() for t in ps[-1]]
    m.cout[l] = mlp[l]
m.eps, m.momentum = 1e-5, 0.1
P = lambda t: ctypes.c_void_p(0 if t is None else t.data_ptr())
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
saved = torch.zeros(lib.pcst_sa_mlp_train_saved_bytes(B, Sq, K, D, c3), dtype=torch.uint8, device=dev)
wsf = torch.zeros(lib.pcst_sa_mlp_train_workspace_bytes(B, Sq, K, D, c3, 0), dtype=torch.uint8, device=dev)
wsb = torch.zeros(lib.pcst_sa_mlp_train_workspace_bytes(B, Sq, K, D, c3, 1), dtype=torch.uint8, device=dev)
out = torch.empty(B, Sq, mlp[2], device=dev)
_lib.check(lib.pcst_sa_mlp_max_bnstats_bf16(P(x), P(f), P(new_xyz), P(idx), B, N, Sq, K, D, ctypes.byref(m), P(out), P(saved),
                                            saved.numel(), P(wsf), wsf.numel(), st))
gout = torch.randn(B, Sq, mlp[2], device=dev)
gr = _lib.Mlp3Grads()
grads = []
for l in range(3):
    gl = [torch.zeros_like(t) for t in ps[l]]
    gr.w[l], gr.bias[l], gr.gamma[l], gr.beta[l] = [t.data_ptr() for t in gl]
    grads.append(gl)
gg = torch.zeros(B, Sq, K, 3 + D, device=dev)
_lib.check(lib.pcst_sa_mlp_max_bwd_bf16(P(x), P(f), P(new_xyz), P(idx), B, N, Sq, K, D, ctypes.byref(m), P(saved), saved.numel(),
                                        P(gout), ctypes.byref(gr), P(gg), P(wsb), wsb.numel(), st))
torch.cuda.synchronize()
# ---- unpack the saved blob and the backward workspace (layout of train_plan) ----
off = 0
Z, stat = [], []
for c in mlp:
    Z.append(saved[off: off + R * c * 4].view(torch.float32).reshape(R, c)); off += al(R * c * 4)
for c in mlp:
    stat.append(saved[off: off + 16 * c].view(torch.float32).reshape(4, c)); off += al(16 * c)
argmax = saved[off: off + G * mlp[2] * 4].view(torch.int32).reshape(G, mlp[2])
cin = [3 + D, mlp[0], mlp[1]]
kp = [(c + 15) // 16 * 16 for c in cin]
off = 0
for l in range(3):
    off += al(mlp[l] * kp[l] * 2)
dY = []
for l in range(2):
    dY.append(wsb[off: off + R * mlp[l] * 2].view(torch.bfloat16).reshape(R, mlp[l]).float()); off += al(R * mlp[l] * 2)
vec = []
for l in range(3):
    off += al(2 * mlp[l] * 8)
    vec.append(wsb[off: off + 20 * mlp[l]].view(torch.float32).reshape(5, mlp[l])); off += al(20 * mlp[l])


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


# ---- torch formulas on the kernel's own Z ----
X = [rows]
for l in range(3):
    X.append(torch.relu(Z[l] * stat[l][2] + stat[l][3]))
dy = torch.zeros(G, K, mlp[2], device=dev)
am = argmax.long()
valid = am >= 0
gi, ci = torch.nonzero(valid, as_tuple=True)
dy[gi, am[gi, ci], ci] = gout.reshape(G, -1)[gi, ci]
dy = dy.reshape(R, -1)
for l in (2, 1, 0):
    xh = (Z[l] - stat[l][0]) * stat[l][1]
    g = ps[l][2] * stat[l][1]
    m1, m2 = dy.mean(0), (dy * xh).mean(0)
    print(f"layer {l}: vec g {rel(vec[l][0], g):.2e} m1 {rel(vec[l][1], m1):.2e} m2 {rel(vec[l][2], m2):.2e} "
          f"mean {rel(vec[l][3], stat[l][0]):.2e} istd {rel(vec[l][4], stat[l][1]):.2e}")
    dz = g * (dy - m1 - xh * m2)
    dw = dz.t() @ X[l]
    print(f"   dW rel {rel(grads[l][0], dw):.3e}   dgamma {rel(grads[l][2], (dy * xh).sum(0)):.2e}  dbeta {rel(grads[l][3], dy.sum(0)):.2e}")
    dx = dz @ ps[l][0]
    if l > 0:
        dy = dx * (X[l] > 0)
        print(f"   dY_{l-1} rel {rel(dY[l-1], dy):.3e}")
    else:
        print(f"   grad_grouped rel {rel(gg.reshape(R, -1), dx):.3e}")

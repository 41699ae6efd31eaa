"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE (test infrastructure).

Runs only in the build container, where the reference is mounted read-only at /root/reference;
the GPU box never has it, so the fixtures written here are committed.  Usage:

    python oracle/gen_golden.py            # rewrites tests/golden/*.npz

Every array is produced by the reference's own functions (file:line cited per block); inputs come
from ``pointcloud_style_transfer_b200.synthetic`` and are stored next to the outputs (or, for the
120k-point scans, re-generated from the seed and guarded by a SHA-256 of their bytes).
"""
from __future__ import annotations

import hashlib
import importlib.util
import os
import sys
import tempfile

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("PCST_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.path.insert(0, REPO)

from pointcloud_style_transfer_b200 import synthetic as S  # noqa: E402

OUT = os.path.join(REPO, "tests", "golden")


def load_reference():
    """Import the reference's hot-path modules; metrics.py is loaded by path
    because the ``evaluation`` package drags in matplotlib/open3d."""
    os.chdir(tempfile.mkdtemp())  # config/config.py:64-67 creates directories in the cwd
    import models.pointnet2_encoder as enc
    import models.losses as losses
    from models.diffusion_model import HierarchicalProcessor

    spec = importlib.util.spec_from_file_location("ref_metrics", os.path.join(REF, "evaluation", "metrics.py"))
    metrics = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(metrics)
    return enc, losses, HierarchicalProcessor, metrics


def sha(t: torch.Tensor) -> str:
    return hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()


def save(name: str, **arrays):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}.npz  {os.path.getsize(path) / 1024:.0f} KiB")


def i32(t):
    return t.numpy().astype(np.int32)


def gen_voxel(HP):
    """Voxel-grid downsample (models/diffusion_model.py:69-122): the reference's outputs with the true RNG (seeded),
    and its per-voxel representative indices, exposed by running it once more with ``torch.randperm`` replaced by
    the identity permutation and ``torch.unique`` wrapped to record the number of voxels."""
    clouds = torch.cat([S.lidar_scan(2, 20000), S.uniform_cloud(17, 1, 20000) * 1.5], 0)   # [2,20000,3]
    target = 5000
    hp = HP(20000, target)
    torch.manual_seed(77)
    down, idx = hp.downsample(clouds)
    real_randperm, real_unique = torch.randperm, torch.unique
    counts = []

    def unique_spy(*a, **k):
        out = real_unique(*a, **k)
        counts.append(int(out[0].numel()))
        return out

    torch.randperm = lambda n, device=None: torch.arange(n)
    torch.unique = unique_spy
    try:
        _, idx_id = hp.downsample(clouds)
    finally:
        torch.randperm, torch.unique = real_randperm, real_unique
    sizes = []
    for b in range(2):
        pts = clouds[b]
        r = pts.max(axis=0)[0] - pts.min(axis=0)[0]
        r[r < 1e-6] = 1.0
        sizes.append(float((r.prod() / target) ** (1 / 3) * 1.2))
    assert torch.equal(down, torch.stack([clouds[b][idx[b]] for b in range(2)]))
    save("voxel_downsample", clouds=clouds.numpy(), target=np.int64(target), indices=i32(idx),
         rep0=i32(idx_id[0, :counts[0]]), rep1=i32(idx_id[1, :counts[1]]), voxel_size=np.array(sizes, np.float32),
         # a cloud that is already small enough is returned unchanged (:70-72)
         small_indices=i32(HP(100, 200).downsample(clouds[:, :100])[1]))


def gen_train(enc):
    """TRAIN-mode encoder step (models/pointnet2_encoder.py:74,106-112 with nn.BatchNorm2d batch statistics; the
    forward / backward of training/trainer.py:78-117 restricted to the encoder): the reference's own forward output,
    parameter gradients (autograd) and updated BatchNorm buffers for a seeded 2 x 2048-point batch."""
    torch.manual_seed(42)
    model = enc.PointNet2Encoder(feature_dim=128)
    g = torch.Generator().manual_seed(7)
    with torch.no_grad():
        for mod in model.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.weight.copy_(torch.rand(mod.num_features, generator=g) * 0.5 + 0.75)
                mod.bias.copy_(torch.randn(mod.num_features, generator=g) * 0.1)
    sd0 = {k: v.clone().numpy() for k, v in model.state_dict().items()}
    model.train()
    x = S.uniform_cloud(21, 2, 2048)
    coef = torch.randn(2, 128, generator=g)
    torch.manual_seed(1234)
    start1 = torch.randint(0, 2048, (2,), dtype=torch.long)
    start2 = torch.randint(0, 512, (2,), dtype=torch.long)
    torch.manual_seed(1234)
    with torch.enable_grad():
        feat = model(x)
        (feat * coef).sum().backward()
    arrays = dict(x=x.numpy(), coef=coef.numpy(), start1=start1.numpy(), start2=start2.numpy(), feature=feat.detach().numpy())
    for k, v in sd0.items():
        arrays["sd0." + k] = v
    for k, v in model.state_dict().items():
        if "running" in k or "num_batches" in k:
            arrays["sd1." + k] = v.numpy()
    for k, prm in model.named_parameters():
        arrays["grad." + k] = prm.grad.numpy()
    save("train_encoder", **arrays)


def gen_noise_predictor():
    """NoisePredictor.forward (models/diffusion_model.py:38-61), eval mode, on a small configuration (feature_dim 64,
    time_embed_dim 32: a 0.3 MB fixture; the default 256 / 128 configuration is compared on the GPU against the same
    formulation, which test_noise_predictor_module_equals_reference_formulation ties to this one)."""
    import models.diffusion_model as rdm
    from config.config import Config as RefConfig

    cfg = RefConfig()
    cfg.feature_dim, cfg.time_embed_dim = 64, 32
    torch.manual_seed(5)
    net = rdm.NoisePredictor(cfg).eval()
    g = torch.Generator().manual_seed(6)
    x = torch.randn(3, 333, 3, generator=g)
    t = torch.tensor([0, 17, 999])
    style = torch.randn(3, 64, generator=g)
    style[2] = 0.0                      # the unconditional branch of classifier-free guidance
    out = net(x, t, style)
    save("noise_predictor", x=x.numpy(), t=t.numpy(), style=style.numpy(), out=out.numpy(),
         **{"sd." + k: v.numpy() for k, v in net.state_dict().items()})


def main():
    torch.set_grad_enabled(False)
    os.makedirs(OUT, exist_ok=True)
    enc, losses, HP, metrics = load_reference()
    if "--only-voxel" in sys.argv:
        gen_voxel(HP)
        return
    if "--only-train" in sys.argv:
        gen_train(enc)
        return
    if "--only-noise" in sys.argv:
        gen_noise_predictor()
        return
    gen_train(enc)
    gen_noise_predictor()
    gen_voxel(HP)
    M = metrics.PointCloudMetrics("cpu")

    # ---- C1: PointNet2Encoder 2x4096, eval, F=256 (models/pointnet2_encoder.py:114-131) ----
    torch.manual_seed(42)
    model = enc.PointNet2Encoder(feature_dim=256)
    # non-trivial BatchNorm statistics so the eval-mode fold is actually exercised
    g = torch.Generator().manual_seed(7)
    for mod in model.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.copy_(torch.randn(mod.num_features, generator=g) * 0.1)
            mod.running_var.copy_(torch.rand(mod.num_features, generator=g) * 0.5 + 0.75)
            mod.weight.copy_(torch.rand(mod.num_features, generator=g) * 0.5 + 0.75)
            mod.bias.copy_(torch.randn(mod.num_features, generator=g) * 0.1)
    model.eval()
    sd = {k: v.numpy() for k, v in model.state_dict().items() if "num_batches" not in k}
    x = S.uniform_cloud(0, 2, 4096)
    torch.manual_seed(1234)
    start1 = torch.randint(0, 4096, (2,), dtype=torch.long)
    start2 = torch.randint(0, 512, (2,), dtype=torch.long)
    torch.manual_seed(1234)
    feat = model(x)
    # the stage-by-stage intermediates, from the reference's own functions with the same RNG stream
    torch.manual_seed(1234)
    fps1 = enc.farthest_point_sample(x, 512)
    l1_xyz = enc.index_points(x, fps1)
    grp1 = enc.query_ball_point(0.2, 32, x, l1_xyz)
    fps2 = enc.farthest_point_sample(l1_xyz, 128)
    l2_xyz = enc.index_points(l1_xyz, fps2)
    grp2 = enc.query_ball_point(0.4, 64, l1_xyz, l2_xyz)
    torch.manual_seed(1234)
    _, l1_pts = model.sa1(x, None)
    _, l2_pts = model.sa2(l1_xyz, l1_pts.permute(0, 2, 1))
    assert torch.equal(fps1[:, 0], start1) and torch.equal(fps2[:, 0], start2)
    save("c1_encoder", x=x.numpy(), start1=start1.numpy(), start2=start2.numpy(), fps1=i32(fps1), fps2=i32(fps2),
         group1=i32(grp1), group2=i32(grp2), l1_points=l1_pts.numpy(), l2_points=l2_pts.numpy(),
         feature=feat.numpy(), **{"sd." + k: v for k, v in sd.items()})

    # ---- square_distance sample (models/pointnet2_encoder.py:8-15) ----
    a, b = S.uniform_cloud(3, 2, 96), S.uniform_cloud(4, 2, 333)
    save("square_distance", src=a.numpy(), dst=b.numpy(), out=enc.square_distance(a, b).numpy())

    # ---- C1 Chamfer / metrics 2x4096x4096 (models/losses.py:8-63, evaluation/metrics.py:20-44,90-105) ----
    p, t = S.uniform_cloud(0, 2, 4096), S.uniform_cloud(100, 2, 4096)
    t = t[:, :3900].contiguous()  # ragged N != M
    # per-point minima exactly as the loss forms them (losses.py:24-25,36-41 / 53-58, unchunked)
    psq, tsq = (p ** 2).sum(-1, keepdim=True), (t ** 2).sum(-1, keepdim=True).transpose(1, 2)
    d1 = torch.clamp(psq + tsq + (-2 * torch.bmm(p, t.transpose(1, 2))), min=0).min(dim=2)[0]
    d2 = torch.clamp(tsq.transpose(1, 2) + psq.transpose(1, 2) + (-2 * torch.bmm(t, p.transpose(1, 2))), min=0).min(dim=2)[0]
    cd = losses.chamfer_distance_chunked_optimized(p, t)
    assert torch.allclose(d1.mean(1) + d2.mean(1), cd, rtol=1e-6)
    dm = torch.cdist(p, t, p=2)
    save("c1_chamfer", pred=p.numpy(), target=t.numpy(), loss_rowmin=d1.numpy(), loss_colmin=d2.numpy(),
         chamfer_loss=cd.numpy(), chamfer_loss_chunk100=losses.chamfer_distance_chunked_optimized(p, t, 100).numpy(),
         metric_rowmin=dm.min(dim=2)[0].numpy(), metric_colmin=dm.min(dim=1)[0].numpy(),
         metric_cd=M.chamfer_distance(p, t).numpy(), metric_cd_oneway=M.chamfer_distance(p, t, bidirectional=False).numpy(),
         metric_hausdorff=M.hausdorff_distance(p, t).numpy())

    # ---- lattice inputs: order-independent exactness + exact ties ----
    xq = S.lattice(S.uniform_cloud(5, 2, 2048), 64)  # coarse lattice -> many FPS / radius ties
    torch.manual_seed(99)
    fq = enc.farthest_point_sample(xq, 256)
    st = fq[:, 0].clone()
    nq = enc.index_points(xq, fq)
    gq = enc.query_ball_point(0.25, 16, xq, nq)
    yq = S.lattice(S.uniform_cloud(6, 2, 1500), 64)
    psq, tsq = (xq ** 2).sum(-1, keepdim=True), (yq ** 2).sum(-1, keepdim=True).transpose(1, 2)
    dq = torch.clamp(psq + tsq + (-2 * torch.bmm(xq, yq.transpose(1, 2))), min=0)
    save("lattice", x=xq.numpy(), y=yq.numpy(), start=st.numpy(), fps=i32(fq), group=i32(gq),
         loss_rowmin=dq.min(dim=2)[0].numpy(), loss_colmin=dq.min(dim=1)[0].numpy(),
         chamfer_loss=losses.chamfer_distance_chunked_optimized(xq, yq).numpy())

    # ---- edge cases of FPS / ball query ----
    xe = S.uniform_cloud(8, 1, 300)
    qe = torch.cat([xe[:, :5], xe[:, :3] + 10.0], 1).contiguous()  # last 3 queries have empty balls
    save("edge_ball_query", x=xe.numpy(), q=qe.numpy(),
         tiny=i32(enc.query_ball_point(0.05, 8, xe, qe)),     # ragged rows, padded with the first hit
         huge=i32(enc.query_ball_point(5.0, 300, xe, qe[:, :5].contiguous())),  # nsample == N, every point in the ball
         # (nsample > N raises IndexError in the reference, :58 -- the drop-in raises too)
         r_tiny=np.float64(0.05), r_huge=np.float64(5.0))
    xd = torch.cat([S.uniform_cloud(9, 1, 40)] * 3, 1).contiguous()  # 120 pts, each position 3 times
    torch.manual_seed(5)
    fd = enc.farthest_point_sample(xd, 100)  # npoint > distinct positions -> all-zero distance ties -> index 0
    xs = torch.zeros(1, 50, 3)
    torch.manual_seed(5)
    fs = enc.farthest_point_sample(xs, 10)
    save("edge_fps", x=xd.numpy(), start=fd[:, 0].numpy(), fps=i32(fd), same_start=fs[:, 0].numpy(), same_fps=i32(fs))

    # ---- C2: one 120 000-point scan, SA1 + SA2 sampling/grouping (indices only) ----
    for name, cloud in (("lidar", S.lidar_scan(0)), ("uniform", S.uniform_cloud(0, 1, 120000))):
        torch.manual_seed(1234)
        f1 = enc.farthest_point_sample(cloud, 512)
        c1 = enc.index_points(cloud, f1)
        g1 = enc.query_ball_point(0.2, 32, cloud, c1)
        f2 = enc.farthest_point_sample(c1, 128)
        c2 = enc.index_points(c1, f2)
        g2 = enc.query_ball_point(0.4, 64, c1, c2)
        save("c2_120k_" + name, sha256=np.array(sha(cloud)), start1=f1[:, 0].numpy(), start2=f2[:, 0].numpy(),
             fps1=i32(f1), group1=i32(g1), fps2=i32(f2), group2=i32(g2))

    # ---- 3-NN inverse-distance upsample (models/diffusion_model.py:127-153) + sklearn metrics ----
    orig = S.uniform_cloud(11, 2, 6000)
    gi = torch.Generator().manual_seed(12)
    idx = torch.stack([torch.randperm(6000, generator=gi)[:1500] for _ in range(2)])
    coarse = torch.randn(2, 1500, 3, generator=gi)
    up = HP(6000, 1500).upsample_knn(coarse, orig, idx)
    pc, tc = S.uniform_cloud(13, 2, 3000), S.uniform_cloud(14, 2, 2500)
    save("upsample_knn", original=orig.numpy(), coarse=coarse.numpy(), indices=idx.numpy(), out=up.numpy(),
         pred=pc.numpy(), target=tc.numpy(), coverage_005=np.float64(M.coverage_score(pc, tc, 0.05)),
         coverage_001=np.float64(M.coverage_score(pc, tc, 0.01)), uniformity_8=np.float64(M.uniformity_score(pc, 8)),
         uniformity_4=np.float64(M.uniformity_score(tc, 4)))


if __name__ == "__main__":
    main()

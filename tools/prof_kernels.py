"""Runs each hot kernel a few times on the BASELINE shapes (for `ncu` captures and quick CUDA-event timings).

    python tools/prof_kernels.py [fps] [ball] [nn] [mlp] [enc] [knn]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointcloud_style_transfer_b200 import ops, synthetic as S  # noqa: E402
from pointcloud_style_transfer_b200.models.pointnet2_encoder import PointNet2Encoder  # noqa: E402

which = set(sys.argv[1:]) or {"fps", "ball", "nn", "mlp", "enc", "knn", "vox"}
dev = torch.device("cuda:0")
x = S.lidar_scan(0).to(dev)
y = S.lidar_scan(100).to(dev)
start = torch.tensor([1234], device=dev)


def timeit(name, fn, reps=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    print(f"{name}: {a.elapsed_time(b) / reps * 1e3:.1f} us", flush=True)


if "fpsprobe" in which:
    from pointcloud_style_transfer_b200 import _lib
    for mode, label in ((0, "full"), (3, "skip all chunks"), (4, "skip all, no block reduction"), (5, "skip all, no cluster exchange")):
        _lib.set_tuning("fps.prune", mode)
        timeit(f"fps 120k->512 probe [{label}]", lambda: ops.fps(x, 512, start), reps=10)
    _lib.set_tuning("fps.prune", 0)
if "fps2nd" in which:
    from pointcloud_style_transfer_b200 import _lib
    x512 = S.lidar_scan(0, 512).to(dev)
    _, c512 = ops.fps(x, 512, start)          # the real stage-2 input: the 512 centroids of a 120k scan
    for t in (32, 64, 128, 256, 512):
        _lib.set_tuning("fps.threads", t)
        timeit(f"fps 512->128 (stage-1 centroids) threads={t}", lambda: ops.fps(c512, 128, start * 0), reps=20)
    _lib.set_tuning("fps.threads", 0)
if "fps" in which:
    from pointcloud_style_transfer_b200 import _lib
    timeit("fps 120k->512 (lidar order)", lambda: ops.fps(x, 512, start))
    _lib.set_tuning("fps.prune", 2)
    timeit("fps 120k->512 (lidar order, no skip test)", lambda: ops.fps(x, 512, start))
    _lib.set_tuning("fps.prune", 3)
    timeit("fps 120k->512 (exchange chain only: every chunk skipped, results invalid)", lambda: ops.fps(x, 512, start))
    _lib.set_tuning("fps.prune", 4)
    timeit("fps 120k->512 (probe: chain without the block-level reduction)", lambda: ops.fps(x, 512, start))
    _lib.set_tuning("fps.prune", 5)
    timeit("fps 120k->512 (probe: chain without the cluster exchange)", lambda: ops.fps(x, 512, start))
    _lib.set_tuning("fps.prune", 0)
    xu = S.uniform_cloud(0, 1, 120000).to(dev)
    timeit("fps 120k->512 (uniform random order)", lambda: ops.fps(xu, 512, start))
    for c in (4, 8, 16):
        _lib.set_tuning("fps.cluster", c)
        x16 = S.uniform_cloud(0, 4, 16384).to(dev)
        timeit(f"fps 4x16384->512 cluster={c}", lambda: ops.fps(x16, 512, torch.zeros(4, dtype=torch.long, device=dev)))
    _lib.set_tuning("fps.cluster", 0)
    x4k = S.uniform_cloud(0, 2, 4096).to(dev)
    for c in (1, 2, 4):
        _lib.set_tuning("fps.cluster", c)
        timeit(f"fps 2x4096->512 cluster={c}", lambda: ops.fps(x4k, 512, torch.zeros(2, dtype=torch.long, device=dev)))
    _lib.set_tuning("fps.cluster", 0)
    for t in (32, 128, 512):
        _lib.set_tuning("fps.threads", t)
        x512 = S.uniform_cloud(0, 1, 512).to(dev)
        timeit(f"fps 512->128 threads={t}", lambda: ops.fps(x512, 128, start * 0))
    _lib.set_tuning("fps.threads", 0)
    x16 = S.uniform_cloud(0, 4, 16384).to(dev)
    st4 = torch.zeros(4, dtype=torch.long, device=dev)
    timeit("fps 4x16384->512", lambda: ops.fps(x16, 512, st4))
    x512 = S.uniform_cloud(0, 1, 512).to(dev)
    timeit("fps 512->128", lambda: ops.fps(x512, 128, start * 0))
if "ballbatch" in which:
    xb = torch.cat([S.lidar_scan(i, 16384) for i in range(32)], 0).to(dev)
    st = torch.zeros(32, dtype=torch.long, device=dev)
    _, c1 = ops.fps(xb, 512, st)
    timeit("ball_query 32 x (512 x 16384) r=0.2 ns=32", lambda: ops.ball_query(xb, c1, 0.04, 32))
    _, c2 = ops.fps(c1, 128, st)
    timeit("ball_query 32 x (128 x 512) r=0.4 ns=64 [bq_small]", lambda: ops.ball_query(c1, c2, 0.16, 64))
    timeit("ball_query 1 x (128 x 512) r=0.4 ns=64 [bq_small]", lambda: ops.ball_query(c1[:1].contiguous(), c2[:1].contiguous(), 0.16, 64))
    timeit("fps 32 x 16384 -> 512", lambda: ops.fps(xb, 512, st))
    timeit("fps 32 x 512 -> 128", lambda: ops.fps(c1, 128, st))
if "ball" in which:
    _, c1 = ops.fps(x, 512, start)
    timeit("ball_query 512x120k r=0.2 ns=32", lambda: ops.ball_query(x, c1, 0.04, 32))
if "nn" in which:
    timeit("nn_min 120k x 120k form0", lambda: ops.nn_min(x, y, 0, False), reps=3)
    timeit("nn_min 120k x 120k form0 +arg", lambda: ops.nn_min(x, y, 0, True), reps=3)
    timeit("nn_min 120k x 120k form1", lambda: ops.nn_min(x, y, 1, False), reps=3)
    timeit("nn_min_pair 120k x 120k form0 (both directions, one sweep)", lambda: ops.nn_min_pair(x, y, 0), reps=3)
    timeit("nn_min_pair 120k x 120k form1", lambda: ops.nn_min_pair(x, y, 1), reps=3)
    timeit("nn_min_pair_arg 120k x 120k (both directions + both argmins, one sweep)", lambda: ops.nn_min_pair_arg(x, y), reps=3)
if "enc" in which or "mlp" in which:
    torch.manual_seed(42)
    for prec in (0, 1):
        enc = PointNet2Encoder(feature_dim=256, mlp_precision=prec).eval().to(dev)
        with torch.no_grad():
            timeit(f"encoder eager precision={prec}", lambda: enc(x))
            ops.start_event_log()
            enc(x)
            log = ops.stop_event_log()
        print("   per-op us:", {k: [round(v * 1e3, 1) for v in vs] for k, vs in log.items()}, flush=True)
if "knn" in which:
    q, r = S.uniform_cloud(1, 1, 90000).to(dev), S.uniform_cloud(2, 1, 30000).to(dev)
    timeit("knn 90k x 30k k=3", lambda: ops.knn(q, r, 3), reps=2)
if "vox" in which:
    from pointcloud_style_transfer_b200.models.diffusion_model import HierarchicalProcessor
    hp = HierarchicalProcessor(120000, 30000)
    box = ops.minmax(x)
    vs = torch.tensor([0.05], device=dev)
    timeit("minmax 120k", lambda: ops.minmax(x))
    timeit("voxel_representatives 120k (hash + radix sort + run means)", lambda: ops.voxel_representatives(x, box[:, :3].contiguous(), vs))
    timeit("HierarchicalProcessor.downsample 120k -> 30k (incl. host scalar math, randperm, gathers)", lambda: hp.downsample(x), reps=3)
    q9 = S.lidar_scan(3)
    timeit("knn self 120k x 120k k=9 (uniformity_score)", lambda: ops.knn(x, x, 9), reps=2)

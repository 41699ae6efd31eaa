#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 point-set hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): SA points/s -- one "step" = one PointNet2Encoder forward (FPS -> ball query ->
grouping -> shared MLP + max, three set-abstraction stages) over one synthetic 120 000-point LiDAR scan
per GPU (BASELINE config[1]; N GPUs = N independent scans, weak scaling, no data-path collective).
`value` is device-resident throughput; `e2e` goes through the public drop-in API from pinned HOST
buffers (H2D of the scan and D2H of the feature inside the timed region).  The Chamfer NN pairs/s of
the same 120k scans is reported in `chamfer` with its own FP32-pipe roofline.  `--impl reference`
times the CPU oracle port of the reference (the reference itself is Python and cannot travel to
the GPU box) on the host cores for the same workload.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

N_POINTS = 120000
FEATURE_DIM = 256
NPOINT1 = 512
WORKLOAD = "PointNet2Encoder fwd (eval, F=256), 1 x 120000-pt synthetic LiDAR scan per GPU"


def peaks():
    try:
        with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", float(p.get("sm_max_mhz", 1965.0))
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


def bf16_peak_tflops():
    """Sustained dense bf16 tensor-core peak for kernels timed inside a step (MEASURED_PEAKS.json), else the
    profiling recipe's fallback."""
    try:
        with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p.get("bf16_tflops_sustained") or p["bf16_tflops"]), "measured, sustained (MEASURED_PEAKS.json)"
    except Exception:
        return 1400.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML.  Started before the warm-up (NVML initialisation takes
    longer than a short timed region); only samples taken between begin() and end() -- the timed region -- count."""

    def __init__(self, index: int, period: float = 0.002):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.max_mhz, self.error = [], None, None   # samples: (time, sm_mhz, reason bits)
        self.t0 = self.t1 = None
        self.ready = threading.Event()
        self._stop_evt = threading.Event()

    NAMES = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            self.ready.set()
            while not self._stop_evt.is_set():
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.samples.append((time.perf_counter(), nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), r))
                time.sleep(self.period)
        except Exception as e:  # NVML missing: report that instead of inventing numbers
            self.error = f"nvml_unavailable:{type(e).__name__}"
            self.ready.set()

    def begin(self):
        self.ready.wait(timeout=10)
        self.t0 = time.perf_counter()

    def end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        inside = [s for s in self.samples if self.t0 is not None and self.t0 <= s[0] <= (self.t1 or float("inf"))]
        reasons = set()
        for _, _, r in inside:
            reasons |= {name for bit, name in self.NAMES.items() if r & bit}
        if self.error:
            reasons.add(self.error)
        return {"sm_mhz": statistics.median(m for _, m, _ in inside) if inside else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(reasons), "samples": len(inside)}


def host_threads() -> int:
    """Threads the CPU baseline may use: the cores this process is allowed to run on."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_oracle(threads: int):
    """The CPU oracle port with ALL of its work inside one OpenMP runtime (distance kernels and the shared MLP in
    oracle/pcst_oracle.c): no second (BLAS / torch) thread pool competes for the cores, which is what slowed the
    round-1 baseline 2-3x.  An OMP_NUM_THREADS set by a launcher (torchrun exports 1) is overridden explicitly."""
    from oracle import ref_oracle as O

    O.build()
    O.set_num_threads(threads)
    return O


def cpu_encoder_baseline(budget_s: float, scan_seed: int = 0):
    """Time the CPU oracle port of the encoder forward on one 120k-point scan (host cores)."""
    import numpy as np
    import torch
    from pointcloud_style_transfer_b200 import synthetic as S
    from pointcloud_style_transfer_b200.models.pointnet2_encoder import PointNet2Encoder

    O = cpu_oracle(host_threads())
    torch.manual_seed(42)
    sd = {k: v.detach().numpy() for k, v in PointNet2Encoder(feature_dim=FEATURE_DIM).eval().state_dict().items()}
    x = S.lidar_scan(scan_seed).numpy()
    s1, s2 = np.array([1234], np.int64), np.array([99], np.int64)
    O.encoder_forward(x, sd, s1, s2, mlp=O.apply_mlp_c)  # warm-up
    times = []
    t_end = time.perf_counter() + budget_s
    while time.perf_counter() < t_end or len(times) < 3:
        t0 = time.perf_counter()
        O.encoder_forward(x, sd, s1, s2, mlp=O.apply_mlp_c)
        times.append(time.perf_counter() - t0)
    return times, O.num_threads()


def cpu_chamfer_knn_baselines():
    """The other half of the metric on the host cores, bounded samples:
    Chamfer = the oracle port of chamfer_distance_chunked_optimized on 30k x 30k points of the same two scans
    (120k x 120k would take ~16x longer: extrapolated pairs/s is the same, the sweep is compute-bound);
    3-NN = scikit-learn NearestNeighbors(k=3, kd-tree, fp64) on 90 000 queries x 30 000 references, the very call
    the reference makes at models/diffusion_model.py:146-147 (third-party library, present in this image)."""
    import numpy as np
    from pointcloud_style_transfer_b200 import synthetic as S

    O = cpu_oracle(host_threads())
    out = {}
    a, b = S.lidar_scan(0, 30000).numpy(), S.lidar_scan(100, 30000).numpy()
    O.chamfer_distance_chunked_optimized(a[:, :2000], b[:, :2000])
    t0 = time.perf_counter()
    O.chamfer_distance_chunked_optimized(a, b)
    dt = time.perf_counter() - t0
    out["chamfer"] = {"value": 2.0 * 30000 * 30000 / dt, "unit": "pairs/s", "seconds": dt, "cores": O.num_threads(),
                      "kind": "port", "sample": "one 30000 x 30000 call of the oracle port of chamfer_distance_chunked_optimized "
                      "(both directions); the 120k x 120k call is 16x this"}
    try:
        from sklearn.neighbors import NearestNeighbors
        x = S.lidar_scan(2).numpy()[0]
        perm = np.random.default_rng(0).permutation(x.shape[0])
        ref, q = x[np.sort(perm[:30000])], x[np.sort(perm[30000:])]
        t0 = time.perf_counter()
        NearestNeighbors(n_neighbors=3, algorithm="auto").fit(ref).kneighbors(q)
        dt = time.perf_counter() - t0
        out["knn3_90k_x_30k"] = {"value": q.shape[0] / dt, "unit": "queries/s", "seconds": dt, "cores": 1, "kind": "reference",
                                 "sample": "sklearn NearestNeighbors(n_neighbors=3).fit(30000 pts).kneighbors(90000 pts): the "
                                 "reference's own call (kd-tree, fp64, single-threaded as the reference leaves n_jobs unset)"}
    except Exception as e:  # sklearn absent: say so instead of inventing a number
        out["knn3_90k_x_30k"] = {"unavailable": f"{type(e).__name__}: {e}"}
    return out


def run_reference(args, rank):
    """--impl reference: the oracle port of the reference's CPU path, all host threads, same workload.
    Under torchrun only rank 0 works; the other ranks exit 0."""
    if rank != 0:
        return
    import numpy as np
    import torch
    from pointcloud_style_transfer_b200 import synthetic as S
    from pointcloud_style_transfer_b200.models.pointnet2_encoder import PointNet2Encoder

    O = cpu_oracle(host_threads())  # all host threads this process may use, one OpenMP runtime (see cpu_oracle)
    steps = max(1, args.steps)
    torch.manual_seed(42)
    sd = {k: v.detach().numpy() for k, v in PointNet2Encoder(feature_dim=FEATURE_DIM).eval().state_dict().items()}
    x = S.lidar_scan(0).numpy()
    s1, s2 = np.array([1234], np.int64), np.array([99], np.int64)
    for _ in range(max(1, args.warmup)):
        O.encoder_forward(x, sd, s1, s2, mlp=O.apply_mlp_c)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.encoder_forward(x, sd, s1, s2, mlp=O.apply_mlp_c)
    dt = (time.perf_counter() - t0) / steps
    value = N_POINTS / dt
    line = {
        "impl": "reference", "metric": "SA points/sec (PointNet2Encoder fwd, 120k-pt scan)", "value": value,
        "unit": "points/s", "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "points": N_POINTS, "feature_dim": FEATURE_DIM,
                   "note": "CPU oracle port of the reference's encoder path (oracle/pcst_oracle.c, OpenMP, %d threads, "
                           "no second thread pool); the Python reference cannot travel to the GPU box" % O.num_threads()},
        "cpu_baseline": {"value": value, "unit": "points/s", "cores": O.num_threads(), "kind": "port",
                         "sample": f"{steps} full encoder forwards of one 120k-point scan"},
        "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mlp-precision", type=int, default=int(os.environ.get("PCST_MLP_PRECISION", "1")),
                    help="1 = bf16 tcgen05 shared MLP (north_star's design, default); 0 = fp32 CUDA-core MLP")
    ap.add_argument("--chamfer-steps", type=int, default=5)
    ap.add_argument("--train-steps", type=int, default=10, help="timed steps of the config-4 training step (0 = skip)")
    ap.add_argument("--sampling-steps", type=int, default=10, help="DDIM steps timed for the sampling-loop line (0 = skip)")
    ap.add_argument("--batched-scans", type=int, default=-1,
                    help="scans per GPU of the secondary batched-throughput line (0 = skip, -1 = as many as FPS clusters fit at once)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    # Everything that writes to fd 1 from here on (NCCL's "NCCL version ..." banner, library chatter) goes to
    # stderr, so that stdout carries exactly ONE line: the JSON result, written through the saved descriptor.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import numpy as np
    import torch
    import torch.distributed as dist

    from pointcloud_style_transfer_b200 import ops, synthetic as S
    from pointcloud_style_transfer_b200.models.pointnet2_encoder import PointNet2Encoder
    from pointcloud_style_transfer_b200.runtime import GraphedEncoder

    assert torch.cuda.is_available(), "bench.py needs a CUDA device; there is no CPU fallback (use --impl reference)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    hbm_peak, peak_src, sm_max = peaks()
    W, K = max(3, args.warmup), max(1, args.steps)

    torch.manual_seed(42)
    enc = PointNet2Encoder(feature_dim=FEATURE_DIM, mlp_precision=args.mlp_precision).eval().to(dev)
    scan = S.lidar_scan(rank % 8)                      # one scan per GPU (seed = rank), [1,120000,3]
    x_dev = scan.to(dev)
    x_host = scan.pin_memory()
    genc = GraphedEncoder(enc)
    # the same weights on the other shared-MLP path (fp32 and bf16 reported separately)
    other_precision = 1 - int(bool(args.mlp_precision))
    enc_other = PointNet2Encoder(feature_dim=FEATURE_DIM, mlp_precision=other_precision).eval().to(dev)
    enc_other.load_state_dict(enc.state_dict())
    genc_other = GraphedEncoder(enc_other)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    # every rank samples its own GPU; with several ranks on one host poll less often (NVML calls share driver locks)
    sampler = ClockSampler(local_rank, period=0.002 if world == 1 else 0.01)
    sampler.start()
    torch.manual_seed(1234 + rank)
    for _ in range(W):
        genc(x_dev)
    feat_host = torch.empty(1, FEATURE_DIM, dtype=torch.float32).pin_memory()

    def timed(fn, steps):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        fn()                                           # untimed: a rank that waited at the barrier has an idle GPU
        for a, b in evs:
            flush.fill_(1)                             # L2 flush between timed iterations (untimed)
            a.record()
            fn()
            b.record()
        barrier()
        return [a.elapsed_time(b) for a, b in evs]

    launches0 = ops.launch_count
    sampler.begin()

    # ---- device-resident: input already in HBM, one graph replay per step ----
    ms_dev = timed(lambda: genc(x_dev), K)
    # graph replays do not pass through ops._call; count the kernels of one captured forward instead
    ops.launch_count = 0
    with torch.no_grad():
        enc(x_dev)
    launches_per_step = ops.launch_count

    # ---- end to end: pinned host scan -> H2D -> encoder -> D2H of the feature, per step ----
    def e2e_step():
        out = genc(x_host)
        feat_host.copy_(out, non_blocking=True)
    for _ in range(2):
        e2e_step()
    ms_e2e = timed(e2e_step, K)
    sampler.end()
    clocks = sampler.stop()

    # ---- the other MLP precision, same scan, same timing recipe ----
    for _ in range(W):
        genc_other(x_dev)
    ms_other = timed(lambda: genc_other(x_dev), K)

    # ---- per-op durations (eager, CUDA events around each C-ABI call), for the roofline ----
    torch.cuda.synchronize()
    per_op = {}
    for _ in range(W):
        with torch.no_grad():
            enc(x_dev)
    for _ in range(min(K, 20)):
        flush.fill_(1)
        ops.start_event_log()
        with torch.no_grad():
            enc(x_dev)
        for name, v in ops.stop_event_log().items():
            per_op.setdefault(name, []).append(v)
    fps_ms = statistics.mean(v[0] for v in per_op["pcst_fps_f32"])          # SA1 FPS (first call of the step)
    op_ms = {name: statistics.mean(sum(v) for v in vals) for name, vals in per_op.items()}

    # ---- Chamfer NN pairs/s on the same 120k scans (second half of the metric) ----
    from pointcloud_style_transfer_b200.models.losses import chamfer_distance_chunked_optimized
    y_dev = S.lidar_scan((rank % 8) + 100).to(dev)
    with torch.no_grad():
        for _ in range(2):
            chamfer_distance_chunked_optimized(x_dev, y_dev)
        ms_ch = timed(lambda: chamfer_distance_chunked_optimized(x_dev, y_dev), max(1, args.chamfer_steps))
    ch_ms = max(statistics.mean(ms_ch), 1e-9)

    # ---- kNN sweeps (config 5 anchor at N = 1, and the 3-NN of upsample_knn at its product shape) ----
    knn_ms = {}
    with torch.no_grad():
        xq = x_dev[:, :N_POINTS]
        for name, (q, r, k) in {"knn3_120k_x_120k": (x_dev, y_dev, 3), "knn9_self_120k": (x_dev, x_dev, 9)}.items():
            for _ in range(2):
                ops.knn(q, r, k)
            knn_ms[name] = statistics.mean(timed(lambda: ops.knn(q, r, k), max(1, args.chamfer_steps)))
        perm = torch.randperm(N_POINTS, generator=torch.Generator().manual_seed(0))
        r30 = x_dev[:, perm[:30000].sort().values.to(dev)].contiguous()
        q90 = x_dev[:, perm[30000:].sort().values.to(dev)].contiguous()
        for _ in range(2):
            ops.knn(q90, r30, 3)
        knn_ms["knn3_90k_x_30k"] = statistics.mean(timed(lambda: ops.knn(q90, r30, 3), max(1, args.chamfer_steps)))
        del xq

    # ---- FP32-pipe peak measured in the same run (packed FMA chains on every SM), for the Chamfer roofline ----
    from pointcloud_style_transfer_b200 import _lib
    import ctypes as _ct
    scratch = torch.zeros(1, device=dev)
    probe = lambda: _lib.load().pcst_fp32_probe(20000, _ct.c_void_p(scratch.data_ptr()),
                                                _ct.c_void_p(torch.cuda.current_stream().cuda_stream))
    probe()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    probe_flop = probe()
    ev1.record()
    torch.cuda.synchronize()
    fp32_measured = probe_flop / (ev0.elapsed_time(ev1) * 1e-3) / 1e12 if probe_flop else None

    # ---- batched throughput: 8 scans per GPU in one graph (FPS = 8 concurrent 16-CTA clusters, one per GPC) ----
    batched = None
    if args.batched_scans < 0:  # auto: as many scans as the FPS clusters that fit the GPU at once
        from pointcloud_style_transfer_b200 import _lib
        args.batched_scans = max(2, min(8, _lib.load().pcst_fps_max_concurrent_clouds(N_POINTS)))
    if args.batched_scans > 1:
        xb = torch.cat([S.lidar_scan((rank * args.batched_scans + i) % 16) for i in range(args.batched_scans)], 0).to(dev)
        for _ in range(3):
            genc(xb)
        ms_b = timed(lambda: genc(xb), max(3, K // 2))
        batched = statistics.mean(ms_b)
        del xb

    # ---- N > 1: ONE 120k x 120k Chamfer with the query points sharded over the ranks (BASELINE config 5):
    # all-gather of the target cloud + one sweep of the local [N/G x M] tile + MIN all-reduce of the column minima
    ch_sharded = knn_sharded = None
    if world > 1:
        from pointcloud_style_transfer_b200 import distributed as D
        p_full, t_full = S.lidar_scan(0), S.lidar_scan(100)
        lo, hi = D.slice_of_rank(N_POINTS, world, rank)
        p_loc, t_loc = p_full[:, lo:hi].contiguous().to(dev), t_full[:, lo:hi].contiguous().to(dev)
        with torch.no_grad():
            gch = D.GraphedShardedChamfer(N_POINTS, N_POINTS)       # sweep + both collectives in one CUDA graph
            for _ in range(2):
                cd_sh = gch(p_loc, t_loc).clone()
            ms_sh = timed(lambda: gch(p_loc, t_loc), max(3, args.chamfer_steps))
            for _ in range(2):
                D.chamfer_query_sharded_one_sweep(p_loc, t_loc, pred_total=N_POINTS, target_total=N_POINTS)
            ms_sh_r1 = timed(lambda: D.chamfer_query_sharded_one_sweep(p_loc, t_loc, pred_total=N_POINTS, target_total=N_POINTS),
                             max(3, args.chamfer_steps))
            cd_one = chamfer_distance_chunked_optimized(p_full.to(dev), t_full.to(dev))
            gch.release()   # the captured graphs hold NCCL kernels: gone before the process group is
        ch_sharded = (statistics.mean(ms_sh), float((cd_sh - cd_one).abs().max() / cd_one.abs().max()), statistics.mean(ms_sh_r1))
        # the kNN sweep of the same configuration: this rank's slice of the queries against ALL reference points
        with torch.no_grad():
            for _ in range(2):
                D.knn_query_sharded(p_loc, t_loc, 3)
            ms_knn = timed(lambda: D.knn_query_sharded(p_loc, t_loc, 3), max(1, args.chamfer_steps))
        knn_sharded = statistics.mean(ms_knn)

    # ---- BASELINE config 4: full diffusion training step, bf16, 4 scans x 16 384 points per GPU (32 x 16 384 on 8 GPUs):
    # style encoder forward + backward on the native train-mode kernels, denoiser on torch autocast, Chamfer branch on
    # (global_points lowered to 4096 as SURVEY.md 8(d) prescribes), one NCCL all-reduce of the flat gradient, AdamW.
    train_c4 = None
    if args.train_steps > 0:
        from pointcloud_style_transfer_b200.config import Config
        from pointcloud_style_transfer_b200.train_step import DiffusionTrainStep
        cfg = Config()
        cfg.total_points, cfg.global_points = 16384, 4096
        BL = 4
        torch.manual_seed(7)
        trainer = DiffusionTrainStep(cfg, dev, mlp_precision=1, world=world)
        sim = torch.cat([S.lidar_scan((rank * BL + i) % 16, 16384) for i in range(BL)], 0).to(dev)
        real = torch.cat([S.lidar_scan((rank * BL + i) % 16 + 100, 16384) for i in range(BL)], 0).to(dev)
        torch.manual_seed(11 + rank)
        torch.cuda.manual_seed(11 + rank)
        for _ in range(3):
            trainer.step(sim, real)
        l0 = ops.launch_count
        ms_tr = timed(lambda: trainer.step(sim, real), args.train_steps)
        train_launches = (ops.launch_count - l0) / (args.train_steps + 1)
        # the same step as ONE CUDA-graph replay (forward, backward, all-reduce, clip, AdamW, EMA)
        for _ in range(2):
            trainer.step_graphed(sim, real)
        ms_tr_graph = timed(lambda: trainer.step_graphed(sim, real), args.train_steps)
        train_c4 = (statistics.mean(ms_tr_graph), train_launches, BL, statistics.mean(ms_tr))
        # where the step's time goes (eager CUDA events around the C-ABI calls of one more step)
        ops.start_event_log()
        trainer.step(sim, real)
        tr_ops = {k: sum(v) for k, v in ops.stop_event_log().items()}
        trainer.release()   # the captured graph holds NCCL kernels: drop it before the process group goes away
        del trainer

    # ---- denoiser (NoisePredictor) on the coarse clouds of one CFG step: 2 x 30 000 points ----
    denoiser = sampling = None
    if args.sampling_steps > 0:
        from pointcloud_style_transfer_b200.config import Config
        from pointcloud_style_transfer_b200.models import diffusion_model as DM
        cfg = Config()
        torch.manual_seed(3)
        model = DM.PointCloudDiffusionModel(cfg, mlp_precision=1).to(dev).eval()
        proc = DM.DiffusionProcess(cfg, device=str(dev))
        xc = torch.randn(2, 30000, 3, device=dev)
        tt = torch.tensor([500, 500], device=dev)
        st = torch.randn(2, 256, device=dev)
        net = model.noise_predictor
        den = {}
        with torch.no_grad():
            for name in ("fused_tcgen05_bf16", "torch_linear_fp32", "torch_linear_bf16_autocast"):
                net.fused_inference = name == "fused_tcgen05_bf16"
                ctx = torch.autocast("cuda", dtype=torch.bfloat16, enabled=name.endswith("autocast"))
                with ctx:
                    for _ in range(3):
                        net(xc, tt, st)
                    den[name] = statistics.mean(timed(lambda: net(xc, tt, st), 10))
            net.fused_inference = True
        denoiser = den
        # ---- one CFG-guided DDIM step on a 120k scan: graph-replayed device-resident step vs the host-driven loop ----
        src, cnd = x_dev, y_dev
        K_s = args.sampling_steps
        with torch.no_grad():
            tg, te = {}, {}
            torch.cuda.manual_seed(5)
            proc.guided_sample_loop_device(model, src, cnd, num_inference_steps=2, graph=True)   # warm-up
            proc.guided_sample_loop_device(model, src, cnd, num_inference_steps=K_s, graph=True, timing=tg)
            proc.guided_sample_loop_device(model, src, cnd, num_inference_steps=K_s, graph=False, timing=te)
            torch.manual_seed(5)
            proc.guided_sample_loop(model, src, cnd, num_inference_steps=1)                        # warm-up
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            proc.guided_sample_loop(model, src, cnd, num_inference_steps=3)
            torch.cuda.synchronize()
            host_ms = (time.perf_counter() - t0) * 1e3 / 3
        sampling = {"steps": K_s, "graph_replay_ms_per_step": tg["ms_per_step"], "device_eager_ms_per_step": te["ms_per_step"],
                    "host_loop_ms_per_step": host_ms}
        del model

    def allmax(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allgather(v):
        if world == 1:
            return [v]
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        out = torch.empty(world, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(out, t)
        return [float(x) for x in out.tolist()]

    rank_ms = allgather(sum(ms_dev) / K)     # per-rank view of the same timed steps (diagnosis of a slow rank)
    rank_worst = allgather(max(ms_dev))
    rank_mhz = allgather(float(clocks["sm_mhz"]) if clocks["sm_mhz"] is not None else -1.0)
    rank_throttled = allgather(float(len([r for r in clocks["reasons"] if not r.startswith("nvml_unavailable")])))
    rank_median = allgather(statistics.median(ms_dev))
    t_other = allmax(sum(ms_other)) / K
    t_dev = allmax(sum(ms_dev)) / K          # ms per step, max over ranks
    t_e2e = allmax(sum(ms_e2e)) / K
    t_ch = allmax(ch_ms)
    fps_ms_max = allmax(fps_ms)
    t_b = allmax(batched) if batched is not None else None
    t_sh = allmax(ch_sharded[0]) if ch_sharded is not None else None
    t_sh_r1 = allmax(ch_sharded[2]) if ch_sharded is not None else None
    t_knn = allmax(knn_sharded) if knn_sharded is not None else None
    t_train = allmax(train_c4[0]) if train_c4 is not None else None
    t_train_host = allmax(train_c4[3]) if train_c4 is not None else None

    if rank == 0:
        value = world * N_POINTS / (t_dev * 1e-3)
        e2e_value = world * N_POINTS / (t_e2e * 1e-3)
        fps_bytes = NPOINT1 * N_POINTS * 20.0           # streaming-model bytes, SURVEY.md §8(d)
        fps_gbs = fps_bytes / (fps_ms_max * 1e-3) / 1e9
        pairs = 2.0 * N_POINTS * N_POINTS               # both directions, as the reference evaluates them
        fp32_nominal = 148 * 128 * 2 * sm_max * 1e6 / 1e12  # TFLOP/s, FP32 CUDA cores at the max SM clock
        fp32_peak = fp32_measured or fp32_nominal
        # one sweep serves both directions (the second matrix is the exact transpose): N*M unique pair evaluations
        ch_tflops = 8.0 * (pairs / 2) / (t_ch * 1e-3) / 1e12
        # ---- per-kernel rooflines from the eager CUDA-event durations of this run (SURVEY.md 8(d) work per unit) ----
        def call_ms(name, i):
            return statistics.mean(v[i] for v in per_op[name] if len(v) > i)

        bf16_peak, bf16_src = bf16_peak_tflops()
        mlp_rows = [NPOINT1 * 32, 128 * 64, 128]
        mlp_macs = [3 * 64 + 64 * 64 + 64 * 128, 131 * 128 + 128 * 128 + 128 * 256, 259 * 256 + 256 * 512 + 512 * FEATURE_DIM]
        tc = args.mlp_precision == 1
        kernels = []
        for i, stage in enumerate(("SA1", "SA2", "SA3 (group_all)")):
            ms = call_ms("pcst_sa_mlp_max_f32", i)
            tf = 2.0 * mlp_rows[i] * mlp_macs[i] / (ms * 1e-3) / 1e12
            peak = bf16_peak if tc else fp32_peak
            kernels.append({"kernel": ("sa_mlp_tc_kernel " if tc else "mlp_layer_kernel ") + stage, "bound": "tensor" if tc else "fp32",
                            "ms": ms, "achieved": tf, "peak": peak, "unit": "TFLOP/s", "frac": tf / peak,
                            "work": "2 * rows * sum(Cin * Cout) = %.3f GFLOP (%d rows)" % (2e-9 * mlp_rows[i] * mlp_macs[i], mlp_rows[i]),
                            "note": "one scan is a few microseconds of tensor work: latency-bound at this shape; "
                                    "the batched shape is under `train_c4` / profiles"})
        bq_ms = call_ms("pcst_ball_query_f32", 1)
        bq_tf = 8.0 * NPOINT1 * N_POINTS / (bq_ms * 1e-3) / 1e12
        kernels.append({"kernel": "bq_mask_kernel + bq_emit_kernel (SA1 ball query, 512 x 120000 full sweep)", "bound": "fp32",
                        "ms": bq_ms, "achieved": bq_tf, "peak": fp32_peak, "unit": "TFLOP/s", "frac": bq_tf / fp32_peak,
                        "work": "8 flop per (query, candidate) pair of the full S x N sweep = 0.49 GFLOP"})
        kernels.append({"kernel": "fps_kernel (SA2, 512 -> 128, one CTA)", "bound": "latency", "ms": call_ms("pcst_fps_f32", 1),
                        "ns_per_iteration": call_ms("pcst_fps_f32", 1) * 1e6 / 128, "note": "on a parallel graph branch"})
        kernels.append({"kernel": "bq_small_kernel (SA2 ball query, 128 x 512)", "bound": "latency", "ms": call_ms("pcst_ball_query_f32", 0)})
        for name, (nq, nr, kk) in {"knn3_90k_x_30k": (90000, 30000, 3), "knn3_120k_x_120k": (N_POINTS, N_POINTS, 3),
                                   "knn9_self_120k": (N_POINTS, N_POINTS, 9)}.items():
            tf = 8.0 * nq * nr / (knn_ms[name] * 1e-3) / 1e12
            kernels.append({"kernel": "knn (%s, k=%d, fp64-exact ranking)" % (name, kk), "bound": "fp32", "ms": knn_ms[name],
                            "achieved": tf, "peak": fp32_peak, "unit": "TFLOP/s", "frac": tf / fp32_peak,
                            "work": "8 flop per pair of the brute-force sweep (%d x %d); a search that visits fewer pairs can exceed 1" % (nq, nr)})

        line = {
            "metric": "SA points/sec (PointNet2Encoder fwd, 120k-pt scan)", "value": value, "unit": "points/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": t_dev, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32" if args.mlp_precision == 0 else "f32 distances/indices, bf16 tcgen05 MLP (f32 accumulate)",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "points": N_POINTS, "feature_dim": FEATURE_DIM,
                       "scans_per_gpu": 1, "l2": "flushed between timed iterations (256 MiB write)",
                       "timing": "CUDA events around each of the K steps, summed; one untimed step after the opening barrier",
                       "launch": "one CUDA-graph replay per step (stage-2 sampling on a parallel graph branch)",
                       "mlp_precision": args.mlp_precision},
            "e2e": {"value": e2e_value, "unit": "points/s", "ms_per_step": t_e2e,
                    "h2d_bytes_per_step": int(scan.numel() * 4 + 16), "d2h_bytes_per_step": FEATURE_DIM * 4},
            "gpu_launches": launches_per_step * K * 2,
            "roofline": {"kernel": "fps_kernel<16> (SA1 farthest point sampling, 16-CTA cluster)", "bound": "hbm",
                         "achieved": fps_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": fps_gbs / hbm_peak,
                         "traffic": 1.46e6, "traffic_source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per "
                         "launch (profiles/r01_ncu_full_kernels.md); = the cloud read once, the kernel keeps it on chip",
                         "peak_source": peak_src, "kernel_ms": fps_ms_max,
                         "model": "streaming-model bytes npoint*N*20 B per launch (SURVEY.md 8(d)): what an FPS that re-reads xyz and "
                                  "running distances every iteration moves; the kernel is bound by its 512 serial cluster exchanges"},
            "mlp_precisions": {("bf16_tcgen05" if args.mlp_precision == 1 else "fp32_cuda_core"): {"ms_per_step": t_dev, "value": value},
                               ("fp32_cuda_core" if args.mlp_precision == 1 else "bf16_tcgen05"):
                                   {"ms_per_step": t_other, "value": world * N_POINTS / (t_other * 1e-3)},
                               "note": "same scan, weights and timing recipe; the headline `value` is the first entry"},
            "fps": {"ns_per_iteration": fps_ms_max * 1e6 / NPOINT1, "iterations": NPOINT1,
                    "note": "the honest figure for this latency-bound kernel: one iteration = one 16-CTA cluster exchange"},
            "kernels": kernels,
            "kernels_note": "eager CUDA-event duration around each C-ABI call (includes launch gaps for microsecond-scale kernels); "
                            "ncu per-launch durations of the same kernels are under profiles/",
            "op_ms_eager": op_ms,
            "chamfer": {"metric": "Chamfer NN pairs/sec (120k x 120k, both directions)", "value": world * pairs / (t_ch * 1e-3),
                        "unit": "pairs/s", "ms_per_call": t_ch,
                        "roofline": {"kernel": "nn_min_pair_kernel<0>", "bound": "fp32", "achieved": ch_tflops, "peak": fp32_peak,
                                     "unit": "TFLOP/s", "frac": ch_tflops / fp32_peak,
                                     "model": "8 flop per UNIQUE pair evaluation (N*M: one sweep yields both directions' minima); "
                                              "peak = packed-FMA micro-benchmark timed in this run (pcst_fp32_probe); nominal "
                                              "148 SMs x 128 lanes x 2 x sm_max_mhz = %.1f" % fp32_nominal,
                                     "peak_source": "measured in this run" if fp32_measured else "computed",
                                     "frac_if_both_directions_counted": 2 * ch_tflops / fp32_peak}},
            "clocks": clocks,
            "ranks": {"ms_per_step": rank_ms, "median_step_ms": rank_median, "slowest_single_step_ms": rank_worst, "sm_mhz": rank_mhz,
                      "throttle_reasons_seen": rank_throttled},
        }
        if t_train is not None:
            BLc = train_c4[2]
            line["train_c4"] = {
                "metric": "training points/sec, full diffusion training step (BASELINE config 4), bf16, %d x 16384 points per GPU "
                          "(%d x 16384 on %d GPU%s)" % (BLc, BLc * world, world, "" if world == 1 else "s"),
                "value": world * BLc * 16384 / (t_train * 1e-3), "unit": "points/s", "ms_per_step": t_train, "scaling": "weak",
                "global_batch": BLc * world, "points_per_scan": 16384, "global_points": 4096,
                "launch": "DiffusionTrainStep.step_graphed: one CUDA-graph replay per step (forward, backward, all-reduce, clip, "
                          "AdamW, EMA); FPS start draws on the CPU generator before the replay",
                "host_loop_ms_per_step": t_train_host,
                "host_loop_note": "DiffusionTrainStep.step: the reference's host behaviour (kernels launched one by one, the loss "
                                  "dict's three .item() syncs per step)",
                "kernel_launches_per_step": train_c4[1], "op_ms_eager": tr_ops,
                "what": "q_sample -> voxel downsample (device) -> style encoder fwd (native train-mode tcgen05 kernels, batch-stat "
                        "BatchNorm) -> denoiser (torch autocast bf16) -> L1 + 0.1 * Chamfer (native) -> backward (native dgrad / "
                        "wgrad / BatchNorm / Chamfer backward) -> one NCCL all-reduce of the flat gradient (10.2 MB) -> clip -> "
                        "fused AdamW -> EMA",
                "dtype": "bf16 GEMM operands (encoder: tcgen05 train kernels; denoiser: autocast), fp32 distances / statistics"}
        if denoiser is not None:
            rows_d = 2 * 30000
            fl = 2.0 * rows_d * (3 * 128 + 128 * 256 + 256 * 256 + 6 * (256 * 512 * 2) + 256 * 256 + 256 * 128 + 128 * 3)
            bf16_peak2, _ = bf16_peak_tflops()
            line["denoiser"] = {"metric": "NoisePredictor forward, 2 x 30000 coarse points (one CFG step), ms",
                                "ms": denoiser, "gflop": fl / 1e9,
                                "roofline": {"kernel": "noise_mlp_kernel", "bound": "tensor", "unit": "TFLOP/s",
                                             "achieved": fl / (denoiser["fused_tcgen05_bf16"] * 1e-3) / 1e12, "peak": bf16_peak2,
                                             "frac": fl / (denoiser["fused_tcgen05_bf16"] * 1e-3) / 1e12 / bf16_peak2}}
        if sampling is not None:
            line["sampling_loop"] = dict(sampling, metric="one CFG-guided DDIM step on a 120k-point scan (downsample x2 -> denoiser -> "
                                         "3-NN upsample x2 -> update), ms per step",
                                         note="graph = guided_sample_loop_device (one CUDA graph replay per step); host = "
                                              "guided_sample_loop (the reference's loop structure with host-side RNG draws)")
        if t_b is not None:
            line["batched"] = {"metric": "SA points/sec, %d x 120k-pt scans per GPU in one graph (= concurrent 16-CTA FPS clusters)" % args.batched_scans,
                               "value": world * args.batched_scans * N_POINTS / (t_b * 1e-3), "unit": "points/s",
                               "ms_per_step": t_b, "scans_per_gpu": args.batched_scans}
        if t_sh is not None:
            line["chamfer_query_sharded"] = {
                "metric": "Chamfer NN pairs/sec, ONE 120k x 120k pair, query points sharded over %d GPUs" % world,
                "value": pairs / (t_sh * 1e-3), "unit": "pairs/s", "ms_per_call": t_sh, "scaling": "strong",
                "collectives": "all-gather of the target cloud (1.44 MB), then ONE all-gather of the packed payloads (column minima + "
                               "fp64 row sum per rank); min over ranks and the means in a finish kernel; the whole call (sweep + both "
                               "collectives) is one CUDA graph launch",
                "rel_diff_vs_single_gpu": ch_sharded[1],
                "ms_per_call_round1_variant": t_sh_r1,
                "round1_variant": "all-gather + sweep + all_reduce(MIN) + all_reduce(SUM) + torch reductions, launched eagerly"}
            line["knn_query_sharded"] = {
                "metric": "3-NN search, 120k queries x 120k references, queries sharded over %d GPUs (fp64-exact ranking)" % world,
                "value": N_POINTS * float(N_POINTS) / (t_knn * 1e-3), "unit": "pairs/s", "ms_per_call": t_knn,
                "scaling": "strong", "collectives": "all-gather of the reference cloud (1.44 MB); results stay sharded"}
        if not args.no_cpu_baseline and world == 1:
            times, cores = cpu_encoder_baseline(budget_s=8.0)
            line["cpu_baseline"] = {"value": N_POINTS / statistics.mean(times), "unit": "points/s", "cores": cores,
                                    "kind": "port", "sample": f"{len(times)} encoder forwards of the same 120k-point scan "
                                    "by the CPU oracle (C, one OpenMP runtime, %d threads), ~8 s" % cores,
                                    "best_ms": min(times) * 1e3, "mean_ms": statistics.mean(times) * 1e3}
            other = cpu_chamfer_knn_baselines()
            line["chamfer"]["cpu_baseline"] = other["chamfer"]
            line["knn_cpu_baseline"] = other["knn3_90k_x_30k"]
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

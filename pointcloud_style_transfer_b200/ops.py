"""torch custom ops ``pcst::*`` over the C ABI of libpcst.so (``include/pcst.h``).

Each op checks its inputs (CUDA, dtype, layout), allocates outputs and workspace with torch, and
launches the kernel on torch's current CUDA stream through ctypes.  PyTorch is plumbing here
(device memory, streams, autograd graph); all arithmetic happens in the CUDA kernels.  There is
no CPU path: CPU tensors raise.
"""

import ctypes
from typing import List, Optional, Tuple

import torch
from torch import Tensor

from . import _lib
from ._lib import Mlp3, check


def _stream() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t: Optional[Tensor]) -> ctypes.c_void_p:
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _need_cuda(*ts: Tensor) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("pointcloud_style_transfer_b200 ops run on CUDA (B200) tensors only; "
                               "there is no CPU fallback")


def _f32c(t: Tensor) -> Tensor:
    """fp32 + contiguous (also opts the distance kernels out of autocast)."""
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _i64c(t: Tensor) -> Tensor:
    if t.dtype != torch.int64:
        t = t.long()
    return t.contiguous()


def _workspace(nbytes: int, device) -> Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


# Launch accounting and optional per-op CUDA-event timing (used by bench.py for `gpu_launches` and the
# roofline's live kernel durations).  KERNELS_PER_CALL counts __global__ launches (memsets excluded).
KERNELS_PER_CALL = {"pcst_l2_prefetch": 1, "pcst_fps_f32": 1, "pcst_ball_query_f32": 2, "pcst_square_distance_f32": 1,
                    "pcst_index_points_f32": 1, "pcst_index_points_bwd_f32": 1, "pcst_group_f32": 1,
                    "pcst_sa_mlp_max_f32": 3, "pcst_sa_mlp_pack_f32": 8, "pcst_nn_min_f32": 4, "pcst_nn_min_pair_f32": 4, "pcst_nn_min_pair_arg_f32": 6, "pcst_chamfer_bwd_f32": 1, "pcst_knn_f32": 1,
                    "pcst_knn_interpolate_f32": 1, "pcst_minmax_f32": 1, "pcst_voxel_representatives_f32": 12}  # hash, 4 x (count, scatter), heads, runs, finish
launch_count = 0
_event_log = None  # None = off; else list of (name, start_event, end_event)


def start_event_log() -> None:
    global _event_log
    _event_log = []


def stop_event_log():
    """-> {c_symbol: [ms, ...]} (synchronises the device)."""
    global _event_log
    log, _event_log = _event_log or [], None
    torch.cuda.synchronize()
    out = {}
    for name, a, b in log:
        out.setdefault(name, []).append(a.elapsed_time(b))
    return out


def _call(name: str, *args, kernels: Optional[int] = None) -> None:
    global launch_count
    fn = getattr(_lib.load(), name)
    if _event_log is None:
        check(fn(*args))
    else:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        check(fn(*args))
        b.record()
        _event_log.append((name, a, b))
    launch_count += KERNELS_PER_CALL[name] if kernels is None else kernels


def l2_prefetch(t: Tensor) -> None:
    """Stream-ordered L2 warm-up of a (contiguous) CUDA tensor's bytes; a scheduling aid, no result."""
    _need_cuda(t)
    with torch.cuda.device(t.device):
        _call("pcst_l2_prefetch", _p(t), t.numel() * t.element_size(), _stream())


# ------------------------------------------------------------------------------------------ FPS


@torch.library.custom_op("pcst::fps", mutates_args=(), device_types="cuda")
def fps(xyz: Tensor, npoint: int, start: Tensor) -> Tuple[Tensor, Tensor]:
    """xyz [B,N,3] fp32, start [B] int64 -> (idx [B,npoint] int64, new_xyz [B,npoint,3])."""
    lib = _lib.load()
    _need_cuda(xyz, start)
    xyz, start = _f32c(xyz), _i64c(start)
    B, N, _ = xyz.shape
    idx = torch.empty(B, npoint, dtype=torch.int64, device=xyz.device)
    new_xyz = torch.empty(B, npoint, 3, dtype=torch.float32, device=xyz.device)
    with torch.cuda.device(xyz.device):
        nb = lib.pcst_fps_workspace_bytes(B, N, npoint)
        ws = _workspace(nb, xyz.device)
        _call("pcst_fps_f32", _p(xyz), B, N, npoint, _p(start), _p(idx), _p(new_xyz), _p(ws), ws.numel(), _stream())
    return idx, new_xyz


@fps.register_fake
def _(xyz, npoint, start):
    B = xyz.shape[0]
    return xyz.new_empty(B, npoint, dtype=torch.int64), xyz.new_empty(B, npoint, 3, dtype=torch.float32)


# ----------------------------------------------------------------------------------- ball query


@torch.library.custom_op("pcst::ball_query", mutates_args=(), device_types="cuda")
def ball_query(xyz: Tensor, new_xyz: Tensor, radius_sq: float, nsample: int) -> Tensor:
    """xyz [B,N,3], new_xyz [B,S,3], radius_sq = fp32(radius**2) -> idx [B,S,nsample] int64."""
    lib = _lib.load()
    _need_cuda(xyz, new_xyz)
    xyz, new_xyz = _f32c(xyz), _f32c(new_xyz)
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    out = torch.empty(B, S, nsample, dtype=torch.int64, device=xyz.device)
    with torch.cuda.device(xyz.device):
        ws = _workspace(lib.pcst_ball_query_workspace_bytes(B, N, S), xyz.device)
        _call("pcst_ball_query_f32", _p(xyz), _p(new_xyz), B, N, S, ctypes.c_float(radius_sq), nsample, _p(out),
              _p(ws), ws.numel(), _stream(), kernels=1 if N <= 2048 else 2)
    return out


@ball_query.register_fake
def _(xyz, new_xyz, radius_sq, nsample):
    return xyz.new_empty(xyz.shape[0], new_xyz.shape[1], nsample, dtype=torch.int64)


@torch.library.custom_op("pcst::square_distance", mutates_args=(), device_types="cuda")
def square_distance(src: Tensor, dst: Tensor) -> Tensor:
    lib = _lib.load()
    _need_cuda(src, dst)
    src, dst = _f32c(src), _f32c(dst)
    B, N, _ = src.shape
    M = dst.shape[1]
    out = torch.empty(B, N, M, dtype=torch.float32, device=src.device)
    with torch.cuda.device(src.device):
        _call("pcst_square_distance_f32", _p(src), _p(dst), B, N, M, _p(out), _stream())
    return out


@square_distance.register_fake
def _(src, dst):
    return src.new_empty(src.shape[0], src.shape[1], dst.shape[1], dtype=torch.float32)


# ---------------------------------------------------------------------------------- index_points


@torch.library.custom_op("pcst::index_points", mutates_args=(), device_types="cuda")
def index_points(points: Tensor, idx: Tensor) -> Tensor:
    """points [B,N,C] fp32, idx [B,...] int64 (clamped) -> [B,...,C]."""
    lib = _lib.load()
    _need_cuda(points, idx)
    points, idx = _f32c(points), _i64c(idx)
    B, N, C = points.shape
    S = idx.numel() // B
    out = torch.empty(*idx.shape, C, dtype=torch.float32, device=points.device)
    with torch.cuda.device(points.device):
        _call("pcst_index_points_f32", _p(points), _p(idx), B, N, C, S, _p(out), _stream())
    return out


@index_points.register_fake
def _(points, idx):
    return points.new_empty(*idx.shape, points.shape[2], dtype=torch.float32)


@torch.library.custom_op("pcst::index_points_bwd", mutates_args=(), device_types="cuda")
def index_points_bwd(grad_out: Tensor, idx: Tensor, N: int) -> Tensor:
    lib = _lib.load()
    _need_cuda(grad_out, idx)
    grad_out, idx = _f32c(grad_out), _i64c(idx)
    B = idx.shape[0]
    C = grad_out.shape[-1]
    S = idx.numel() // B
    grad_points = torch.zeros(B, N, C, dtype=torch.float32, device=grad_out.device)
    with torch.cuda.device(grad_out.device):
        _call("pcst_index_points_bwd_f32", _p(grad_out), _p(idx), B, N, C, S, _p(grad_points), _stream())
    return grad_points


@index_points_bwd.register_fake
def _(grad_out, idx, N):
    return grad_out.new_empty(idx.shape[0], N, grad_out.shape[-1], dtype=torch.float32)


def _index_points_setup(ctx, inputs, output):
    points, idx = inputs
    ctx.save_for_backward(idx)
    ctx.N = points.shape[1]


def _index_points_backward(ctx, grad):
    (idx,) = ctx.saved_tensors
    return index_points_bwd(grad, idx, ctx.N), None


index_points.register_autograd(_index_points_backward, setup_context=_index_points_setup)


# ----------------------------------------------------------------------------------------- group


@torch.library.custom_op("pcst::group", mutates_args=(), device_types="cuda")
def group(xyz: Tensor, feats: Optional[Tensor], new_xyz: Tensor, idx: Tensor) -> Tensor:
    """cat([xyz[idx] - new_xyz[:, :, None], feats[idx]], -1) -> [B,S,K,3+D]."""
    lib = _lib.load()
    _need_cuda(xyz, feats, new_xyz, idx)
    xyz, new_xyz, idx = _f32c(xyz), _f32c(new_xyz), _i64c(idx)
    feats = None if feats is None else _f32c(feats)
    B, N, _ = xyz.shape
    _, S, K = idx.shape
    D = 0 if feats is None else feats.shape[2]
    out = torch.empty(B, S, K, 3 + D, dtype=torch.float32, device=xyz.device)
    with torch.cuda.device(xyz.device):
        _call("pcst_group_f32", _p(xyz), _p(feats), _p(new_xyz), _p(idx), B, N, S, K, D, _p(out), _stream())
    return out


@group.register_fake
def _(xyz, feats, new_xyz, idx):
    D = 0 if feats is None else feats.shape[2]
    return xyz.new_empty(*idx.shape, 3 + D, dtype=torch.float32)


def _group_setup(ctx, inputs, output):
    xyz, feats, new_xyz, idx = inputs
    ctx.save_for_backward(idx)
    ctx.N = xyz.shape[1]
    ctx.has_feats = feats is not None


def _group_backward(ctx, grad):
    (idx,) = ctx.saved_tensors
    g_xyz = index_points_bwd(grad[..., :3].contiguous(), idx, ctx.N)
    g_new = -grad[..., :3].sum(dim=2)
    g_feats = index_points_bwd(grad[..., 3:].contiguous(), idx, ctx.N) if ctx.has_feats else None
    return g_xyz, g_feats, g_new, None


group.register_autograd(_group_backward, setup_context=_group_setup)


# ---------------------------------------------------------------------------- SA MLP + max-pool


def _cout3(couts: List[int]):
    if len(couts) != 3:
        raise ValueError("sa_mlp: the kernels implement the reference's 3-layer shared MLP")
    return (ctypes.c_int * 3)(*[int(c) for c in couts])


@torch.library.custom_op("pcst::sa_mlp_pack", mutates_args=(), device_types="cuda")
def sa_mlp_pack(weights: List[Tensor], scales: List[Tensor], shifts: List[Tensor], D: int, precision: int,
                cluster: int) -> Tensor:
    """Pack the folded parameters of a 3-layer shared MLP once (bf16 UMMA operand blocks for the tensor-core
    path, split into ``cluster`` N slices; a plain fp32 blob otherwise) -> uint8 tensor, reused by every
    ``sa_mlp_max`` call with the same (precision, cluster)."""
    lib = _lib.load()
    _need_cuda(*weights, *scales, *shifts)
    if len(weights) != 3 or len(scales) != 3 or len(shifts) != 3:
        raise ValueError("sa_mlp_pack: the kernels implement the reference's 3-layer shared MLP")
    weights = [_f32c(w.reshape(w.shape[0], -1)) for w in weights]
    scales = [_f32c(s) for s in scales]
    shifts = [_f32c(s) for s in shifts]
    m = Mlp3()
    cin = 3 + D
    for l in range(3):
        if weights[l].shape[1] != cin:
            raise ValueError(f"sa_mlp_pack: layer {l} expects Cin={weights[l].shape[1]}, got {cin}")
        m.w[l], m.scale[l], m.shift[l] = weights[l].data_ptr(), scales[l].data_ptr(), shifts[l].data_ptr()
        m.cout[l] = cin = weights[l].shape[0]
    dev = weights[0].device
    with torch.cuda.device(dev):
        nb = lib.pcst_sa_mlp_packed_bytes(D, m.cout, precision, cluster)
        if nb == 0:
            raise ValueError("sa_mlp_pack: unsupported layer widths (Cout must be a multiple of 32, <= 1024) or cluster size")
        packed = torch.empty(nb, dtype=torch.uint8, device=dev)
        _call("pcst_sa_mlp_pack_f32", ctypes.byref(m), D, precision, cluster, _p(packed), nb, _stream())
    return packed


@sa_mlp_pack.register_fake
def _(weights, scales, shifts, D, precision, cluster):
    return weights[0].new_empty(1, dtype=torch.uint8)


def sa_mlp_pick_cluster(B: int, S: int, K: int, D: int, couts: List[int], precision: int) -> int:
    """CTAs per 128-row tile the tensor-core kernel should use for this shape (1 on the fp32 path)."""
    return int(_lib.load().pcst_sa_mlp_pick_cluster(B, S, K, D, _cout3(couts), precision))


@torch.library.custom_op("pcst::sa_mlp_max", mutates_args=(), device_types="cuda")
def sa_mlp_max(xyz: Tensor, feats: Optional[Tensor], new_xyz: Optional[Tensor], idx: Optional[Tensor],
               packed: Tensor, couts: List[int], precision: int, cluster: int) -> Tensor:
    """Fused grouping gather + 3 x relu(scale * (W x) + shift) + max over each group -> [B,S,Cout] (POINT-major;
    the reference's channel-first layout is ``.permute(0, 2, 1)``).

    ``idx is None`` = group_all (one group holding the whole cloud, no centroid subtraction);
    ``packed`` comes from ``sa_mlp_pack`` with the same ``couts`` / ``precision`` / ``cluster`` / D."""
    lib = _lib.load()
    _need_cuda(xyz, feats, new_xyz, idx, packed)
    xyz = _f32c(xyz)
    feats = None if feats is None else _f32c(feats)
    new_xyz = None if new_xyz is None else _f32c(new_xyz)
    idx = None if idx is None else _i64c(idx)
    B, N, _ = xyz.shape
    D = 0 if feats is None else feats.shape[2]
    if idx is None:
        S, K = 1, N
    else:
        _, S, K = idx.shape
    c3 = _cout3(couts)
    out = torch.empty(B, S, int(couts[2]), dtype=torch.float32, device=xyz.device)
    with torch.cuda.device(xyz.device):
        if packed.numel() < lib.pcst_sa_mlp_packed_bytes(D, c3, precision, cluster):
            raise ValueError("sa_mlp_max: `packed` does not match (D, couts, precision, cluster)")
        nb = lib.pcst_sa_mlp_max_workspace_bytes(B, N, S, K, D, c3, precision)
        ws = _workspace(nb, xyz.device)
        _call("pcst_sa_mlp_max_f32", _p(xyz), _p(feats), _p(new_xyz), _p(idx), B, N, S, K, D, c3, precision, cluster,
              _p(packed), _p(out), _p(ws), ws.numel(), _stream(),
              kernels=int(lib.pcst_sa_mlp_max_kernel_launches(B, N, S, K, D, c3, precision)))
    return out


@sa_mlp_max.register_fake
def _(xyz, feats, new_xyz, idx, packed, couts, precision, cluster):
    S = 1 if idx is None else idx.shape[1]
    return xyz.new_empty(xyz.shape[0], S, couts[2], dtype=torch.float32)


# --------------------------------------------------------------------------------------- NN-min


@torch.library.custom_op("pcst::nn_min", mutates_args=(), device_types="cuda")
def nn_min(a: Tensor, b: Tensor, form: int, want_arg: bool) -> Tuple[Tensor, Tensor]:
    """Row minima (and argmins if ``want_arg``) of the pair matrix between a [B,N,3] and b [B,M,3].
    form 0 = loss (clamped squared, expanded), 1 = cdist rows (a = x1), 2 = cdist columns (a = x2)."""
    lib = _lib.load()
    _need_cuda(a, b)
    a, b = _f32c(a), _f32c(b)
    B, N, _ = a.shape
    M = b.shape[1]
    rowmin = torch.empty(B, N, dtype=torch.float32, device=a.device)
    rowarg = torch.empty((B, N) if want_arg else (0,), dtype=torch.int64, device=a.device)
    with torch.cuda.device(a.device):
        ws = _workspace(lib.pcst_nn_min_workspace_bytes(B, N, M), a.device)
        _call("pcst_nn_min_f32", _p(a), _p(b), B, N, M, form, _p(rowmin), _p(rowarg) if want_arg else None, _p(ws),
                                  ws.numel(), _stream())
    return rowmin, rowarg


@nn_min.register_fake
def _(a, b, form, want_arg):
    B, N = a.shape[0], a.shape[1]
    return a.new_empty(B, N, dtype=torch.float32), a.new_empty((B, N) if want_arg else (0,), dtype=torch.int64)


@torch.library.custom_op("pcst::nn_min_pair", mutates_args=(), device_types="cuda")
def nn_min_pair(a: Tensor, b: Tensor, form: int) -> Tuple[Tensor, Tensor]:
    """Row AND column minima of the pair matrix between a [B,N,3] and b [B,M,3] in one sweep (every pair is
    evaluated once; the second direction's matrix is exactly the transpose).  form 0 = loss form,
    form 1 = cdist (a = x1).  -> (rowmin [B,N], colmin [B,M]), bit-identical to two ``nn_min`` calls."""
    lib = _lib.load()
    _need_cuda(a, b)
    a, b = _f32c(a), _f32c(b)
    B, N, _ = a.shape
    M = b.shape[1]
    rowmin = torch.empty(B, N, dtype=torch.float32, device=a.device)
    colmin = torch.empty(B, M, dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        ws = _workspace(lib.pcst_nn_min_pair_workspace_bytes(B, N, M), a.device)
        _call("pcst_nn_min_pair_f32", _p(a), _p(b), B, N, M, form, _p(rowmin), _p(colmin), _p(ws), ws.numel(), _stream())
    return rowmin, colmin


@nn_min_pair.register_fake
def _(a, b, form):
    return (a.new_empty(a.shape[0], a.shape[1], dtype=torch.float32),
            a.new_empty(b.shape[0], b.shape[1], dtype=torch.float32))


@torch.library.custom_op("pcst::nn_min_pair_arg", mutates_args=(), device_types="cuda")
def nn_min_pair_arg(a: Tensor, b: Tensor) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """Loss-form row / column minima AND their argmins from one sweep -> (rowmin [B,N], rowarg [B,N] int64,
    colmin [B,M], colarg [B,M] int64); ties go to the lowest index like ``torch.min``."""
    lib = _lib.load()
    _need_cuda(a, b)
    a, b = _f32c(a), _f32c(b)
    B, N, _ = a.shape
    M = b.shape[1]
    rowmin = torch.empty(B, N, dtype=torch.float32, device=a.device)
    colmin = torch.empty(B, M, dtype=torch.float32, device=a.device)
    rowarg = torch.empty(B, N, dtype=torch.int64, device=a.device)
    colarg = torch.empty(B, M, dtype=torch.int64, device=a.device)
    with torch.cuda.device(a.device):
        ws = _workspace(lib.pcst_nn_min_pair_arg_workspace_bytes(B, N, M), a.device)
        _call("pcst_nn_min_pair_arg_f32", _p(a), _p(b), B, N, M, _p(rowmin), _p(rowarg), _p(colmin), _p(colarg), _p(ws),
              ws.numel(), _stream())
    return rowmin, rowarg, colmin, colarg


@nn_min_pair_arg.register_fake
def _(a, b):
    B, N, M = a.shape[0], a.shape[1], b.shape[1]
    return (a.new_empty(B, N, dtype=torch.float32), a.new_empty(B, N, dtype=torch.int64),
            a.new_empty(B, M, dtype=torch.float32), a.new_empty(B, M, dtype=torch.int64))


@torch.library.custom_op("pcst::chamfer_bwd", mutates_args=(), device_types="cuda")
def chamfer_bwd(pred: Tensor, target: Tensor, arg_pt: Tensor, arg_tp: Tensor, grad_out: Tensor) -> Tuple[Tensor, Tensor]:
    lib = _lib.load()
    _need_cuda(pred, target, arg_pt, arg_tp, grad_out)
    pred, target, grad_out = _f32c(pred), _f32c(target), _f32c(grad_out)
    arg_pt, arg_tp = _i64c(arg_pt), _i64c(arg_tp)
    B, N, _ = pred.shape
    M = target.shape[1]
    gp, gt = torch.empty_like(pred), torch.empty_like(target)
    with torch.cuda.device(pred.device):
        _call("pcst_chamfer_bwd_f32", _p(pred), _p(target), _p(arg_pt), _p(arg_tp), _p(grad_out), B, N, M, _p(gp),
                                       _p(gt), _stream())
    return gp, gt


@chamfer_bwd.register_fake
def _(pred, target, arg_pt, arg_tp, grad_out):
    return torch.empty_like(pred, dtype=torch.float32), torch.empty_like(target, dtype=torch.float32)


class _ChamferLoss(torch.autograd.Function):
    """chamfer[b] = mean_i min_j D(p_i, t_j) + mean_j min_i D(t_j, p_i) (models/losses.py:61)."""

    @staticmethod
    def forward(ctx, pred, target):
        need_grad = pred.requires_grad or target.requires_grad
        if need_grad:  # the backward needs both argmins: still one sweep (block tracking + exact fix-up)
            d1, a1, d2, a2 = nn_min_pair_arg(pred, target)
            ctx.save_for_backward(pred, target, a1, a2)
        else:          # values only: one sweep serves both directions
            d1, d2 = nn_min_pair(pred, target, 0)
        out = d1.mean(dim=1) + d2.mean(dim=1)
        # The kernels' minima drop NaN (FMNMX returns the other operand).  The reference's clamp / min propagate it: a
        # non-finite coordinate anywhere in a cloud makes every column minimum of that element NaN, hence the loss.  Keep
        # that signal (a diverging training run must show up as NaN): one fused check per call.
        finite = torch.isfinite(pred.detach().sum(dim=(1, 2)) + target.detach().sum(dim=(1, 2)))
        return torch.where(finite, out, torch.full_like(out, float("nan")))

    @staticmethod
    def backward(ctx, grad):
        pred, target, a1, a2 = ctx.saved_tensors
        gp, gt = chamfer_bwd(pred, target, a1, a2, grad)
        return gp.to(pred.dtype), gt.to(target.dtype)


def chamfer_loss(pred: Tensor, target: Tensor) -> Tensor:
    _need_cuda(pred, target)
    return _ChamferLoss.apply(pred, target)


# ------------------------------------------------------------------------------------------ kNN


@torch.library.custom_op("pcst::knn", mutates_args=(), device_types="cuda")
def knn(query: Tensor, ref: Tensor, k: int) -> Tuple[Tensor, Tensor]:
    """query [B,Q,3], ref [B,R,3] -> (dist [B,Q,k] fp64 ascending, idx [B,Q,k] int64); sklearn order."""
    lib = _lib.load()
    _need_cuda(query, ref)
    if not 1 <= k <= 16:
        raise ValueError(f"knn: k={k} outside the kernels' range 1..16 (the sorted candidate list lives in registers); "
                         "the reference's sklearn call accepts any k -- split the query or use k <= 16")
    query, ref = _f32c(query), _f32c(ref)
    B, Q, _ = query.shape
    R = ref.shape[1]
    idx = torch.empty(B, Q, k, dtype=torch.int64, device=query.device)
    dist = torch.empty(B, Q, k, dtype=torch.float64, device=query.device)
    with torch.cuda.device(query.device):
        nb = lib.pcst_knn_workspace_bytes(B, Q, R, k)
        ws = _workspace(nb, query.device)
        _call("pcst_knn_f32", _p(query), _p(ref), B, Q, R, k, _p(idx), _p(dist), _p(ws), ws.numel(), _stream(),
              kernels=int(lib.pcst_knn_kernel_launches(B, Q, R, k, int(query.data_ptr() == ref.data_ptr()))))
    return dist, idx


@knn.register_fake
def _(query, ref, k):
    B, Q = query.shape[0], query.shape[1]
    return query.new_empty(B, Q, k, dtype=torch.float64), query.new_empty(B, Q, k, dtype=torch.int64)


@torch.library.custom_op("pcst::knn_interpolate", mutates_args=(), device_types="cuda")
def knn_interpolate(feat: Tensor, idx: Tensor, dist: Tensor) -> Tensor:
    """feat [B,R,C] fp32, idx/dist [B,Q,k] -> [B,Q,C] inverse-distance weighted (fp64 weights)."""
    lib = _lib.load()
    _need_cuda(feat, idx, dist)
    feat, idx = _f32c(feat), _i64c(idx)
    dist = dist.double().contiguous()
    B, R, C = feat.shape
    _, Q, k = idx.shape
    out = torch.empty(B, Q, C, dtype=torch.float32, device=feat.device)
    with torch.cuda.device(feat.device):
        _call("pcst_knn_interpolate_f32", _p(feat), _p(idx), _p(dist), B, R, Q, k, C, _p(out), _stream())
    return out


@knn_interpolate.register_fake
def _(feat, idx, dist):
    return feat.new_empty(feat.shape[0], idx.shape[1], feat.shape[2], dtype=torch.float32)


# ----------------------------------------------------------------------------- voxel-grid downsample


@torch.library.custom_op("pcst::minmax", mutates_args=(), device_types="cuda")
def minmax(xyz: Tensor) -> Tensor:
    """xyz [B,N,3] -> [B,6] = (min x, min y, min z, max x, max y, max z) per cloud."""
    _need_cuda(xyz)
    xyz = _f32c(xyz)
    B, N, _ = xyz.shape
    out = torch.empty(B, 6, dtype=torch.float32, device=xyz.device)
    with torch.cuda.device(xyz.device):
        _call("pcst_minmax_f32", _p(xyz), B, N, _p(out), _stream())
    return out


@minmax.register_fake
def _(xyz):
    return xyz.new_empty(xyz.shape[0], 6, dtype=torch.float32)


@torch.library.custom_op("pcst::voxel_representatives", mutates_args=(), device_types="cuda")
def voxel_representatives(xyz: Tensor, xyz_min: Tensor, voxel_size: Tensor) -> Tuple[Tensor, Tensor]:
    """Deterministic part of the reference's voxel-grid downsample (models/diffusion_model.py:86-93).
    xyz [B,N,3], xyz_min [B,3], voxel_size [B] -> (rep [B,N] int64, count [B] int32): row b's first count[b]
    entries are the per-voxel mean point indices in torch.unique (ascending hash) order."""
    lib = _lib.load()
    _need_cuda(xyz, xyz_min, voxel_size)
    xyz, xyz_min, voxel_size = _f32c(xyz), _f32c(xyz_min), _f32c(voxel_size)
    B, N, _ = xyz.shape
    rep = torch.empty(B, N, dtype=torch.int64, device=xyz.device)
    count = torch.empty(B, dtype=torch.int32, device=xyz.device)
    with torch.cuda.device(xyz.device):
        ws = _workspace(lib.pcst_voxel_representatives_workspace_bytes(B, N), xyz.device)
        _call("pcst_voxel_representatives_f32", _p(xyz), B, N, _p(xyz_min), _p(voxel_size), _p(rep), _p(count), _p(ws),
              ws.numel(), _stream())
    return rep, count


@voxel_representatives.register_fake
def _(xyz, xyz_min, voxel_size):
    B, N = xyz.shape[0], xyz.shape[1]
    return xyz.new_empty(B, N, dtype=torch.int64), xyz.new_empty(B, dtype=torch.int32)


# ------------------------------------------------------------- SA MLP, training mode (batch-stat BatchNorm + autograd)


def sa_mlp_train_supported(D: int, couts: List[int]) -> bool:
    """Shapes the train-mode tensor-core kernels cover: 3 layers, Cout a multiple of 16 and <= 512, 3 + D <= 512."""
    return len(couts) == 3 and all(c % 16 == 0 and 0 < c <= 512 for c in couts) and 3 + D <= 512


class _SaMlpTrain(torch.autograd.Function):
    """relu(bn(conv(.))) x 3 + max over each group, nn.BatchNorm2d in TRAINING mode (models/pointnet2_encoder.py:74,
    106-112), forward and backward on the tcgen05 kernels of csrc/sa_mlp_train.cu.  Running statistics and
    num_batches_tracked are updated in place by the forward, exactly once per call, like the module would."""

    @staticmethod
    def forward(ctx, xyz, feats, new_xyz, idx, eps, momentum, precision, bn_buffers, *params):
        # params = (w0, b0, gamma0, beta0, w1, ..., beta2); bn_buffers = [(running_mean, running_var, num_batches_tracked)] * 3
        lib = _lib.load()
        _need_cuda(xyz, feats, new_xyz, idx, *params)
        xyz = _f32c(xyz.detach())
        featsc = None if feats is None else _f32c(feats.detach())
        new_xyzc = None if new_xyz is None else _f32c(new_xyz.detach())
        idxc = None if idx is None else _i64c(idx)
        B, N, _ = xyz.shape
        D = 0 if featsc is None else featsc.shape[2]
        S, K = (1, N) if idxc is None else (idxc.shape[1], idxc.shape[2])
        ps = [_f32c(p.detach().reshape(p.shape[0], -1)) if p.dim() > 1 else _f32c(p.detach()) for p in params]
        m = _lib.Mlp3Train()
        for l in range(3):
            w, b, g, be = ps[4 * l: 4 * l + 4]
            m.w[l], m.bias[l], m.gamma[l], m.beta[l] = w.data_ptr(), b.data_ptr(), g.data_ptr(), be.data_ptr()
            m.cout[l] = w.shape[0]
            rm, rv, nbt = bn_buffers[l]
            m.running_mean[l] = 0 if rm is None else rm.data_ptr()
            m.running_var[l] = 0 if rv is None else rv.data_ptr()
            m.num_batches_tracked[l] = 0 if nbt is None else nbt.data_ptr()
        m.eps, m.momentum = float(eps), float(momentum)
        couts = [int(m.cout[l]) for l in range(3)]
        c3 = _cout3(couts)
        dev = xyz.device
        out = torch.empty(B, S, couts[2], dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            nsaved = lib.pcst_sa_mlp_train_saved_bytes(B, S, K, D, c3)
            if nsaved == 0:
                raise ValueError("sa_mlp_train: unsupported layer widths (Cout must be a multiple of 16, <= 512; 3 + D <= 512)")
            saved = torch.empty(nsaved, dtype=torch.uint8, device=dev)
            ws = _workspace(lib.pcst_sa_mlp_train_workspace_bytes(B, S, K, D, c3, int(precision), 0), dev)
            _call("pcst_sa_mlp_max_bnstats_bf16", _p(xyz), _p(featsc), _p(new_xyzc), _p(idxc), B, N, S, K, D, ctypes.byref(m),
                  int(precision), _p(out), _p(saved), saved.numel(), _p(ws), ws.numel(), _stream())
        ctx.save_for_backward(xyz, featsc, new_xyzc, idxc, saved, *ps)
        ctx.dims = (B, N, S, K, D, couts, float(eps), float(momentum), int(precision))
        ctx.param_shapes = [p.shape for p in params]
        ctx.grouped_grad = (feats is not None and feats.requires_grad) or xyz.requires_grad or \
            (new_xyz is not None and new_xyz.requires_grad)
        ctx.needs = (ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2])
        return out

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        xyz, feats, new_xyz, idx, saved, *ps = ctx.saved_tensors
        B, N, S, K, D, couts, eps, momentum, precision = ctx.dims
        dev = xyz.device
        grad_out = _f32c(grad_out)
        m = _lib.Mlp3Train()
        gr = _lib.Mlp3Grads()
        grads = []
        for l in range(3):
            w, b, g, be = ps[4 * l: 4 * l + 4]
            m.w[l], m.bias[l], m.gamma[l], m.beta[l] = w.data_ptr(), b.data_ptr(), g.data_ptr(), be.data_ptr()
            m.cout[l] = w.shape[0]
            gl = [torch.empty_like(t) for t in (w, b, g, be)]
            gr.w[l], gr.bias[l], gr.gamma[l], gr.beta[l] = [t.data_ptr() for t in gl]
            grads += gl
        m.eps, m.momentum = eps, momentum
        c3 = _cout3(couts)
        want_grouped = any(ctx.needs)
        gg = torch.empty(B, S, K, 3 + D, dtype=torch.float32, device=dev) if want_grouped else None
        with torch.cuda.device(dev):
            ws = _workspace(lib.pcst_sa_mlp_train_workspace_bytes(B, S, K, D, c3, precision, 1), dev)
            _call("pcst_sa_mlp_max_bwd_bf16", _p(xyz), _p(feats), _p(new_xyz), _p(idx), B, N, S, K, D, ctypes.byref(m),
                  precision, _p(saved), saved.numel(), _p(grad_out), ctypes.byref(gr), _p(gg), _p(ws), ws.numel(), _stream())
        g_xyz = g_feats = g_new = None
        if want_grouped:
            if idx is None:   # group_all: the group is the cloud itself, in order, nothing subtracted
                if ctx.needs[0]:
                    g_xyz = gg[:, 0, :, :3].contiguous()
                if ctx.needs[1]:
                    g_feats = gg[:, 0, :, 3:].contiguous()
            else:
                if ctx.needs[0]:
                    g_xyz = index_points_bwd(gg[..., :3].contiguous(), idx, N)
                if ctx.needs[1]:
                    g_feats = index_points_bwd(gg[..., 3:].contiguous(), idx, N)
                if ctx.needs[2]:
                    g_new = -gg[..., :3].sum(dim=2)
        pg = [g.reshape(s) for g, s in zip(grads, ctx.param_shapes)]
        return (g_xyz, g_feats, g_new, None, None, None, None, None, *pg)


def sa_mlp_train_unpack_saved(saved: Tensor, rows: int, groups: int, couts: List[int]):
    """Views into the ``saved`` blob of the train-mode forward (layout of train_plan in csrc/sa_mlp_train.cu):
    -> ([Z_l fp32 [rows, C_l]], [stat_l fp32 [4, C_l] = mean | 1/sqrt(var + eps) | a | b], argmax int32 [groups, C_2];
    -1 where the pooled value was clipped by the ReLU).  For tests and debugging."""
    al = lambda v: (v + 255) // 256 * 256
    off, Z, stat = 0, [], []
    for c in couts:
        Z.append(saved[off: off + rows * c * 4].view(torch.float32).reshape(rows, c))
        off += al(rows * c * 4)
    for c in couts:
        stat.append(saved[off: off + 16 * c].view(torch.float32).reshape(4, c))
        off += al(16 * c)
    argmax = saved[off: off + groups * couts[2] * 4].view(torch.int32).reshape(groups, couts[2])
    return Z, stat, argmax


KERNELS_PER_CALL["pcst_sa_mlp_max_bnstats_bf16"] = 13   # 3 x (pack, GEMM, column sums, finalize) + pool
KERNELS_PER_CALL["pcst_sa_mlp_max_bwd_bf16"] = 18       # 3 packs + 3 x (sums, finalize, wgrad, reduce, dgrad)


def sa_mlp_train(xyz: Tensor, feats: Optional[Tensor], new_xyz: Optional[Tensor], idx: Optional[Tensor],
                 convs, bns, precision: int = 1) -> Tensor:
    """Train-mode shared MLP + max-pool of one SetAbstraction stage -> [B,S,Cout] point-major.
    ``convs`` / ``bns``: the module's three Conv2d(1x1) / BatchNorm2d layers (parameters differentiable, running
    statistics updated in place).  ``precision``: 1 = bf16 operands, 0 = split (bf16x3, fp32-faithful) operands."""
    params, buffers = [], []
    for conv, bn in zip(convs, bns):
        params += [conv.weight, conv.bias, bn.weight, bn.bias]
        if bn.track_running_stats:
            buffers.append((bn.running_mean, bn.running_var, bn.num_batches_tracked))
        else:
            buffers.append((None, None, None))
    return _SaMlpTrain.apply(xyz, feats, new_xyz, idx, bns[0].eps, bns[0].momentum, int(precision), buffers, *params)


# ------------------------------------------------------------------------- NoisePredictor (fused per-point MLP chain)

KERNELS_PER_CALL["pcst_noise_predictor_f32"] = 2          # conditioning prep + the fused chain
KERNELS_PER_CALL["pcst_noise_predictor_pack_f32"] = 0      # size dependent: pcst_noise_predictor_pack_launches


def noise_predictor_supported(feature_dim: int, time_dim: int, nblocks: int) -> bool:
    return feature_dim % 16 == 0 and 16 <= feature_dim <= 256 and time_dim % 2 == 0 and 4 <= time_dim <= 1024 and 0 <= nblocks <= 8


def noise_predictor_pack(pe, time_proj, style_proj, blocks, out) -> Tensor:
    """Pack a NoisePredictor's nn.Linear parameters once per parameter version -> uint8 blob (bf16 UMMA weight blocks in
    the kernel's step order, fp32 biases and conditioning projections).  ``pe`` / ``out``: the three Linear layers of
    point_encoder / output_mlp; ``blocks``: [(Linear(F,2F), Linear(2F,F))]."""
    lib = _lib.load()
    m = _lib.NoiseMlp()
    keep = []

    def ptr(t):
        t = _f32c(t.detach())
        keep.append(t)
        return t.data_ptr()

    for i in range(3):
        m.pe_w[i], m.pe_b[i] = ptr(pe[i].weight), ptr(pe[i].bias)
        m.out_w[i], m.out_b[i] = ptr(out[i].weight), ptr(out[i].bias)
    m.time_w, m.time_b = ptr(time_proj.weight), ptr(time_proj.bias)
    m.style_w, m.style_b = ptr(style_proj.weight), ptr(style_proj.bias)
    for i, (l1, l2) in enumerate(blocks):
        m.blk_w1[i], m.blk_b1[i], m.blk_w2[i], m.blk_b2[i] = ptr(l1.weight), ptr(l1.bias), ptr(l2.weight), ptr(l2.bias)
    m.feature_dim, m.time_dim, m.nblocks = style_proj.out_features, time_proj.in_features, len(blocks)
    dev = pe[0].weight.device
    _need_cuda(pe[0].weight)
    with torch.cuda.device(dev):
        nb = lib.pcst_noise_predictor_packed_bytes(m.feature_dim, m.time_dim, m.nblocks)
        if nb == 0:
            raise ValueError("noise_predictor_pack: unsupported sizes (feature_dim a multiple of 16 in [16, 256], <= 8 blocks)")
        packed = torch.empty(nb, dtype=torch.uint8, device=dev)
        _call("pcst_noise_predictor_pack_f32", ctypes.byref(m), _p(packed), nb, _stream(),
              kernels=int(lib.pcst_noise_predictor_pack_launches(m.feature_dim, m.time_dim, m.nblocks)))
    return packed


@torch.library.custom_op("pcst::noise_predictor", mutates_args=(), device_types="cuda")
def noise_predictor(points: Tensor, timestep: Tensor, style: Tensor, packed: Tensor, feature_dim: int, time_dim: int,
                    nblocks: int) -> Tensor:
    """points [B,N,3] fp32, timestep [B] int64, style [B,F] -> predicted noise [B,N,3] (eval-mode NoisePredictor)."""
    lib = _lib.load()
    _need_cuda(points, timestep, style, packed)
    points, style, timestep = _f32c(points), _f32c(style), _i64c(timestep)
    B, N, _ = points.shape
    out = torch.empty(B, N, 3, dtype=torch.float32, device=points.device)
    with torch.cuda.device(points.device):
        ws = _workspace(lib.pcst_noise_predictor_workspace_bytes(B, feature_dim, nblocks), points.device)
        _call("pcst_noise_predictor_f32", _p(points), _p(timestep), _p(style), B, N, feature_dim, time_dim, nblocks, _p(packed),
              _p(out), _p(ws), ws.numel(), _stream())
    return out


@noise_predictor.register_fake
def _(points, timestep, style, packed, feature_dim, time_dim, nblocks):
    return points.new_empty(points.shape[0], points.shape[1], 3, dtype=torch.float32)


# ------------------------------------------------------------------- query-sharded Chamfer: pack / finish kernels

KERNELS_PER_CALL["pcst_chamfer_shard_pack_f32"] = 1
KERNELS_PER_CALL["pcst_chamfer_shard_finish_f32"] = 1


def chamfer_shard_pack(rowmin: Tensor, colmin: Tensor) -> Tensor:
    """rowmin [B,n] (this rank's complete row minima), colmin [B,M] (partial column minima) -> payload [B, M + 256] fp32 =
    colmin | 128 fp64 partial row sums as float pairs: what the ranks all-gather."""
    lib = _lib.load()
    _need_cuda(rowmin, colmin)
    rowmin, colmin = _f32c(rowmin), _f32c(colmin)
    B, M = colmin.shape
    n = rowmin.shape[1]
    payload = torch.empty(B, lib.pcst_chamfer_shard_payload_floats(M), dtype=torch.float32, device=colmin.device)
    with torch.cuda.device(colmin.device):
        _call("pcst_chamfer_shard_pack_f32", _p(rowmin), _p(colmin), B, n, M, _p(payload), _stream())
    return payload


def chamfer_shard_finish(gathered: Tensor, n_total: int, form: int) -> Tensor:
    """gathered [G, B, M + 256] (the all-gathered payloads) -> chamfer [B] fp32 (identical on every rank)."""
    lib = _lib.load()
    _need_cuda(gathered)
    gathered = _f32c(gathered)
    G, B, P = gathered.shape
    M = P - (lib.pcst_chamfer_shard_payload_floats(1) - 1)
    out = torch.empty(B, dtype=torch.float32, device=gathered.device)
    with torch.cuda.device(gathered.device):
        ws = _workspace(B * 128 * 8, gathered.device)
        _call("pcst_chamfer_shard_finish_f32", _p(gathered), G, B, M, int(n_total), int(form), _p(out), _p(ws), ws.numel(),
              _stream(), kernels=2)
    return out

"""CPU ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy / ctypes front-end of ``oracle/pcst_oracle.c``: a CPU restatement of the reference's
point-set hot path (wangxy0820/PointCloud_style_transfer).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this module, and only as the checker or the timed CPU baseline.  The product package
``pointcloud_style_transfer_b200`` never imports it.

Parity pin: the reference has no tests or golden vectors; this restatement is pinned against
outputs of the reference itself (``oracle/gen_golden.py`` -> ``tests/golden/*.npz``, and
``oracle/pin_against_reference.py`` -> ``oracle/PINNING.md`` at the full 120k-point sizes).

All arrays are numpy, C-contiguous; floating data fp32, indices int64 -- the reference's dtypes.
Every function cites the reference file:line it restates (paths relative to /root/reference).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Dict, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None

_f32p = ctypes.POINTER(ctypes.c_float)
_f64p = ctypes.POINTER(ctypes.c_double)
_i64p = ctypes.POINTER(ctypes.c_int64)


def build(force: bool = False) -> str:
    """Compile the C restatement with the committed Makefile (gcc, a few seconds)."""
    src = os.path.join(_HERE, "pcst_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.oracle_num_threads.restype = ctypes.c_int
    return _lib


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def set_num_threads(n: int) -> None:
    lib().oracle_set_num_threads(ctypes.c_int(int(n)))


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a: np.ndarray, t):
    return a.ctypes.data_as(t)


# --------------------------------------------------------------------------- encoder path


def square_distance(src, dst) -> np.ndarray:
    """models/pointnet2_encoder.py:8-15 -> [B,N,M] fp32."""
    src, dst = _f32(src), _f32(dst)
    B, N, _ = src.shape
    M = dst.shape[1]
    out = np.empty((B, N, M), np.float32)
    lib().oracle_square_distance(_ptr(src, _f32p), _ptr(dst, _f32p), B, N, M, _ptr(out, _f32p))
    return out


def index_points(points, idx) -> np.ndarray:
    """models/pointnet2_encoder.py:17-28: batched gather with idx clamped to [0, N-1]."""
    points = np.asarray(points)
    idx = np.clip(np.asarray(idx, dtype=np.int64), 0, points.shape[1] - 1)
    b = np.arange(points.shape[0]).reshape((-1,) + (1,) * (idx.ndim - 1))
    return points[b, idx]


def farthest_point_sample(xyz, npoint: int, start) -> np.ndarray:
    """models/pointnet2_encoder.py:30-45 with the random start index passed in -> [B,npoint] int64."""
    xyz = _f32(xyz)
    B, N, _ = xyz.shape
    start = np.ascontiguousarray(start, dtype=np.int64).reshape(B)
    out = np.empty((B, npoint), np.int64)
    lib().oracle_fps(_ptr(xyz, _f32p), B, N, int(npoint), _ptr(start, _i64p), _ptr(out, _i64p))
    return out


def radius_sq_f32(radius: float) -> np.float32:
    """``sqrdists > radius ** 2`` (:54): the Python double is cast to fp32 by type promotion."""
    return np.float32(float(radius) ** 2)


def query_ball_point(radius: float, nsample: int, xyz, new_xyz) -> np.ndarray:
    """models/pointnet2_encoder.py:47-59 -> [B,S,nsample] int64."""
    xyz, new_xyz = _f32(xyz), _f32(new_xyz)
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    out = np.empty((B, S, nsample), np.int64)
    lib().oracle_ball_query(_ptr(xyz, _f32p), _ptr(new_xyz, _f32p), B, N, S,
                            ctypes.c_float(float(radius_sq_f32(radius))), int(nsample), _ptr(out, _i64p))
    return out


def apply_mlp(points, layers: Sequence[Dict[str, np.ndarray]], eps: float = 1e-5) -> np.ndarray:
    """SetAbstraction.apply_mlp, models/pointnet2_encoder.py:106-112, eval-mode BatchNorm.

    points [B,S,K,C] -> [B,C_out,S].  ``layers`` = list of dicts with conv ``weight`` [Co,Ci]
    (or [Co,Ci,1,1]), ``bias`` [Co] and BN ``gamma, beta, mean, var`` [Co].
    """
    x = _f32(points)
    for L in layers:
        W = _f32(L["weight"]).reshape(L["weight"].shape[0], -1)
        y = x @ W.T + _f32(L["bias"])
        y = (y - _f32(L["mean"])) / np.sqrt(_f32(L["var"]) + np.float32(eps)) * _f32(L["gamma"]) + _f32(L["beta"])
        x = np.maximum(y, np.float32(0)).astype(np.float32)
    return np.ascontiguousarray(x.max(axis=2).transpose(0, 2, 1))


def apply_mlp_c(points, layers: Sequence[Dict[str, np.ndarray]], eps: float = 1e-5) -> np.ndarray:
    """``apply_mlp`` evaluated by the C oracle (``oracle_apply_mlp3``, OpenMP over groups) instead of numpy/BLAS:
    same formula, sequential summation.  bench.py's CPU baseline uses it so that the whole encoder runs inside ONE
    OpenMP runtime (an OpenMP pool next to a BLAS pool oversubscribes the cores and slows both)."""
    x = _f32(points)
    B, S, K, C0 = x.shape
    assert len(layers) == 3
    fp = ctypes.POINTER(ctypes.c_float)
    keep = []

    def arr(key, transpose=False):
        vals = []
        for L in layers:
            v = _f32(L[key])
            if transpose:
                v = np.ascontiguousarray(v.reshape(v.shape[0], -1).T)
            vals.append(v)
        keep.append(vals)
        return (fp * 3)(*[_ptr(v, fp) for v in vals])

    cout = (ctypes.c_int * 3)(*[int(L["weight"].shape[0]) for L in layers])
    out = np.empty((B * S, cout[2]), np.float32)
    lib().oracle_apply_mlp3(_ptr(x, fp), ctypes.c_long(B * S), int(K), int(C0), cout, arr("weight", True), arr("bias"),
                            arr("gamma"), arr("beta"), arr("mean"), arr("var"), ctypes.c_float(eps), _ptr(out, fp))
    return np.ascontiguousarray(out.reshape(B, S, -1).transpose(0, 2, 1))


def layers_from_state_dict(sd: Dict[str, np.ndarray], prefix: str) -> list:
    """Collect ``{prefix}mlp_convs.i.*`` / ``{prefix}mlp_bns.i.*`` (models/pointnet2_encoder.py:68-75)."""
    out, i = [], 0
    while f"{prefix}mlp_convs.{i}.weight" in sd:
        out.append(dict(weight=np.asarray(sd[f"{prefix}mlp_convs.{i}.weight"]),
                        bias=np.asarray(sd[f"{prefix}mlp_convs.{i}.bias"]),
                        gamma=np.asarray(sd[f"{prefix}mlp_bns.{i}.weight"]),
                        beta=np.asarray(sd[f"{prefix}mlp_bns.{i}.bias"]),
                        mean=np.asarray(sd[f"{prefix}mlp_bns.{i}.running_mean"]),
                        var=np.asarray(sd[f"{prefix}mlp_bns.{i}.running_var"])))
        i += 1
    return out


def set_abstraction(xyz, points, layers, npoint, radius, nsample, start, group_all=False, mlp=None):
    """SetAbstraction.forward, models/pointnet2_encoder.py:78-104 (eval mode).

    Returns (new_xyz [B,S,3], new_points [B,C_out,S] or [B,C_out] for group_all, fps_idx, group_idx).
    ``mlp`` = ``apply_mlp`` (numpy, default) or ``apply_mlp_c``.
    """
    apply_mlp = mlp or globals()["apply_mlp"]
    xyz = _f32(xyz)
    B, N, _ = xyz.shape
    if group_all:
        new_xyz = np.zeros((B, 1, 3), np.float32)
        g = xyz.reshape(B, 1, N, 3)
        if points is not None:
            g = np.concatenate([g, _f32(points).reshape(B, 1, N, -1)], axis=-1)
        return new_xyz, apply_mlp(g, layers)[:, :, 0], None, None
    fps_idx = farthest_point_sample(xyz, npoint, start)
    new_xyz = index_points(xyz, fps_idx)
    group_idx = query_ball_point(radius, nsample, xyz, new_xyz)
    g = index_points(xyz, group_idx) - new_xyz.reshape(B, npoint, 1, 3)
    if points is not None:
        g = np.concatenate([g, index_points(_f32(points), group_idx)], axis=-1)
    return new_xyz, apply_mlp(g, layers), fps_idx, group_idx


def encoder_forward(xyz, sd: Dict[str, np.ndarray], start1, start2, mlp=None) -> Dict[str, np.ndarray]:
    """PointNet2Encoder.forward, models/pointnet2_encoder.py:114-131 (eval mode).

    ``start1`` / ``start2`` are the FPS start indices of sa1 / sa2 (two consecutive
    ``torch.randint`` draws on the CPU generator in the reference, :36).
    """
    l1_xyz, l1_pts, f1, g1 = set_abstraction(xyz, None, layers_from_state_dict(sd, "sa1."), 512, 0.2, 32, start1, mlp=mlp)
    l2_xyz, l2_pts, f2, g2 = set_abstraction(l1_xyz, l1_pts.transpose(0, 2, 1), layers_from_state_dict(sd, "sa2."),
                                             128, 0.4, 64, start2, mlp=mlp)
    _, g, _, _ = set_abstraction(l2_xyz, l2_pts.transpose(0, 2, 1), layers_from_state_dict(sd, "sa3."),
                                 None, None, None, None, group_all=True, mlp=mlp)
    return dict(feature=g, l1_xyz=l1_xyz, l1_points=l1_pts, l2_xyz=l2_xyz, l2_points=l2_pts,
                fps1=f1, fps2=f2, group1=g1, group2=g2)


# --------------------------------------------------------------------------- NN reductions


def nn_min(a, b, form: int = 0, want_arg: bool = False):
    """Row minima of the pair matrix; form 0 = models/losses.py:36-41 (clamped squared,
    expanded), form 1 = evaluation/metrics.py:32 (torch.cdist mm path, Euclidean; ``a`` is cdist's
    x1), form 2 = the same cdist matrix reduced along the other axis (``a`` is cdist's x2, ``b`` its x1)."""
    a, b = _f32(a), _f32(b)
    B, N, _ = a.shape
    M = b.shape[1]
    rowmin = np.empty((B, N), np.float32)
    rowarg = np.empty((B, N), np.int64) if want_arg else None
    lib().oracle_nn_min(_ptr(a, _f32p), _ptr(b, _f32p), B, N, M, int(form), _ptr(rowmin, _f32p),
                        _ptr(rowarg, _i64p) if want_arg else None)
    return (rowmin, rowarg) if want_arg else rowmin


def _mean_f32(x: np.ndarray) -> np.ndarray:
    # torch.mean over fp32 accumulates in a vectorised/pairwise order that is implementation
    # defined; a float64 accumulation rounded once is within 1 ulp-ish of any such order.
    return x.astype(np.float64).mean(axis=1).astype(np.float32)


def chamfer_distance_chunked_optimized(pred, target, chunk_size: int = 1024) -> np.ndarray:
    """models/losses.py:8-63 -> [B] fp32: mean_i min_j D + mean_j min_i D (squared distances).
    ``chunk_size`` only bounds the reference's temporaries; it does not change any value."""
    d1 = nn_min(pred, target, 0)
    d2 = nn_min(target, pred, 0)
    return (_mean_f32(d1) + _mean_f32(d2)).astype(np.float32)


def metrics_chamfer_distance(pred, target, bidirectional: bool = True) -> np.ndarray:
    """evaluation/metrics.py:20-44 -> [B] fp32 (Euclidean, (mean+mean)/2)."""
    d1 = _mean_f32(nn_min(pred, target, 1))
    if not bidirectional:
        return d1
    d2 = _mean_f32(nn_min(target, pred, 2))
    return ((d1 + d2) / np.float32(2)).astype(np.float32)


def metrics_hausdorff_distance(pred, target) -> np.ndarray:
    """evaluation/metrics.py:90-105 -> [B] fp32."""
    return np.maximum(nn_min(pred, target, 1).max(axis=1), nn_min(target, pred, 2).max(axis=1))


def knn(query, ref, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """sklearn NearestNeighbors(n_neighbors=k).fit(ref).kneighbors(query) restated as fp64 brute
    force (models/diffusion_model.py:146-147) -> (dist [B,Q,k] fp64 ascending, idx [B,Q,k] int64)."""
    query, ref = _f32(query), _f32(ref)
    B, Q, _ = query.shape
    R = ref.shape[1]
    assert 1 <= k <= min(R, 64)
    idx = np.empty((B, Q, k), np.int64)
    dist = np.empty((B, Q, k), np.float64)
    lib().oracle_knn(_ptr(query, _f32p), _ptr(ref, _f32p), B, Q, R, int(k), _ptr(idx, _i64p), _ptr(dist, _f64p))
    return dist, idx


def upsample_knn(coarse_points, original_points, coarse_indices) -> np.ndarray:
    """HierarchicalProcessor.upsample_knn, models/diffusion_model.py:127-153 -> [B,N,3] fp32."""
    coarse_points, original_points = _f32(coarse_points), _f32(original_points)
    coarse_indices = np.asarray(coarse_indices, dtype=np.int64)
    B, N, _ = original_points.shape
    outs = []
    for b in range(B):
        result = np.zeros_like(original_points[b])
        ind = coarse_indices[b]
        valid = ind[ind < N]
        vals = coarse_points[b][: len(valid)]
        result[valid] = vals
        mask = np.ones(N, bool)
        mask[valid] = False
        unknown = np.where(mask)[0]
        if len(unknown) > 0 and len(valid) > 0:
            k = min(3, len(valid))
            fit = original_points[b][valid]
            dist, nbr = knn(original_points[b][unknown][None], fit[None], k)
            w = 1.0 / (dist[0] + 1e-8)
            w = w / w.sum(axis=1, keepdims=True)
            result[unknown] = np.sum(vals[nbr[0]] * w[..., None], axis=1)
        outs.append(result)
    return np.stack(outs)


def coverage_score(pred, target, threshold: float = 0.01) -> float:
    """evaluation/metrics.py:107-134: fraction of target points whose 1-NN in pred is < threshold."""
    dist, _ = knn(target, pred, 1)
    d = dist[..., 0]
    return float(np.mean([(d[b] < threshold).sum() / d.shape[1] for b in range(d.shape[0])]))


def uniformity_score(points, k: int = 8) -> float:
    """evaluation/metrics.py:136-170: 1/(1+cv) of the mean distance to the k nearest non-self points."""
    dist, _ = knn(points, points, k + 1)
    scores = []
    for b in range(dist.shape[0]):
        m = dist[b][:, 1:].mean(axis=1)
        mean = m.mean()
        scores.append(1.0 / (1.0 + m.std() / mean) if mean > 0 else 0.0)
    return float(np.mean(scores))


def voxel_representatives(points_b, xyz_min, voxel_size) -> np.ndarray:
    """Deterministic part of HierarchicalProcessor._voxel_grid_downsample_torch for ONE cloud
    (models/diffusion_model.py:86-93): voxel index, int32 hash, torch.unique order, truncated float32 mean of
    the member indices.  points_b [N,3] fp32, xyz_min [3] fp32, voxel_size fp32 scalar -> int64 [U]."""
    p = _f32(points_b)
    q = np.floor((p - _f32(xyz_min)) / np.float32(voxel_size)).astype(np.int32)          # :86
    u = q.astype(np.uint32)                                                               # wrapping int32 multiply
    h = (u[:, 0] * np.uint32(73856093)) ^ (u[:, 1] * np.uint32(19349663)) ^ (u[:, 2] * np.uint32(83492791))
    h = h.view(np.int32)                                                                  # :87
    _, inverse = np.unique(h, return_inverse=True)                                        # :89 (ascending, signed)
    sums = np.zeros(inverse.max() + 1, np.int64)
    np.add.at(sums, inverse, np.arange(len(p), dtype=np.int64))                           # :91-92
    counts = np.bincount(inverse).astype(np.int64)
    return (sums.astype(np.float32) / counts.astype(np.float32)).astype(np.int64)         # :93 (float32 true division)


def voxel_size_like_reference(points_b, target_size: int) -> np.float32:
    """voxel_size of models/diffusion_model.py:78-85: fp32 range / product / division; the cube root of the
    0-dim tensor is evaluated by torch in double and rounded to fp32 (pinned in tests/test_oracle_golden.py against
    torch's own expression on random clouds), then multiplied by fp32(1.2)."""
    p = _f32(points_b)
    rng = p.max(axis=0) - p.min(axis=0)
    rng = np.where(rng < np.float32(1e-6), np.float32(1.0), rng).astype(np.float32)
    prod = np.float32(np.float32(rng[0] * rng[1]) * rng[2])
    x = np.float32(prod / np.float32(target_size))
    vs = np.float32(np.float32(float(x) ** (1 / 3)) * np.float32(1.2))
    return np.float32(1e-3) if vs < 1e-6 else vs


def voxel_grid_downsample(points, target_size: int, randperm) -> np.ndarray:
    """HierarchicalProcessor._voxel_grid_downsample_torch (models/diffusion_model.py:69-122) -> indices [B,target].
    ``randperm(n)`` must return the permutation the reference would draw at that point (the tests pass
    ``lambda n: torch.randperm(n).numpy()`` after seeding torch like the reference run)."""
    points = _f32(points)
    B, N, _ = points.shape
    if N <= target_size:
        return np.broadcast_to(np.arange(N, dtype=np.int64), (B, N)).copy()
    out = []
    for b in range(B):
        pts = points[b]
        rep = voxel_representatives(pts, pts.min(axis=0), voxel_size_like_reference(pts, target_size))
        cur = len(rep)
        if cur > target_size:
            final = rep[randperm(cur)[:target_size]]
        elif cur < target_size:
            mask = np.ones(N, bool)
            mask[rep] = False
            pool = np.arange(N, dtype=np.int64)[mask]
            if len(pool) > 0:
                k = min(target_size - cur, len(pool))
                final = np.concatenate([rep, pool[randperm(len(pool))[:k]]])
            else:
                final = rep
        else:
            final = rep
        out.append(final)
    return np.stack(out)

"""Seeded synthetic inputs shared by the tests and bench.py.

All generators run on the CPU and return fp32 ``torch.Tensor``s so that the CPU oracle and the
CUDA path see identical bits.  Nothing here touches the oracle or the reference.
"""
from __future__ import annotations

import numpy as np
import torch


def uniform_cloud(seed: int, B: int, N: int) -> torch.Tensor:
    """U(seed,B,N): ``torch.rand(B,N,3)*2-1`` -- the family of examples/benchmark.py:35,86."""
    g = torch.Generator().manual_seed(int(seed))
    return torch.rand(B, N, 3, generator=g, dtype=torch.float32) * 2 - 1


def normalize_point_cloud(points: np.ndarray) -> np.ndarray:
    """Centre on the mean and scale max-abs to 1.8 (the reference's normalisation,
    data/preprocessing.py:21-38, restated)."""
    centred = points - points.mean(axis=0, keepdims=True)
    scale = np.abs(centred).max()
    if scale > 0:
        centred = centred * (1.8 / scale)
    return centred.astype(np.float32)


def lidar_scan(seed: int, N: int = 120000) -> torch.Tensor:
    """L(seed,N): a synthetic 64-beam spinning-LiDAR scan, [1,N,3] fp32, normalised to +-1.8.

    64 beams (elevation -24.8..+2.0 deg) x 1875 azimuth steps = 120 000 rays from a sensor at
    z = 1.73 m over a ground plane and 40 seeded axis-aligned boxes, range cap 80 m, Gaussian
    range noise sigma 0.02 m.  Rays are ordered beam-major like a real sensor, so index order is
    spatially coherent (which is what the ball-query early exit and FPS plane ties care about).
    Smaller N = a seeded random subset in the original order.
    """
    rs = np.random.RandomState(int(seed))
    beams, steps = 64, 1875
    elev = np.deg2rad(np.linspace(-24.8, 2.0, beams))
    azim = np.linspace(0.0, 2 * np.pi, steps, endpoint=False)
    el, az = np.meshgrid(elev, azim, indexing="ij")
    d = np.stack([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el)], -1).reshape(-1, 3)
    origin = np.array([0.0, 0.0, 1.73])
    t = np.full(d.shape[0], 80.0)
    down = d[:, 2] < -1e-6
    t[down] = np.minimum(t[down], -origin[2] / d[down, 2])
    centres = rs.uniform(-40, 40, size=(40, 2))
    sizes = rs.uniform(1.0, 6.0, size=(40, 3))
    for c, s in zip(centres, sizes):
        lo = np.array([c[0] - s[0] / 2, c[1] - s[1] / 2, 0.0])
        hi = np.array([c[0] + s[0] / 2, c[1] + s[1] / 2, s[2]])
        with np.errstate(divide="ignore", invalid="ignore"):
            t0 = (lo - origin) / d
            t1 = (hi - origin) / d
        tn = np.nanmax(np.minimum(t0, t1), axis=1)
        tf = np.nanmin(np.maximum(t0, t1), axis=1)
        hit = (tn <= tf) & (tn > 0.5)
        t[hit] = np.minimum(t[hit], tn[hit])
    t = np.minimum(t + rs.normal(0.0, 0.02, size=t.shape), 80.0)
    pts = origin + d * t[:, None]
    if N < pts.shape[0]:
        keep = np.sort(rs.permutation(pts.shape[0])[:N])
        pts = pts[keep]
    elif N > pts.shape[0]:
        raise ValueError("lidar_scan generates at most 120000 points")
    return torch.from_numpy(normalize_point_cloud(pts))[None].contiguous()


def lattice(x: torch.Tensor, denom: int = 1024) -> torch.Tensor:
    """Q(.): snap to the k/denom lattice.  For |x|<=1 with denom 1024 (or |x|<=1.8 with 512)
    every product and partial sum of the path's fp32 distance arithmetic is exact, so results do
    not depend on evaluation order, and exact ties become common."""
    return torch.round(x * denom) / denom


def fps_start(seed: int, B: int, N: int) -> torch.Tensor:
    """The reference's start draw (models/pointnet2_encoder.py:36) with an explicit generator."""
    g = torch.Generator().manual_seed(int(seed))
    return torch.randint(0, N, (B,), generator=g, dtype=torch.long)

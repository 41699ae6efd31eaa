"""Drop-in for the reference's ``models/pointnet2_encoder.py`` on B200.

Same names, signatures, tensor layouts, index dtypes and ``state_dict`` keys as the reference
(file:line citations are relative to the reference root); the arithmetic runs in the sm_100a
kernels of libpcst.so through the ``pcst::*`` custom ops.  CUDA tensors only -- no CPU fallback.
"""
from typing import List, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops


_SIDE_STREAMS = {}  # device -> side stream of the overlapped inference schedule (PointNet2Encoder._forward_overlapped)


def _cuda_only(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor; pointcloud_style_transfer_b200 has no CPU fallback")


def square_distance(src: torch.Tensor, dst: torch.Tensor) -> torch.Tensor:
    """models/pointnet2_encoder.py:8-15.  [B,N,3], [B,M,3] -> [B,N,M] fp32, bit-exact with the
    reference's fp32 CPU result (always fp32: the op opts out of autocast)."""
    _cuda_only(src, "square_distance")
    return ops.square_distance(src, dst)


def index_points(points: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """models/pointnet2_encoder.py:17-28.  points [B,N,C], idx [B,S] or [B,S,K] (clamped to
    [0, N-1]) -> [B,S,C] / [B,S,K,C]; differentiable w.r.t. ``points``."""
    _cuda_only(points, "index_points")
    return ops.index_points(points, idx)


def farthest_point_sample(xyz: torch.Tensor, npoint: int) -> torch.Tensor:
    """models/pointnet2_encoder.py:30-45.  xyz [B,N,3] -> [B,npoint] int64.

    The start index is drawn exactly like the reference does (:36): ``torch.randint`` on the CPU
    default generator, then moved to the device, so a seeded run consumes the same RNG stream."""
    _cuda_only(xyz, "farthest_point_sample")
    B, N, _ = xyz.shape
    farthest = torch.randint(0, N, (B,), dtype=torch.long).to(xyz.device)
    idx, _ = ops.fps(xyz, int(npoint), farthest)
    return idx


def query_ball_point(radius: float, nsample: int, xyz: torch.Tensor, new_xyz: torch.Tensor) -> torch.Tensor:
    """models/pointnet2_encoder.py:47-59.  -> [B,S,nsample] int64.  ``radius ** 2`` is rounded to
    fp32 as torch's type promotion does at :54.  Like the reference, nsample > N raises IndexError."""
    _cuda_only(xyz, "query_ball_point")
    if nsample > xyz.shape[1]:
        raise IndexError(f"query_ball_point: nsample={nsample} exceeds the number of points N={xyz.shape[1]} "
                         "(the reference raises IndexError at models/pointnet2_encoder.py:58)")
    return ops.ball_query(xyz, new_xyz, float(np.float32(float(radius) ** 2)), int(nsample))


class SetAbstraction(nn.Module):
    """models/pointnet2_encoder.py:61-112.  Parameters live in the same ``mlp_convs`` / ``mlp_bns``
    ModuleLists (Conv2d 1x1 + BatchNorm2d) so reference checkpoints load unchanged.

    Inference (``eval()`` and no gradient required) takes the fused path: FPS (+ centroid gather) ->
    ball query -> one fused gather + MLP + max-pool op with conv bias and BatchNorm folded.
    Training (``train()``) runs the NATIVE train-mode kernels (csrc/sa_mlp_train.cu): tcgen05 GEMMs for forward, dgrad
    and wgrad, batch-statistic BatchNorm with in-place running-stat updates, ReLU and max-pool gradients -- no cuDNN /
    cuBLAS.  ``mlp_precision == 1`` feeds bf16 operands (the autocast mode), ``mlp_precision == 0`` split bf16x3
    operands (fp32-faithful: tracks the reference's fp32 autograd).  The composition of the differentiable ``pcst``
    gather ops with torch's own Conv2d / BatchNorm2d remains only for eval-mode gradients (frozen BatchNorm), for layer
    widths outside the kernels' range, and as the explicit ``train_backend = "torch"`` comparison path of the tests."""

    #: 0 = fp32 CUDA-core MLP (exact-parity path), 1 = bf16 tcgen05 tensor-core MLP
    mlp_precision: int = 0
    #: "native" = train mode on the kernels of csrc/sa_mlp_train.cu; "torch" = the Conv2d / BatchNorm2d composition
    train_backend: str = "native"

    def __init__(self, npoint: int, radius: float, nsample: int,
                 in_channel: int, mlp: List[int], group_all: bool = False):
        super().__init__()
        self.npoint = npoint
        self.radius = radius
        self.nsample = nsample
        self.mlp_convs = nn.ModuleList()
        self.mlp_bns = nn.ModuleList()
        last_channel = in_channel + 3
        for out_channel in mlp:
            self.mlp_convs.append(nn.Conv2d(last_channel, out_channel, 1))
            self.mlp_bns.append(nn.BatchNorm2d(out_channel))
            last_channel = out_channel
        self.group_all = group_all
        self._fold_key = None
        self._fold = None
        self._packed = {}
        self._native_steps = 0  # train-mode kernel calls: they update the BatchNorm buffers behind torch's version counters

    # -- eval-mode folding of conv bias + BatchNorm into per-channel (scale, shift) ------------------
    def _folded(self):
        tensors = []
        for conv, bn in zip(self.mlp_convs, self.mlp_bns):
            tensors += [conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var]
        key = tuple((t.data_ptr(), t._version, t.device) for t in tensors) + (self._native_steps,)
        if key != self._fold_key:
            ws, scs, shs = [], [], []
            cin_pad = 0
            with torch.no_grad():
                for conv, bn in zip(self.mlp_convs, self.mlp_bns):
                    w = conv.weight.detach().reshape(conv.out_channels, -1).float()
                    scale = (bn.weight.double() / torch.sqrt(bn.running_var.double() + bn.eps))
                    shift = (conv.bias.double() - bn.running_mean.double()) * scale + bn.bias.double()
                    if cin_pad:  # previous layer was padded: its extra channels are identically 0
                        w = F.pad(w, (0, cin_pad))
                    cout_pad = (-conv.out_channels) % 32  # kernels need Cout % 32 == 0
                    if cout_pad:
                        w = F.pad(w, (0, 0, 0, cout_pad))
                        scale = F.pad(scale, (0, cout_pad))
                        shift = F.pad(shift, (0, cout_pad))
                    ws.append(w.contiguous())
                    scs.append(scale.float().contiguous())
                    shs.append(shift.float().contiguous())
                    cin_pad = cout_pad
            self._fold, self._fold_key = (ws, scs, shs), key
            self._packed = {}
        return self._fold

    def _packed_params(self, B: int, S: int, K: int):
        """The folded parameters in the layout the kernels read, packed once per parameter version, precision and
        cluster split (bf16 UMMA operand blocks for the tensor-core path; stages with few rows split every layer's
        channels over a thread-block cluster) -> (uint8 blob, padded Couts, cluster)."""
        ws, scs, shs = self._folded()
        prec = int(self.mlp_precision)
        couts = [int(w.shape[0]) for w in ws]
        D = ws[0].shape[1] - 3
        cluster = ops.sa_mlp_pick_cluster(B, S, K, D, couts, prec)
        key = (prec, cluster)
        if key not in self._packed:
            self._packed[key] = ops.sa_mlp_pack(ws, scs, shs, D, prec, cluster)
        return self._packed[key], couts, cluster

    def _fused_ok(self, *tensors) -> bool:
        if self.training or len(self.mlp_convs) != 3:
            return False
        if not torch.is_grad_enabled():
            return True
        needs = any(t is not None and t.requires_grad for t in tensors) or any(p.requires_grad for p in self.parameters())
        return not needs

    def _native_train_ok(self, D: int) -> bool:
        """Train mode on the native tensor-core kernels: three layers of supported widths, ordinary BatchNorm2d
        (affine, a numeric momentum, one eps), precision 1."""
        if not self.training or len(self.mlp_convs) != 3 or self.train_backend != "native":
            return False
        bns = list(self.mlp_bns)
        if any((not bn.affine) or bn.momentum is None or bn.eps != bns[0].eps or bn.momentum != bns[0].momentum for bn in bns):
            return False
        return ops.sa_mlp_train_supported(D, [conv.out_channels for conv in self.mlp_convs])

    def _mlp_train_native(self, xyz, points, new_xyz, group_idx) -> torch.Tensor:
        """-> [B,C_out,S] (channel-first view of the kernels' point-major output)."""
        out = ops.sa_mlp_train(xyz, points, new_xyz, group_idx, list(self.mlp_convs), list(self.mlp_bns),
                               precision=int(self.mlp_precision))
        self._native_steps += 1
        return out.permute(0, 2, 1)

    def forward(self, xyz: torch.Tensor, points: Optional[torch.Tensor] = None,
                start: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """``start`` (optional, [B] int64 on the device) overrides the FPS start draw; by default it is
        drawn from the CPU generator exactly like the reference (:36)."""
        _cuda_only(xyz, "SetAbstraction.forward")
        B, N, C = xyz.shape
        fused = self._fused_ok(xyz, points)
        cout = self.mlp_convs[-1].out_channels

        if self.group_all:
            new_xyz = torch.zeros(B, 1, 3, device=xyz.device)  # :82
            if fused:
                packed, couts, cluster = self._packed_params(B, 1, N)
                new_points = ops.sa_mlp_max(xyz, points, None, None, packed, couts, self.mlp_precision, cluster)  # [B,1,C]
                return new_xyz, new_points[:, 0, :cout]
            if self._native_train_ok(0 if points is None else points.shape[2]):
                return new_xyz, self._mlp_train_native(xyz, points, None, None)[:, :, 0]
            if points is not None:
                grouped_points = torch.cat([xyz.view(B, 1, N, 3), points.view(B, 1, N, -1)], dim=-1)
            else:
                grouped_points = xyz.view(B, 1, N, 3)
            new_points = self.apply_mlp(grouped_points)
            return new_xyz, new_points.squeeze(-1)

        new_xyz, group_idx = self._sample_group(xyz, start)
        if fused:
            return new_xyz, self._mlp_fused(xyz, points, new_xyz, group_idx)
        if self._native_train_ok(0 if points is None else points.shape[2]):
            return new_xyz, self._mlp_train_native(xyz, points, new_xyz, group_idx)
        new_points = ops.group(xyz, points, new_xyz, group_idx)                # :94-101
        new_points = self.apply_mlp(new_points)
        return new_xyz, new_points

    def _sample_group(self, xyz: torch.Tensor, start: Optional[torch.Tensor] = None):
        """FPS (+ centroid gather, :91-92) and ball query (:93) -> (new_xyz [B,S,3], group_idx [B,S,K])."""
        B, N, _ = xyz.shape
        if start is None:
            start = torch.randint(0, N, (B,), dtype=torch.long).to(xyz.device)  # :36, CPU generator
        fps_idx, new_xyz = ops.fps(xyz, int(self.npoint), start)
        if xyz.requires_grad and torch.is_grad_enabled():
            # the reference's new_xyz = index_points(xyz, fps_idx) (:92) is differentiable w.r.t. xyz: take it from the
            # differentiable gather instead of the sampling kernel's by-product
            new_xyz = ops.index_points(xyz, fps_idx)
        return new_xyz, query_ball_point(self.radius, self.nsample, xyz, new_xyz.detach())

    def _mlp_fused(self, xyz, points, new_xyz, group_idx) -> torch.Tensor:
        """Grouping gather + folded MLP + max-pool in one op -> [B,C_out,S] (a channel-first VIEW of the
        kernel's point-major output, so the caller's permute back, :128-129, is free)."""
        B, S, K = group_idx.shape
        packed, couts, cluster = self._packed_params(B, S, K)
        out = ops.sa_mlp_max(xyz, points, new_xyz, group_idx, packed, couts, self.mlp_precision, cluster)  # [B,S,C]
        return out[:, :, :self.mlp_convs[-1].out_channels].permute(0, 2, 1)

    def apply_mlp(self, points):
        """models/pointnet2_encoder.py:106-112.  points [B,S,K,C] -> [B,C_out,S]."""
        if self._fused_ok(points) and points.is_cuda:
            # a pre-grouped tensor: run it as B*S clouds of K points, one group each (group_all form)
            B, S, K, C = points.shape
            packed, couts, cluster = self._packed_params(B * S, 1, K)
            flat = points.reshape(B * S, K, C)
            xyz = flat[..., :3].contiguous()
            feats = flat[..., 3:].contiguous() if C > 3 else None
            out = ops.sa_mlp_max(xyz, feats, None, None, packed, couts, self.mlp_precision, cluster)  # [B*S,1,C]
            return out[:, 0, :self.mlp_convs[-1].out_channels].reshape(B, S, -1).permute(0, 2, 1)
        points = points.permute(0, 3, 1, 2)
        for conv, bn in zip(self.mlp_convs, self.mlp_bns):
            points = F.relu(bn(conv(points)))
        return torch.max(points, 3)[0]


class PointNet2Encoder(nn.Module):
    """models/pointnet2_encoder.py:114-131.  xyz [B,N,3] -> [B,feature_dim]; ``input_channels`` is
    accepted and ignored, as in the reference."""

    def __init__(self, input_channels: int = 3, feature_dim: int = 512, mlp_precision: int = 0):
        super().__init__()
        self.sa1 = SetAbstraction(512, 0.2, 32, in_channel=0, mlp=[64, 64, 128])
        self.sa2 = SetAbstraction(128, 0.4, 64, in_channel=128, mlp=[128, 128, 256])
        self.sa3 = SetAbstraction(npoint=None, radius=None, nsample=None,
                                  in_channel=256, mlp=[256, 512, feature_dim], group_all=True)
        self.set_mlp_precision(mlp_precision)

    def _forward_overlapped(self, xyz: torch.Tensor, s1, s2) -> torch.Tensor:
        """Inference schedule of the same three stages.  Stage 2's sampling and grouping (FPS over the 512
        stage-1 centroids, then ball query) only need stage 1's CENTROIDS, not its features, so they run on a
        side stream next to stage 1's ball query + MLP; the two streams join before stage 2's MLP.  Under
        ``runtime.GraphedEncoder`` the fork / join is captured into the CUDA graph as parallel branches."""
        B, N, _ = xyz.shape
        dev = xyz.device
        if s1 is None:  # the reference's two draws, in its order (:36 inside sa1, then inside sa2)
            s1 = torch.randint(0, N, (B,), dtype=torch.long).to(dev)
            s2 = torch.randint(0, self.sa1.npoint, (B,), dtype=torch.long).to(dev)
        main = torch.cuda.current_stream(dev)
        side = _SIDE_STREAMS.get(dev)
        if side is None:
            side = _SIDE_STREAMS[dev] = torch.cuda.Stream(device=dev)
        packs = [self.sa1._packed_params(B, self.sa1.npoint, self.sa1.nsample)[0],
                 self.sa2._packed_params(B, self.sa2.npoint, self.sa2.nsample)[0],
                 self.sa3._packed_params(B, 1, self.sa2.npoint)[0]]
        side.wait_stream(main)
        with torch.cuda.stream(side):
            for blob in packs:  # 1.3 MB of packed weights -> L2 while the first FPS occupies 16 of the 148 SMs
                ops.l2_prefetch(blob)
        _, l1_xyz = ops.fps(xyz, int(self.sa1.npoint), s1)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            l2_xyz, idx2 = self.sa2._sample_group(l1_xyz, s2)
        idx1 = query_ball_point(self.sa1.radius, self.sa1.nsample, xyz, l1_xyz)
        l1_points = self.sa1._mlp_fused(xyz, None, l1_xyz, idx1)                  # [B,128,512]
        main.wait_stream(side)
        if not torch.cuda.is_current_stream_capturing():
            for t in (l2_xyz, idx2):  # allocated on the side stream, consumed on the main stream
                t.record_stream(main)
        l2_points = self.sa2._mlp_fused(l1_xyz, l1_points.permute(0, 2, 1), l2_xyz, idx2)   # [B,256,128]
        # group_all stage without materialising its all-zero new_xyz (:82), which nothing downstream reads
        packed, couts, cluster = self.sa3._packed_params(B, 1, l2_xyz.shape[1])
        g = ops.sa_mlp_max(l2_xyz, l2_points.permute(0, 2, 1), None, None, packed, couts, self.sa3.mlp_precision, cluster)
        return g[:, 0, :self.sa3.mlp_convs[-1].out_channels].reshape(B, -1)

    def set_mlp_precision(self, precision: int) -> "PointNet2Encoder":
        """0 = fp32 CUDA cores (parity within rtol 1e-4), 1 = bf16 tcgen05 tensor cores (rtol 2e-2)."""
        for sa in (self.sa1, self.sa2, self.sa3):
            sa.mlp_precision = int(precision)
        return self

    #: FPS start indices (sa1, sa2) as static device buffers, read when ``forward`` gets no ``starts``: a captured CUDA
    #: graph cannot contain the reference's CPU-generator draw (:36), so its owner refills these before every replay
    static_starts: Optional[Tuple[torch.Tensor, torch.Tensor]] = None

    def forward(self, xyz: torch.Tensor, starts: Optional[Tuple[torch.Tensor, torch.Tensor]] = None) -> torch.Tensor:
        """``starts`` (optional) = the FPS start indices of sa1 / sa2 as device tensors; used by
        ``runtime.GraphedEncoder`` to keep the CPU RNG draws outside a captured CUDA graph."""
        B, N, C = xyz.shape
        points = None
        if starts is None:
            starts = self.static_starts   # set by a CUDA-graph runner that owns the encoder (train_step.DiffusionTrainStep)
        s1, s2 = starts if starts is not None else (None, None)
        if xyz.is_cuda and all(sa._fused_ok(xyz) for sa in (self.sa1, self.sa2, self.sa3)):
            return self._forward_overlapped(xyz, s1, s2)
        l1_xyz, l1_points = self.sa1(xyz, points, s1)                           # [B,128,512]
        l2_xyz, l2_points = self.sa2(l1_xyz, l1_points.permute(0, 2, 1), s2)    # [B,256,128]
        _, global_feature = self.sa3(l2_xyz, l2_points.permute(0, 2, 1))        # [B,feature_dim]
        return global_feature.view(B, -1)

"""Drop-in for ``HierarchicalProcessor`` of the reference's models/diffusion_model.py: the 3-NN
inverse-distance upsample (the north-star's "feature-propagation interpolation", §8(a) row a12) and the
voxel-grid downsample in front of the encoder (first "next" row)."""
import math
from typing import Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from .pointnet2_encoder import PointNet2Encoder


class HierarchicalProcessor:
    """models/diffusion_model.py:64-153."""

    #: which torch generator the random thinning / top-up of the downsample draws from.  The reference calls
    #: ``torch.randperm(n, device=pts.device)`` (:97,108); parity is defined against its CPU path, so the default
    #: consumes the CPU generator exactly like a CPU run of the reference and moves the permutation to the device.
    rng_device = "cpu"

    def __init__(self, total_points: int = 120000, global_points: int = 30000):
        self.total_points = total_points
        self.global_points = global_points

    def _randperm(self, n: int, device) -> torch.Tensor:
        if self.rng_device == "cpu":
            return torch.randperm(n).to(device)
        return torch.randperm(n, device=device)

    def _voxel_grid_downsample_torch(self, points: torch.Tensor, target_size: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """models/diffusion_model.py:69-122.  points [B,N,3] -> (downsampled [B,target,3], indices [B,target]).

        The bounding boxes, voxel hashes, ``torch.unique`` ordering and per-voxel mean indices of ALL batch elements
        come from two kernel launches (no per-element Python loop over 120k-point tensors); ``voxel_size`` is formed
        on the host with the reference's own scalar fp32 expression so that every voxel boundary falls where the
        reference puts it; the random thinning / top-up keeps the reference's draw order."""
        if points.shape[1] <= target_size:
            indices = torch.arange(points.shape[1], device=points.device)
            return points, indices.expand(points.shape[0], -1)
        if not points.is_cuda:
            raise RuntimeError("HierarchicalProcessor.downsample: expected CUDA tensors; there is no CPU fallback")
        B, N, _ = points.shape
        device = points.device
        pts_all = points.detach().float().contiguous()
        box = ops.minmax(pts_all).cpu()                       # [B,6]; the reference also syncs here (:83)
        sizes = []
        for b in range(B):
            xyz_range = box[b, 3:] - box[b, :3]               # :80
            xyz_range[xyz_range < 1e-6] = 1.0                 # :81
            voxel_size = (xyz_range.prod() / target_size) ** (1 / 3) * 1.2   # :83, fp32 tensor arithmetic
            if voxel_size < 1e-6:                             # :84-85
                voxel_size = torch.tensor(1e-3, dtype=torch.float32)
            sizes.append(voxel_size.reshape(()))
        rep, count = ops.voxel_representatives(pts_all, box[:, :3].to(device), torch.stack(sizes).to(device))
        counts = count.cpu().tolist()
        downsampled_list, indices_list = [], []
        for b in range(B):
            pts = points[b]
            current_size = counts[b]
            representative_indices = rep[b, :current_size]
            if current_size > target_size:                    # :96-98
                rand_indices = self._randperm(current_size, device)[:target_size]
                final_indices = representative_indices[rand_indices]
            elif current_size < target_size:                  # :99-112
                remaining_needed = target_size - current_size
                mask = torch.ones(N, dtype=torch.bool, device=device)
                mask[representative_indices] = False
                pool = torch.arange(N, device=device)[mask]
                if len(pool) > 0:
                    num_to_sample = min(remaining_needed, len(pool))
                    additional_indices = pool[self._randperm(len(pool), device)[:num_to_sample]]
                    final_indices = torch.cat([representative_indices, additional_indices])
                else:
                    final_indices = representative_indices
            else:
                final_indices = representative_indices
            downsampled_list.append(pts[final_indices])
            indices_list.append(final_indices)
        return torch.stack(downsampled_list), torch.stack(indices_list)

    def downsample(self, points: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """models/diffusion_model.py:124-125.  ``rng_device == "cuda"`` takes the device-resident variant."""
        if self.rng_device == "cuda" and points.is_cuda and points.shape[1] > self.global_points:
            return self.downsample_device(points)
        return self._voxel_grid_downsample_torch(points, self.global_points)

    def downsample_device(self, points: torch.Tensor, generator: Optional[torch.Generator] = None):
        """The same voxel-grid downsample (:69-122) with NOTHING leaving the device: no host synchronisation, fixed
        output shapes, capturable in a CUDA graph.  Bounding box and representatives come from
        the same kernels; ``voxel_size`` is formed by the reference's own tensor expression (:80-85) evaluated on the
        device, as the reference does when it runs on CUDA; the random thinning / top-up (:95-112) is one draw of uniform
        keys per point and a top-k: with at least ``target`` occupied voxels a uniformly random subset of the
        representatives (in random order), otherwise all representatives in ``torch.unique`` order followed by a
        uniformly random subset of the remaining points -- the reference's selection rule, with the device generator's
        stream instead of ``torch.randperm``'s."""
        target = self.global_points
        B, N, _ = points.shape
        device = points.device
        pts = points.detach().float().contiguous()
        box = ops.minmax(pts)                                                   # [B,6]
        xyz_range = box[:, 3:] - box[:, :3]                                     # :80
        xyz_range = torch.where(xyz_range < 1e-6, torch.ones_like(xyz_range), xyz_range)       # :81
        voxel_size = (xyz_range.prod(dim=1) / target) ** (1 / 3) * 1.2          # :83
        voxel_size = torch.where(voxel_size < 1e-6, torch.full_like(voxel_size, 1e-3), voxel_size)   # :84-85
        rep, count = ops.voxel_representatives(pts, box[:, :3].contiguous(), voxel_size)
        # Candidates: the N slots of the representative list (slot j valid while j < count; the list may name a point twice,
        # two voxels can share a truncated mean index, and the reference keeps both) followed by the N points themselves
        # as the top-up pool (points that are representatives excluded, :102-104).
        ar = torch.arange(N, device=device)
        valid = ar[None, :] < count[:, None]
        is_rep = torch.zeros(B, N, dtype=torch.bool, device=device)
        is_rep.scatter_(1, torch.where(valid, rep, rep[:, :1].expand(B, N)), torch.ones_like(valid))   # count >= 1 always
        key = torch.rand(B, 2 * N, device=device, generator=generator)
        enough = (count >= target)[:, None]
        never = torch.full((), 3.0, device=device)
        slot_score = torch.where(valid, torch.where(enough, key[:, :N], (ar[None, :] - N).float().expand(B, N)), never)
        pool_score = torch.where(enough | is_rep, never, key[:, N:])
        pick = torch.topk(torch.cat([slot_score, pool_score], 1), target, dim=1, largest=False, sorted=True).indices
        final_indices = torch.gather(torch.cat([rep, ar[None, :].expand(B, N)], 1), 1, pick)
        return torch.gather(points, 1, final_indices[..., None].expand(-1, -1, 3)), final_indices

    def upsample_knn_device(self, coarse_points: torch.Tensor, original_points: torch.Tensor,
                            coarse_indices: torch.Tensor) -> torch.Tensor:
        """``upsample_knn`` (:127-153) for the case the sampling loop produces -- ``coarse_indices`` unique and in range
        (the output of ``downsample``) -- without host synchronisation or data-dependent shapes: every point of the cloud
        queries its 3 nearest KNOWN points (a known point finds itself at distance 0) and the known points' own values
        are written over the interpolated ones, which is what the reference's ``result[idx] = coarse`` does."""
        C = coarse_points.shape[2]
        original = original_points.detach().float().contiguous()
        coarse = coarse_points.detach().float().contiguous()
        fit = torch.gather(original, 1, coarse_indices[..., None].expand(-1, -1, 3))
        k = min(3, fit.shape[1])
        dist, nbr = ops.knn(original, fit, k)
        out = ops.knn_interpolate(coarse, nbr, dist)
        return out.scatter(1, coarse_indices[..., None].expand(-1, -1, C), coarse)

    def upsample_knn(self, coarse_points: torch.Tensor, original_points: torch.Tensor,
                     coarse_indices: torch.Tensor) -> torch.Tensor:
        """models/diffusion_model.py:127-153.  coarse [B,M,C], original [B,N,3], indices [B,M] -> [B,N,C] fp32.

        Known points keep their value (``result[idx] = coarse``; on duplicate indices the last write
        wins, as in numpy); every other point gets the inverse-distance (1/(d+1e-8), un-squared d)
        average of its 3 nearest known points, searched and weighted in fp64 like sklearn/numpy.
        Everything stays on the device: no .cpu()/.numpy() round trip per element per step."""
        if not coarse_points.is_cuda:
            raise RuntimeError("upsample_knn: expected CUDA tensors; there is no CPU fallback")
        B, N_orig, _ = original_points.shape
        device = coarse_points.device
        original_points = original_points.detach().float()
        coarse_points = coarse_points.detach().float()
        outs = []
        for b in range(B):
            ind = coarse_indices[b]
            valid = ind[ind < N_orig]
            vals = coarse_points[b][: valid.numel()]
            result = torch.zeros(N_orig, coarse_points.shape[2], dtype=torch.float32, device=device)
            if valid.numel() > 0:
                # numpy fancy assignment: the LAST occurrence of a duplicated index wins
                order = torch.arange(valid.numel(), device=device)
                last = torch.full((N_orig,), -1, dtype=torch.long, device=device)
                last.scatter_reduce_(0, valid, order, reduce="amax", include_self=True)
                known = last >= 0
                result[known] = vals[last[known]]
            else:
                known = torch.zeros(N_orig, dtype=torch.bool, device=device)
            unknown = (~known).nonzero(as_tuple=True)[0]
            if unknown.numel() > 0 and valid.numel() > 0:
                k = min(3, valid.numel())
                fit = original_points[b][valid]
                dist, nbr = ops.knn(original_points[b][unknown][None], fit[None], k)
                result[unknown] = ops.knn_interpolate(vals[None], nbr, dist)[0]
            outs.append(result)
        return torch.stack(outs)


# ----------------------------------------------------------------------------------------------------------------------
# The callers either side of the hot path: the reference's model classes with the same
# names, constructor arguments, attribute / state_dict names and forward signatures (models/diffusion_model.py:15-61,
# 156-300), so that its checkpoints load and its trainer / inference scripts run against this package.  The per-point
# denoiser runs as one fused tcgen05 kernel at inference (ops.noise_predictor); in training its nn.Linear stack stays
# on torch autograd (library GEMMs: out of the hot-path scope) while the style encoder under it
# trains on the native kernels.
# ----------------------------------------------------------------------------------------------------------------------
class TimeEmbedding(nn.Module):
    """models/diffusion_model.py:15-26: sinusoidal embedding [sin(t f_j), cos(t f_j)], f_j = 10000^(-j / (dim/2 - 1))."""

    def __init__(self, dim: int):
        super().__init__()
        self.dim = dim

    def forward(self, t: torch.Tensor) -> torch.Tensor:
        half = self.dim // 2
        freq = torch.exp(torch.arange(half, device=t.device) * -(math.log(10000) / (half - 1)))
        arg = t[:, None] * freq[None, :]
        return torch.cat((arg.sin(), arg.cos()), dim=-1)


class StyleEncoder(nn.Module):
    """models/diffusion_model.py:28-36: PointNet2Encoder -> Linear-ReLU-Dropout-Linear-ReLU."""

    def __init__(self, feature_dim: int = 256, mlp_precision: int = 0):
        super().__init__()
        self.encoder = PointNet2Encoder(input_channels=3, feature_dim=feature_dim, mlp_precision=mlp_precision)
        self.style_mlp = nn.Sequential(nn.Linear(feature_dim, 512), nn.ReLU(), nn.Dropout(0.1),
                                       nn.Linear(512, feature_dim), nn.ReLU())

    def forward(self, points: torch.Tensor) -> torch.Tensor:
        return self.style_mlp(self.encoder(points))


class NoisePredictor(nn.Module):
    """models/diffusion_model.py:38-61.  ``eval()`` without gradients on CUDA runs the fused tensor-core kernel
    (csrc/noise_mlp_tc.cu, bf16 operands / fp32 accumulate, rtol 2e-2); training runs the nn.Linear stack."""

    #: False forces the nn.Linear formulation at inference too (fp32 comparison path of the tests)
    fused_inference: bool = True

    def __init__(self, config):
        super().__init__()
        self.config = config
        fd = config.feature_dim
        self.point_encoder = nn.Sequential(nn.Linear(3, 128), nn.ReLU(), nn.Linear(128, 256), nn.ReLU(), nn.Linear(256, fd))
        self.time_embedding = TimeEmbedding(config.time_embed_dim)
        self.time_proj = nn.Linear(config.time_embed_dim, fd)
        self.style_proj = nn.Linear(fd, fd)
        self.layers = nn.ModuleList([nn.Sequential(nn.Linear(fd, fd * 2), nn.ReLU(), nn.Linear(fd * 2, fd), nn.Dropout(0.1))
                                     for _ in range(6)])
        self.output_mlp = nn.Sequential(nn.Linear(fd, 256), nn.ReLU(), nn.Linear(256, 128), nn.ReLU(), nn.Linear(128, 3))
        self._packed = None
        self._packed_key = None

    def _packed_params(self) -> torch.Tensor:
        key = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if key != self._packed_key:
            self._packed = ops.noise_predictor_pack(
                [self.point_encoder[0], self.point_encoder[2], self.point_encoder[4]], self.time_proj, self.style_proj,
                [(blk[0], blk[2]) for blk in self.layers],
                [self.output_mlp[0], self.output_mlp[2], self.output_mlp[4]])
            self._packed_key = key
        return self._packed

    def _fused_ok(self, *tensors) -> bool:
        if self.training or not self.fused_inference or not tensors[0].is_cuda:
            return False
        if not ops.noise_predictor_supported(self.style_proj.out_features, self.time_proj.in_features, len(self.layers)):
            return False
        if not torch.is_grad_enabled():
            return True
        return not (any(t.requires_grad for t in tensors) or any(p.requires_grad for p in self.parameters()))

    def forward(self, noisy_points: torch.Tensor, timestep: torch.Tensor, style_feat: torch.Tensor) -> torch.Tensor:
        if self._fused_ok(noisy_points, style_feat):
            return ops.noise_predictor(noisy_points, timestep, style_feat, self._packed_params(),
                                       self.style_proj.out_features, self.time_proj.in_features, len(self.layers))
        point_feat = self.point_encoder(noisy_points)                                                   # :54
        time_feat = self.time_proj(self.time_embedding(timestep)).unsqueeze(1)                          # :55
        x = point_feat + time_feat + self.style_proj(style_feat).unsqueeze(1)                           # :56-57
        for layer in self.layers:                                                                       # :58-59
            x = layer(x) + x
        return self.output_mlp(x)                                                                       # :60


class PointCloudDiffusionModel(nn.Module):
    """models/diffusion_model.py:156-190."""

    def __init__(self, config, mlp_precision: int = 0):
        super().__init__()
        self.config = config
        self.style_encoder = StyleEncoder(feature_dim=config.feature_dim, mlp_precision=mlp_precision)
        self.noise_predictor = NoisePredictor(config)
        self.hierarchical_processor = HierarchicalProcessor(total_points=config.total_points,
                                                            global_points=config.global_points)

    def forward(self, noisy_points: torch.Tensor, timestep: torch.Tensor, condition_points: torch.Tensor,
                cond_drop_prob: float = 0.0, use_hierarchical: bool = True):
        hp, gp = self.hierarchical_processor, self.config.global_points
        if use_hierarchical and condition_points.shape[1] > gp:                                         # :169-173
            condition_points, _ = hp.downsample(condition_points)
        style_feat = self.style_encoder(condition_points)
        if cond_drop_prob > 0:                                                                          # :175-177
            keep = torch.rand(style_feat.shape[0], 1, device=style_feat.device) > cond_drop_prob
            style_feat = style_feat * keep
        if use_hierarchical and noisy_points.shape[1] > gp:                                             # :179-184
            noisy_coarse, noise_indices = hp.downsample(noisy_points)
            return self.noise_predictor(noisy_coarse, timestep, style_feat), noise_indices
        return self.noise_predictor(noisy_points, timestep, style_feat), None                           # :186-189


class DiffusionProcess:
    """models/diffusion_model.py:193-300: noise schedules, ``q_sample``, and the CFG-guided / plain DDIM sampling loops.
    The loops keep the reference's arithmetic step for step; the hierarchical pieces they call (voxel downsample, fused
    denoiser, 3-NN upsample) are this package's kernels, so nothing leaves the device inside a step."""

    def __init__(self, config, device: str = "cuda"):
        self.num_timesteps = config.num_timesteps
        self.device = device
        self.betas = self._get_beta_schedule(config.beta_schedule, config.noise_schedule_offset).to(device)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.alphas_cumprod_prev = F.pad(self.alphas_cumprod[:-1], (1, 0), value=1.0)
        self.sqrt_alphas_cumprod = torch.sqrt(self.alphas_cumprod)
        self.sqrt_one_minus_alphas_cumprod = torch.sqrt(1.0 - self.alphas_cumprod)

    def _get_beta_schedule(self, schedule_name: str, offset: float = 0.0) -> torch.Tensor:
        T = self.num_timesteps
        if schedule_name == "cosine":                                                                   # :205-209
            x = torch.linspace(0, T, T + 1, device=self.device)
            ac = torch.cos(((x / T) + 0.008 + offset) / 1.008 * torch.pi * 0.5) ** 2
            ac = ac / ac[0]
            return torch.clip(1 - (ac[1:] / ac[:-1]), 0.0001, 0.9999)
        if schedule_name == "linear":                                                                   # :210-211
            return torch.linspace(0.0001, 0.02, T, device=self.device)
        raise NotImplementedError(f"unknown beta schedule: {schedule_name}")

    def q_sample(self, x_start: torch.Tensor, t: torch.Tensor, noise: Optional[torch.Tensor] = None):
        """:214-219 -> (noisy points, the noise that was added)."""
        if noise is None:
            noise = torch.randn_like(x_start)
        t = torch.clamp(t, 0, self.num_timesteps - 1)
        a = self.sqrt_alphas_cumprod[t].view(-1, 1, 1)
        s = self.sqrt_one_minus_alphas_cumprod[t].view(-1, 1, 1)
        return a * x_start + s * noise, noise

    def _apply_geometric_constraints(self, points: torch.Tensor, target_range: float = 1.8) -> torch.Tensor:
        return torch.tanh(points / target_range) * target_range                                         # :221-222

    def _ddim_update(self, x, noise, alpha_t, alpha_prev, source_points=None):
        """One DDIM step (:253-260 / :289-297): predicted x0, optional pull towards the source, tanh range constraint."""
        pred_x0 = (x - torch.sqrt(1.0 - alpha_t) * noise) / (torch.sqrt(alpha_t) + 1e-8)
        if source_points is not None:
            pred_x0 = pred_x0 + 0.1 * (source_points - pred_x0)
        pred_x0 = self._apply_geometric_constraints(pred_x0)
        return torch.sqrt(alpha_prev) * pred_x0 + torch.sqrt(1.0 - alpha_prev) * noise

    @torch.no_grad()
    def guided_sample_loop(self, model, source_points: torch.Tensor, condition_points: torch.Tensor,
                           num_inference_steps: int = 50, guidance_scale: float = 7.5) -> torch.Tensor:
        """:225-261.  Classifier-free-guided DDIM: every step denoises the conditional and the unconditional copy of the
        current sample on its coarse (voxel-downsampled) points and interpolates the predicted noise back to all points."""
        device, shape = source_points.device, source_points.shape
        B = shape[0]
        hp = model.hierarchical_processor
        style_feat = model.style_encoder(hp.downsample(condition_points)[0])
        style_in = torch.cat([style_feat, torch.zeros_like(style_feat)])
        x = torch.randn(shape, device=device)
        timesteps_host = torch.linspace(self.num_timesteps - 1, 0, num_inference_steps).long()   # :236, computed on the CPU
        timesteps = timesteps_host.to(device)
        timesteps_host = timesteps_host.tolist()
        ac = self.alphas_cumprod
        one = torch.ones((), device=device)
        for i in range(num_inference_steps):
            t = timesteps[i]
            x_in = torch.cat([x, x])
            t_in = t.expand(2 * B).contiguous()
            x_coarse, x_indices = hp.downsample(x_in)
            noise_coarse = model.noise_predictor(x_coarse, t_in, style_in)
            both = hp.upsample_knn(noise_coarse, x_in, x_indices)
            cond, uncond = both.chunk(2)
            noise = uncond + guidance_scale * (cond - uncond)
            # the reference reads the previous timestep from the list while t > 0, else "no previous step" (:251-252)
            alpha_prev = ac[timesteps[i + 1]] if (i + 1 < num_inference_steps and timesteps_host[i] > 0) else one
            x = self._ddim_update(x, noise, ac[t], alpha_prev, source_points)
        return x

    @torch.no_grad()
    def guided_sample_loop_device(self, model, source_points: torch.Tensor, condition_points: torch.Tensor,
                                  num_inference_steps: int = 50, guidance_scale: float = 7.5,
                                  graph: bool = True, x_init: Optional[torch.Tensor] = None,
                                  timing: Optional[dict] = None) -> torch.Tensor:
        """The same CFG-guided DDIM loop (:225-261) with every step device-resident: voxel
        downsample, fused denoiser, 3-NN upsample and the DDIM update run without a host round trip
        (``downsample_device`` / ``upsample_knn_device``), so ONE step is captured into a CUDA graph and replayed
        ``num_inference_steps`` times (``graph=True``); the step index, timestep and the two alpha-bar values are read from
        device tables by an in-graph counter.  ``graph=False`` runs the identical step eagerly (the tests compare the two).
        The random thinning of the downsample draws from the CUDA default generator in both modes.  Requires the
        hierarchical path (more points than ``config.global_points``).  ``timing`` (optional dict) receives
        ``ms_per_step``: CUDA-event time of the step loop alone (no style encoder, warm-up or capture)."""
        device, shape = source_points.device, source_points.shape
        B, N, _ = shape
        hp = model.hierarchical_processor
        if not (source_points.is_cuda and N > hp.global_points):
            raise RuntimeError("guided_sample_loop_device: CUDA tensors with more points than config.global_points expected")
        cond = condition_points
        if cond.shape[1] > hp.global_points:
            cond = hp.downsample_device(cond)[0]
        style_feat = model.style_encoder(cond)
        style_in = torch.cat([style_feat, torch.zeros_like(style_feat)]).contiguous()
        ts = torch.linspace(self.num_timesteps - 1, 0, num_inference_steps).long()               # :236 (CPU)
        prev = [int(ts[i + 1]) if (i + 1 < num_inference_steps and int(ts[i]) > 0) else -1 for i in range(num_inference_steps)]
        ac = self.alphas_cumprod
        tab_t = ts.to(device)
        tab_at = ac[tab_t].contiguous()
        tab_ap = torch.stack([ac[p] if p >= 0 else torch.ones((), device=device) for p in prev]).contiguous()
        x0 = torch.randn(shape, device=device) if x_init is None else x_init.to(device).float()
        x = x0.clone()
        counter = torch.zeros(1, dtype=torch.long, device=device)
        source = source_points.float().contiguous()

        def step():
            t = tab_t.index_select(0, counter)                      # [1]
            at = tab_at.index_select(0, counter).view(())
            ap = tab_ap.index_select(0, counter).view(())
            x_in = torch.cat([x, x])
            x_coarse, x_indices = hp.downsample_device(x_in)
            noise_coarse = model.noise_predictor(x_coarse, t.expand(2 * B).contiguous(), style_in)
            both = hp.upsample_knn_device(noise_coarse, x_in, x_indices)
            nc, nu = both.chunk(2)
            noise = nu + guidance_scale * (nc - nu)
            x.copy_(self._ddim_update(x, noise, at, ap, source))
            counter.add_(1)

        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) if timing is not None else None

        def finish():
            if ev is not None:
                ev[1].record()
                ev[1].synchronize()
                timing["ms_per_step"] = ev[0].elapsed_time(ev[1]) / num_inference_steps
            return x

        if not graph:
            if ev is not None:
                ev[0].record()
            for _ in range(num_inference_steps):
                step()
            return finish()
        rng = torch.cuda.get_rng_state(device)
        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            step()                                                  # warm-up: library load, packing, allocator
        torch.cuda.current_stream(device).wait_stream(side)
        torch.cuda.synchronize(device)
        counter.zero_()
        x.copy_(x0)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            step()
        counter.zero_()
        x.copy_(x0)
        torch.cuda.set_rng_state(rng, device)                       # the replays start from the caller's generator state
        if ev is not None:
            ev[0].record()
        for _ in range(num_inference_steps):
            g.replay()
        return finish()

    @torch.no_grad()
    def ddim_sample_loop(self, model, shape, condition_points: torch.Tensor, num_inference_steps: int = 50) -> torch.Tensor:
        """:263-300 (unguided)."""
        device = condition_points.device
        x = torch.randn(shape, device=device)
        timesteps = torch.linspace(self.num_timesteps - 1, 0, num_inference_steps, dtype=torch.long, device=device)
        hier = shape[1] > model.config.global_points
        ac = self.alphas_cumprod
        one = torch.ones((), device=device)
        for i in range(num_inference_steps):
            t = timesteps[i]
            batch_t = t.expand(shape[0]).contiguous()
            if hier:
                noise_coarse, indices = model(x, batch_t, condition_points, cond_drop_prob=0, use_hierarchical=True)
                noise = model.hierarchical_processor.upsample_knn(noise_coarse, x, indices)
            else:
                noise, _ = model(x, batch_t, condition_points, cond_drop_prob=0, use_hierarchical=False)
            alpha_prev = ac[timesteps[i + 1]] if i + 1 < num_inference_steps else one
            x = self._ddim_update(x, noise, ac[t], alpha_prev)
        return x

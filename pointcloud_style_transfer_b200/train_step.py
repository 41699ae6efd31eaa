"""One optimisation step of the reference's trainer (training/trainer.py:64-125, the body of ``train_one_epoch``) with the
hot path on this package's kernels -- the unit BASELINE config 4 is quoted on ("full diffusion training step ... bf16,
batch 32 x 16384 points on 8 x B200").

What runs where:
  * q_sample, the CFG condition drop, the coarse gathers and the x0 prediction: the reference's tensor expressions (:75-101);
  * voxel-grid downsample of condition and noisy cloud: ``HierarchicalProcessor.downsample_device`` (device-resident);
  * style encoder (PointNet++ set abstraction x 3) forward AND backward: the native train-mode kernels
    (csrc/sa_mlp_train.cu, csrc/fps.cu, csrc/ball_query.cu) -- no cuDNN / cuBLAS;
  * Chamfer loss forward + backward: csrc/nn_min.cu;
  * the denoiser's nn.Linear stack and the two-layer style MLP: torch autograd under bf16 autocast (library GEMMs; the
    reference's own formulation, outside the hot-path scope -- SURVEY.md section 2 row 5);
  * gradient averaging across ranks: ONE NCCL all-reduce of a flat fp32 gradient buffer (2 549 827 parameters, 10.2 MB),
    the parameters' ``.grad`` tensors being views into it; then clip-by-norm 1.0, fused AdamW(0.9, 0.95) and the EMA update
    (:119-125).
"""
from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist

from .models.diffusion_model import DiffusionProcess, PointCloudDiffusionModel
from .models.losses import DiffusionLoss


def attach_flat_grad(params) -> torch.Tensor:
    """Make every parameter's ``.grad`` a view into ONE flat fp32 buffer (returned), so that the cross-rank gradient average
    is a single collective over 4 * numel bytes instead of one per tensor."""
    params = list(params)
    flat = torch.zeros(sum(p.numel() for p in params), dtype=torch.float32, device=params[0].device)
    off = 0
    for p in params:
        p.grad = flat[off: off + p.numel()].view_as(p)
        off += p.numel()
    return flat


def average_gradients(flat_grad: torch.Tensor, world: int) -> None:
    """DDP-style gradient averaging: all-reduce(SUM) of the flat buffer, then divide by the number of ranks."""
    if world > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
        flat_grad.div_(world)


class DiffusionTrainStep:
    def __init__(self, config, device, mlp_precision: int = 1, world: int = 1, amp_dtype: Optional[torch.dtype] = torch.bfloat16,
                 quiet: bool = True):
        self.config = config
        self.device = torch.device(device)
        self.world = world
        self.amp_dtype = amp_dtype if getattr(config, "use_amp", True) else None
        self.model = PointCloudDiffusionModel(config, mlp_precision=mlp_precision).to(self.device).train()
        self.model.hierarchical_processor.rng_device = "cuda"
        self.process = DiffusionProcess(config, device=str(self.device))
        if quiet:
            import contextlib
            import io
            with contextlib.redirect_stdout(io.StringIO()):   # the reference's DiffusionLoss prints its weights (:76-78)
                self.loss_fn = DiffusionLoss(noise_weight=1.0, chamfer_weight=config.lambda_chamfer)
        else:
            self.loss_fn = DiffusionLoss(noise_weight=1.0, chamfer_weight=config.lambda_chamfer)
        params = [p for p in self.model.parameters() if p.requires_grad]
        self.flat_grad = attach_flat_grad(params)   # every .grad is a view: the cross-rank average is one collective
        self.params = params
        self.optimizer = torch.optim.AdamW(params, lr=config.learning_rate, weight_decay=config.weight_decay, betas=(0.9, 0.95),
                                           fused=True)
        self.ema = [p.detach().clone() for p in params]
        self.ema_decay = getattr(config, "ema_decay", 0.999)
        self.clip = 1.0                                                                      # trainer.py:61

    def loss(self, sim_points: torch.Tensor, real_points: torch.Tensor, t: Optional[torch.Tensor] = None,
             noise: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, Dict[str, float]]:
        """trainer.py:71-113: forward of one batch -> (loss, loss_dict)."""
        cfg = self.config
        B, N, C = sim_points.shape
        if t is None:
            t = torch.randint(0, cfg.num_timesteps, (B,), device=self.device).long()
        noisy, actual_noise = self.process.q_sample(sim_points, t, noise)
        with torch.autocast(device_type="cuda", dtype=self.amp_dtype or torch.bfloat16, enabled=self.amp_dtype is not None):
            pred, indices = self.model(noisy_points=noisy, timestep=t, condition_points=real_points,
                                       cond_drop_prob=cfg.cond_drop_prob, use_hierarchical=cfg.use_hierarchical)
            if indices is not None:                                                          # hierarchical path, :90-108
                ind = indices.unsqueeze(-1).expand(-1, -1, C)
                noise_coarse = torch.gather(actual_noise, 1, ind)
                pred_x0 = sim_coarse = None
                if cfg.lambda_chamfer > 0:
                    noisy_coarse = torch.gather(noisy, 1, ind)
                    sim_coarse = torch.gather(sim_points, 1, ind)
                    a = self.process.sqrt_alphas_cumprod[t].view(B, 1, 1)
                    s = self.process.sqrt_one_minus_alphas_cumprod[t].view(B, 1, 1)
                    pred_x0 = (noisy_coarse - s * pred) / (a + 1e-8)
                return self.loss_fn(predicted_noise=pred, actual_noise=noise_coarse, predicted_points_coarse=pred_x0,
                                    target_points_coarse=sim_coarse)
            return self.loss_fn(predicted_noise=pred, actual_noise=actual_noise)             # direct path, :109-113

    def step(self, sim_points: torch.Tensor, real_points: torch.Tensor, t: Optional[torch.Tensor] = None,
             noise: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, Dict[str, float]]:
        """One batch: forward, backward, gradient average over the ranks, clip, AdamW, EMA (:78-125 with
        gradient_accumulation_steps = 1)."""
        loss, loss_dict = self.loss(sim_points, real_points, t, noise)
        loss.backward()
        average_gradients(self.flat_grad, self.world)
        torch.nn.utils.clip_grad_norm_(self.params, self.clip, foreach=True)
        self.optimizer.step()
        self.optimizer.zero_grad(set_to_none=False)
        torch._foreach_lerp_(self.ema, [p.detach() for p in self.params], 1.0 - self.ema_decay)
        return loss.detach(), loss_dict

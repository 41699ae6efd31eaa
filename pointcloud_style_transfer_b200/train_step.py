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

``step`` keeps the reference's host behaviour (the loss dictionary is read back with ``.item()`` three times per step,
FPS start indices drawn on the CPU generator inside the forward).  ``step_graphed`` is the same arithmetic as ONE CUDA
graph replay per step: a step is ~600 kernel launches of a few microseconds each and is bound by the host's launch rate
otherwise.  The graph holds forward, backward, the NCCL all-reduce, clip, AdamW and EMA; the timestep / noise / dropout
draws use the device generator as in the reference (:75-77, captured as graph-safe Philox offsets); the two FPS start
draws stay on the CPU generator (reference :36) and are copied into static buffers before each replay.
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
                                           fused=True, capturable=True)   # capturable: the step counter lives on the device
        self.ema = [p.detach().clone() for p in params]
        self.ema_decay = getattr(config, "ema_decay", 0.999)
        self.clip = 1.0                                                                      # trainer.py:61
        self._graph = None

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

    # ------------------------------------------------------------------ one CUDA graph per step
    def _device_step(self, sim_points: torch.Tensor, real_points: torch.Tensor) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
        """``step`` without host syncs: loss dictionary as device tensors."""
        self.loss_fn.sync_items = False
        try:
            loss, loss_dict = self.loss(sim_points, real_points)
        finally:
            self.loss_fn.sync_items = True
        loss.backward()
        average_gradients(self.flat_grad, self.world)
        torch.nn.utils.clip_grad_norm_(self.params, self.clip, foreach=True)
        self.optimizer.step()
        self.flat_grad.zero_()
        torch._foreach_lerp_(self.ema, [p.detach() for p in self.params], 1.0 - self.ema_decay)
        return loss.detach(), loss_dict

    def _capture(self, sim_points: torch.Tensor, real_points: torch.Tensor) -> None:
        enc = self.model.style_encoder.encoder
        B = sim_points.shape[0]
        n1 = min(real_points.shape[1], self.config.global_points) if self.config.use_hierarchical else real_points.shape[1]
        st = {"sim": sim_points.clone(), "real": real_points.clone(),
              "starts": torch.zeros(2, B, dtype=torch.long, device=self.device),
              # pinned staging ring: a slot is rewritten only after its copy has completed (the host runs ahead of replays)
              "starts_host": [torch.zeros(2, B, dtype=torch.long).pin_memory() for _ in range(8)],
              "copied": [None] * 8, "slot": 0, "n1": n1}
        enc.static_starts = (st["starts"][0], st["starts"][1])
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(3):                      # allocator, lazy initialisations, optimizer state
                self._draw_starts(st)
                self._device_step(st["sim"], st["real"])
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            st["loss"], st["loss_dict"] = self._device_step(st["sim"], st["real"])
        st["graph"] = g
        self._graph = st

    def _draw_starts(self, st: dict) -> None:
        """The reference's two FPS start draws of the style encoder (sa1 on the coarse condition cloud, then sa2), CPU
        default generator (models/pointnet2_encoder.py:36)."""
        enc = self.model.style_encoder.encoder
        slot = st["slot"]
        st["slot"] = (slot + 1) % len(st["starts_host"])
        if st["copied"][slot] is not None:
            st["copied"][slot].synchronize()
        host = st["starts_host"][slot]
        torch.randint(0, st["n1"], (host.shape[1],), dtype=torch.long, out=host[0])
        torch.randint(0, enc.sa1.npoint, (host.shape[1],), dtype=torch.long, out=host[1])
        st["starts"].copy_(host, non_blocking=True)
        ev = st["copied"][slot] = st["copied"][slot] or torch.cuda.Event()
        ev.record()

    def step_graphed(self, sim_points: torch.Tensor, real_points: torch.Tensor) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
        """One batch as one CUDA-graph replay.  Returns the graph's static (loss, loss_dict) device tensors (valid until the
        next call); read them with ``float()`` when a host value is needed."""
        if self._graph is None or self._graph["sim"].shape != sim_points.shape or self._graph["real"].shape != real_points.shape:
            self.release()
            self._capture(sim_points, real_points)
        st = self._graph
        if sim_points.data_ptr() != st["sim"].data_ptr():
            st["sim"].copy_(sim_points, non_blocking=True)
        if real_points.data_ptr() != st["real"].data_ptr():
            st["real"].copy_(real_points, non_blocking=True)
        self._draw_starts(st)
        st["graph"].replay()
        return st["loss"], st["loss_dict"]

    def release(self) -> None:
        """Drop the captured graph (it holds NCCL kernels when world > 1: must go before ``destroy_process_group``)."""
        if self._graph is not None:
            torch.cuda.synchronize(self.device)
            self.model.style_encoder.encoder.static_starts = None
            self._graph = None

"""CUDA-graph runner for the inference path of the encoder.

A 120k-point encoder forward is ~15 kernel launches of a few microseconds to a few hundred
microseconds each; launched one by one through Python the step is CPU-bound.  ``GraphedEncoder``
captures the whole eval-mode forward (FPS -> ball query -> fused gather+MLP+max, three stages) into
one CUDA graph per input shape and replays it with a single launch.  The FPS start indices are still
drawn from the CPU generator exactly like the reference (models/pointnet2_encoder.py:36) on every
call and copied into the graph's static buffer, so a seeded run is identical to the eager path.

A captured graph holds raw pointers to the packed (BatchNorm-folded) weight blobs of the three
stages.  Every call therefore compares the parameter version of the encoder (data pointers and
in-place version counters of all conv / BatchNorm tensors, plus the MLP precision) with the one
the graph was captured under; on a mismatch (``optimizer.step()``, ``load_state_dict``,
``set_mlp_precision`` ...) the graph is dropped and re-captured, and the blobs a graph reads are kept
alive in the graph's own state for as long as the graph exists.
"""
from typing import Dict, Tuple

import torch

from .models.pointnet2_encoder import PointNet2Encoder

_RING = 4  # pinned staging slots for the start indices: a slot is rewritten only after its copy has completed


class GraphedEncoder:
    def __init__(self, encoder: PointNet2Encoder, warmup: int = 2):
        if encoder.training:
            raise RuntimeError("GraphedEncoder captures the eval-mode (BatchNorm-folded) forward; call .eval() first")
        self.encoder = encoder
        self.warmup = warmup
        self._graphs: Dict[Tuple, dict] = {}

    def _param_key(self) -> Tuple:
        """Changes whenever a replay would read stale weights: any conv / BatchNorm tensor re-allocated or
        modified in place, or the MLP precision switched."""
        enc = self.encoder
        key = []
        for sa in (enc.sa1, enc.sa2, enc.sa3):
            key.append((int(sa.mlp_precision), sa._native_steps))
            for conv, bn in zip(sa.mlp_convs, sa.mlp_bns):
                for t in (conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var):
                    key.append((t.data_ptr(), t._version))
        return tuple(key)

    def _capture(self, B: int, N: int, device) -> dict:
        enc = self.encoder
        st = {
            "x": torch.zeros(B, N, 3, dtype=torch.float32, device=device),
            "starts": torch.zeros(2, B, dtype=torch.long, device=device),       # row 0: sa1, row 1: sa2
            "starts_host": [torch.zeros(2, B, dtype=torch.long).pin_memory() for _ in range(_RING)],
            "copied": [None] * _RING,
            "slot": 0,
        }
        s1, s2 = st["starts"][0], st["starts"][1]
        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(self.warmup):  # loads the library, sets function attributes, warms the allocator
                enc(st["x"], (s1, s2))
        torch.cuda.current_stream(device).wait_stream(side)
        torch.cuda.synchronize(device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g), torch.no_grad():
            st["out"] = enc(st["x"], (s1, s2))
        st["graph"] = g
        # the blobs the captured kernels read: owned by the graph state, not only by the modules' caches
        st["blobs"] = [dict(sa._packed) for sa in (enc.sa1, enc.sa2, enc.sa3)]
        st["param_key"] = self._param_key()
        return st

    def _state(self, B: int, N: int) -> dict:
        """The captured graph for shape (B, N), (re-)captured if absent or stale."""
        device = next(self.encoder.parameters()).device
        key = (B, N, device)
        st = self._graphs.get(key)
        if st is not None and st["param_key"] != self._param_key():
            torch.cuda.synchronize(device)  # replays of the stale graph still in flight read the old blobs
            del self._graphs[key]
            st = None
        if st is None:
            with torch.cuda.device(device), torch.no_grad():
                st = self._graphs[key] = self._capture(B, N, device)
        return st

    @torch.no_grad()
    def __call__(self, xyz: torch.Tensor) -> torch.Tensor:
        """xyz [B,N,3]: a CUDA tensor, or a (preferably pinned) CPU tensor that is copied asynchronously.
        Returns the graph's static output buffer [B,feature_dim] (valid until the next call)."""
        if self.encoder.training:
            raise RuntimeError("GraphedEncoder replays the eval-mode forward; the encoder was switched to train()")
        B, N, _ = xyz.shape
        st = self._state(B, N)
        slot = st["slot"]
        st["slot"] = (slot + 1) % _RING
        if st["copied"][slot] is not None:
            st["copied"][slot].synchronize()  # _RING calls ago: long done, this never blocks in practice
        host = st["starts_host"][slot]
        # the reference's two start draws (sa1 then sa2), CPU default generator
        torch.randint(0, N, (B,), dtype=torch.long, out=host[0])
        torch.randint(0, self.encoder.sa1.npoint, (B,), dtype=torch.long, out=host[1])
        st["starts"].copy_(host, non_blocking=True)
        ev = st["copied"][slot] = st["copied"][slot] or torch.cuda.Event()
        ev.record()
        if xyz.data_ptr() != st["x"].data_ptr():
            st["x"].copy_(xyz, non_blocking=True)
        st["graph"].replay()
        return st["out"]

    def static_input(self, B: int, N: int) -> torch.Tensor:
        """The graph's own input buffer for shape (B, N): callers that produce the scan on the device can write
        into it directly and pass it to ``__call__``, which then skips the device-to-device copy."""
        return self._state(B, N)["x"]

"""Drop-in for ``HierarchicalProcessor.upsample_knn`` of the reference's models/diffusion_model.py.

Only the 3-NN inverse-distance upsample (the north-star's "feature-propagation interpolation") is
on the hot path; the voxel-grid downsample is a "next" row and is not built yet."""
import torch

from .. import ops


class HierarchicalProcessor:
    """models/diffusion_model.py:64-153 (upsample_knn only)."""

    def __init__(self, total_points: int = 120000, global_points: int = 30000):
        self.total_points = total_points
        self.global_points = global_points

    def downsample(self, points: torch.Tensor):
        raise NotImplementedError("voxel-grid downsample (models/diffusion_model.py:69-125) is a 'next' row of "
                                  "SURVEY.md §8(f); use the reference's implementation for it")

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

"""Drop-in for ``HierarchicalProcessor`` of the reference's models/diffusion_model.py: the 3-NN
inverse-distance upsample (the north-star's "feature-propagation interpolation", §8(a) row a12) and the
voxel-grid downsample in front of the encoder (first "next" row)."""
from typing import Tuple

import torch

from .. import ops


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
        """models/diffusion_model.py:124-125."""
        return self._voxel_grid_downsample_torch(points, self.global_points)

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

"""Drop-in for ``calculate_similarity`` of the reference's ``compare.py`` (:6-43) on B200 -- SURVEY.md §8(f) rank 3.

The reference builds two scipy cKD-trees and thresholds the fp64 1-NN distances in both directions.  Here the two
1-NN searches run on the kNN kernel (fp32 prefilter, exact fp64 ranking in cKDTree's summation order), so the
precision / recall / F1 values are identical for float32 point clouds."""
from typing import Tuple, Union

import numpy as np
import torch

from . import ops

ArrayLike = Union[np.ndarray, torch.Tensor]


def _to_cuda_f32(p: ArrayLike, device) -> torch.Tensor:
    t = torch.as_tensor(p)
    if t.dtype == torch.float64:
        t32 = t.float()
        if not torch.equal(t32.double(), t):
            raise ValueError("calculate_similarity: float64 coordinates that are not exactly representable in float32 "
                             "are not supported by the fp32-input kernels (the reference searches them in fp64)")
        t = t32
    return t.float().reshape(1, -1, 3).to(device)


def calculate_similarity(pcd1: ArrayLike, pcd2: ArrayLike, threshold: float,
                         device: str = "cuda") -> Tuple[float, float, float]:
    """compare.py:6-43.  pcd1 [N,3] (reference / ground truth), pcd2 [M,3] (generated) ->
    (precision %, recall %, F1): the share of pcd2 points whose nearest pcd1 point is closer than ``threshold``,
    the share of pcd1 points whose nearest pcd2 point is, and their harmonic mean."""
    if not torch.cuda.is_available():
        raise RuntimeError("calculate_similarity: a CUDA device (B200) is required; there is no CPU fallback")
    a, b = _to_cuda_f32(pcd1, device), _to_cuda_f32(pcd2, device)
    d21, _ = ops.knn(b, a, 1)   # every pcd2 point -> nearest pcd1 point (fp64 distances)
    d12, _ = ops.knn(a, b, 1)
    # np.mean of a boolean array = exact integer count / n
    precision = int((d21[0, :, 0] < threshold).sum().item()) / d21.shape[1]
    recall = int((d12[0, :, 0] < threshold).sum().item()) / d12.shape[1]
    f_score = 0.0 if precision + recall == 0 else 2 * (precision * recall) / (precision + recall)
    return precision * 100, recall * 100, f_score

"""Host-side mirror of the reference's ``models`` package for the hot path (same module, class and
function names as models/pointnet2_encoder.py, models/losses.py and the HierarchicalProcessor of
models/diffusion_model.py)."""
from .pointnet2_encoder import (PointNet2Encoder, SetAbstraction, farthest_point_sample, index_points,  # noqa: F401
                                query_ball_point, square_distance)
from .losses import DiffusionLoss, chamfer_distance_chunked_optimized  # noqa: F401
from .diffusion_model import HierarchicalProcessor  # noqa: F401

"""pointcloud_style_transfer_b200 -- B200 (sm_100a) kernels behind the call surface of the point-set
hot path of wangxy0820/PointCloud_style_transfer.

    from pointcloud_style_transfer_b200.models.pointnet2_encoder import PointNet2Encoder, ...
    from pointcloud_style_transfer_b200.models.losses import chamfer_distance_chunked_optimized, DiffusionLoss
    from pointcloud_style_transfer_b200.evaluation.metrics import PointCloudMetrics
    from pointcloud_style_transfer_b200.models.diffusion_model import HierarchicalProcessor

or ``patch_reference()`` to swap these into an unmodified checkout of the reference.
The compute lives in ``csrc/libpcst.so`` (C ABI: ``include/pcst.h``); importing this package does
not load it, the first op call does and raises if it is missing -- there is no fallback path.
"""
import sys

__version__ = "0.1.0"


def patch_reference() -> None:
    """Replace the hot-path symbols of an importable reference checkout (``models.pointnet2_encoder``,
    ``models.losses``, ``evaluation.metrics`` and ``HierarchicalProcessor.upsample_knn``) with the
    B200 implementations (plus the two "next" rows that are built: ``HierarchicalProcessor.downsample`` and
    ``compare.calculate_similarity``), so that the reference's trainer / inference scripts run unmodified."""
    import importlib

    from .evaluation import metrics as our_metrics
    from .models import diffusion_model as our_dm
    from .models import losses as our_losses
    from .models import pointnet2_encoder as our_enc

    ref_enc = importlib.import_module("models.pointnet2_encoder")
    for name in ("square_distance", "index_points", "farthest_point_sample", "query_ball_point",
                 "SetAbstraction", "PointNet2Encoder"):
        setattr(ref_enc, name, getattr(our_enc, name))
    ref_losses = importlib.import_module("models.losses")
    for name in ("chamfer_distance_chunked_optimized", "DiffusionLoss"):
        setattr(ref_losses, name, getattr(our_losses, name))
    ref_dm = importlib.import_module("models.diffusion_model")
    ref_dm.PointNet2Encoder = our_enc.PointNet2Encoder
    ref_dm.HierarchicalProcessor.upsample_knn = our_dm.HierarchicalProcessor.upsample_knn
    for name in ("rng_device", "_randperm", "_voxel_grid_downsample_torch", "downsample"):
        setattr(ref_dm.HierarchicalProcessor, name, getattr(our_dm.HierarchicalProcessor, name))
    if "compare" in sys.modules:
        from . import compare as our_compare
        sys.modules["compare"].calculate_similarity = our_compare.calculate_similarity
    if "evaluation.metrics" in sys.modules:
        sys.modules["evaluation.metrics"].PointCloudMetrics = our_metrics.PointCloudMetrics
    for mod in ("training.trainer", "scripts.test"):
        if mod in sys.modules:
            m = sys.modules[mod]
            if hasattr(m, "DiffusionLoss"):
                m.DiffusionLoss = our_losses.DiffusionLoss
            if hasattr(m, "PointCloudMetrics"):
                m.PointCloudMetrics = our_metrics.PointCloudMetrics

"""Host-side mirror of the reference's ``evaluation.metrics`` for the NN-reduction metrics."""
from .metrics import PointCloudMetrics  # noqa: F401

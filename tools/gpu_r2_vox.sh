set -x
timeout 600 python -m pytest tests -m gpu -x -q -k "voxel or downsample or sampling" 2>&1 | tail -4
timeout 200 python - <<'PY'
import torch, sys, os
sys.path.insert(0, os.getcwd())
from pointcloud_style_transfer_b200 import ops, synthetic as S
from pointcloud_style_transfer_b200.models import diffusion_model as DM
dev = torch.device("cuda:0")
x = S.lidar_scan(0).to(dev)
hp = DM.HierarchicalProcessor(30000)
hp.rng_device = "cuda"
def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
print("downsample_device 120k -> 30k: %.3f ms" % timeit(lambda: hp.downsample_device(x)))
log = ops.timed_calls(lambda: hp.downsample_device(x)) if hasattr(ops, "timed_calls") else None
print(log)
PY

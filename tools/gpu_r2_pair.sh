set -x
timeout 900 python -m pytest tests -m gpu -x -q -k "chamfer or nn_min or hausdorff or metrics or sharded or loss" 2>&1 | tail -3
timeout 200 python - <<'PY'
import torch, sys, os
sys.path.insert(0, os.getcwd())
from pointcloud_style_transfer_b200 import ops, synthetic as S
dev = torch.device("cuda:0")
x, y = S.lidar_scan(0).to(dev), S.lidar_scan(100).to(dev)
def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
with torch.no_grad():
    for n in (120000, 60000, 30000, 15000, 7500):
        p = x[:, :n].contiguous()
        t = timeit(lambda: ops.nn_min_pair(p, y, 0))
        print(f"nn_min_pair {n} x 120000: {t*1e3:.1f} us  ({n*120000/t/1e9:.2f} e12 pairs/s; full-size rate would give {3014*n/120000:.1f} us)", flush=True)
PY

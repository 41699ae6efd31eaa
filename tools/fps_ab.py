import sys, os
sys.path.insert(0, os.getcwd())
import torch
from pointcloud_style_transfer_b200 import ops, synthetic as S, _lib
dev=torch.device("cuda:0")
x=S.lidar_scan(0).to(dev); start=torch.tensor([1234],device=dev)
def timeit(name, fn, reps=10):
    fn(); torch.cuda.synchronize()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    print(f"{name}: {a.elapsed_time(b)/reps*1e3:.1f} us", flush=True)
for mode,label in ((2,"one sample per exchange (old kernel)"),(1,"look-ahead"),(3,"look-ahead machinery, runner-up never accepted (511 rounds)"),(4,"look-ahead, every chunk skipped (invalid), rounds vary")):
    _lib.set_tuning("fps.lookahead", mode)
    timeit("fps 120k->512 "+label, lambda: ops.fps(x,512,start))
_lib.set_tuning("fps.lookahead", 0)

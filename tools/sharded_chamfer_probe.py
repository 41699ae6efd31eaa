"""Query-sharded Chamfer (config 5) at N GPUs: where the call's time goes.  Run under torchrun; rank 0 prints.
Phases are timed eagerly with CUDA events (a barrier + synchronize before each repetition, max over ranks); the graph
replay of the whole call is timed the same way."""
import os
import sys

sys.path.insert(0, os.getcwd())
import torch
import torch.distributed as dist

from pointcloud_style_transfer_b200 import distributed as D, ops, synthetic as S

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
x, y = S.lidar_scan(0).to(dev), S.lidar_scan(100).to(dev)
N = x.shape[1]
lo, hi = D.slice_of_rank(N, world, rank)
p_loc, t_loc = x[:, lo:hi].contiguous(), y[:, lo:hi].contiguous()


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    tot = 0.0
    for _ in range(reps):
        dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    t = torch.tensor([tot / reps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


with torch.no_grad():
    t_all = D.all_gather_ragged(t_loc, None, N)
    rowmin, colmin = ops.nn_min_pair(p_loc, t_all, 0)
    payload = ops.chamfer_shard_pack(rowmin, colmin)
    gathered = payload.new_empty(world * payload.shape[0], payload.shape[1])
    res = {
        "all_gather target": timed(lambda: D.all_gather_ragged(t_loc, None, N)),
        "sweep n/G x M": timed(lambda: ops.nn_min_pair(p_loc, t_all, 0)),
        "pack": timed(lambda: ops.chamfer_shard_pack(rowmin, colmin)),
        "all_gather payload": timed(lambda: dist.all_gather_into_tensor(gathered, payload)),
        "finish": timed(lambda: ops.chamfer_shard_finish(gathered.view(world, payload.shape[0], -1), N, 0)),
        "eager call": timed(lambda: D.chamfer_query_sharded_fused(p_loc, t_loc, N, N)),
    }
    g = D.GraphedShardedChamfer(N, N)
    res["graph replay"] = timed(lambda: g(p_loc, t_loc))
    tiny = torch.zeros(8, device=dev)
    big = torch.zeros(world * 8, device=dev)
    res["all_gather 32 B (latency floor)"] = timed(lambda: dist.all_gather_into_tensor(big, tiny))
    res["all_reduce MIN 480 KB"] = timed(lambda: dist.all_reduce(colmin, op=dist.ReduceOp.MIN))
    g.release()
if rank == 0:
    tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("NCCL_") and k not in ("NCCL_DEBUG",))
    print(f"world={world} [{tag or 'default NCCL settings'}]")
    for k, v in res.items():
        print(f"  {k}: {v * 1e3:.1f} us")
torch.cuda.synchronize()
dist.destroy_process_group()

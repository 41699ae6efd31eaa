"""CPU, world_size 2, gloo: the host-side sharding logic of pointcloud_style_transfer_b200.distributed
(ragged all-gather, query-sharded Chamfer, scan-sharded encoding) with the CPU oracle injected as the
local compute.  The CUDA kernels are not involved here; the same functions run on NCCL in production."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pointcloud_style_transfer_b200 import distributed as D
from pointcloud_style_transfer_b200 import synthetic as S


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_nn_min(a, b, form):
    from oracle import ref_oracle as O
    return torch.from_numpy(O.nn_min(a.numpy(), b.numpy(), form))


def _oracle_pair(a, b, form):
    from oracle import ref_oracle as O
    rows = O.nn_min(a.numpy(), b.numpy(), 0 if form == 0 else 1)
    cols = O.nn_min(b.numpy(), a.numpy(), 0 if form == 0 else 2)
    return torch.from_numpy(rows), torch.from_numpy(cols)


def _oracle_knn(q, r, k):
    from oracle import ref_oracle as O
    d, i = O.knn(q.numpy(), r.numpy(), k)
    return torch.from_numpy(d), torch.from_numpy(i)


_PARTS = 128  # fp64 partial row sums per payload row (pcst_chamfer_shard_payload_floats(M) - M = 256 float slots)


def _cpu_pack(rowmin, colmin):
    """CPU stand-in of pcst_chamfer_shard_pack_f32: colmin | 32 fp64 partial row sums as float pairs."""
    B, M = colmin.shape
    out = np.zeros((B, M + 2 * _PARTS), np.float32)
    out[:, :M] = colmin.numpy()
    parts = np.zeros((B, _PARTS), np.float64)
    parts[:, 0] = rowmin.numpy().astype(np.float64).sum(axis=1) if rowmin.shape[1] else 0.0
    out[:, M:] = parts.view(np.float32)
    return torch.from_numpy(out)


def _cpu_finish(gathered, n_total, form):
    g = gathered.numpy()
    M = g.shape[2] - 2 * _PARTS
    cols = g[:, :, :M].min(axis=0).astype(np.float64).sum(axis=1)
    rows = np.ascontiguousarray(g[:, :, M:]).view(np.float64).sum(axis=(0, 2))
    v = rows / n_total + cols / M
    return torch.from_numpy((v / 2 if form else v).astype(np.float32))


class _ToyEncoder(torch.nn.Module):
    def forward(self, x):  # [S,N,3] -> [S,4]: any per-scan function will do for the plumbing test
        return torch.cat([x.mean(dim=1), x.abs().amax(dim=(1, 2))[:, None]], dim=1)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pred, target = S.uniform_cloud(0, 2, 1001), S.uniform_cloud(100, 2, 777)  # ragged over 2 ranks
        lo, hi = D.slice_of_rank(1001, world, rank)
        lo2, hi2 = D.slice_of_rank(777, world, rank)
        g = D.all_gather_ragged(pred[:, lo:hi].contiguous())
        ok_gather = torch.equal(g, pred)
        cd = D.chamfer_query_sharded(pred[:, lo:hi].contiguous(), target[:, lo2:hi2].contiguous(), nn_min_fn=_oracle_nn_min)
        cdm = D.chamfer_query_sharded(pred[:, lo:hi].contiguous(), target[:, lo2:hi2].contiguous(), nn_min_fn=_oracle_nn_min, form=1)
        cd1 = D.chamfer_query_sharded_one_sweep(pred[:, lo:hi].contiguous(), target[:, lo2:hi2].contiguous(), pair_fn=_oracle_pair)
        cdm1 = D.chamfer_query_sharded_one_sweep(pred[:, lo:hi].contiguous(), target[:, lo2:hi2].contiguous(), pair_fn=_oracle_pair, form=1,
                                                 pred_total=1001, target_total=777)  # known totals: no size exchange
        # the fused variant: one result collective (all-gather of packed payloads), pack / finish injected
        cdf = D.chamfer_query_sharded_fused(pred[:, lo:hi].contiguous(), target[:, lo2:hi2].contiguous(), 1001, 777,
                                            pair_fn=_oracle_pair, pack_fn=_cpu_pack, finish_fn=_cpu_finish)
        cdfm = D.chamfer_query_sharded_fused(pred[:, lo:hi].contiguous(), target[:, lo2:hi2].contiguous(), 1001, 777, form=1,
                                             pair_fn=_oracle_pair, pack_fn=_cpu_pack, finish_fn=_cpu_finish)
        # a rank with an EMPTY query slice (more ranks than points would do this): rank 1 holds no pred points
        elo, ehi = (0, 5) if rank == 0 else (5, 5)
        cde = D.chamfer_query_sharded_one_sweep(pred[:, :5][:, elo:ehi].contiguous(), target[:, lo2:hi2].contiguous(), pair_fn=_oracle_pair)
        even = S.uniform_cloud(9, 3, 64)                       # equal shards take the single-buffer all-gather
        elo, ehi = D.slice_of_rank(64, world, rank)
        ok_gather = ok_gather and torch.equal(D.all_gather_ragged(even[:, elo:ehi].contiguous(), total=64), even)
        scans = S.uniform_cloud(5, 5, 64)
        feats = D.encode_scans_sharded(_ToyEncoder(), scans)
        # 3-NN with the queries sharded and the (ragged) reference shards all-gathered: results stay sharded
        kd, ki = D.knn_query_sharded(pred[:, lo:hi].contiguous(), target[:, lo2:hi2].contiguous(), 3, knn_fn=_oracle_knn)
        # DDP-style gradient averaging of the config-4 training step: every .grad a view into one flat buffer, one all-reduce
        from pointcloud_style_transfer_b200.train_step import attach_flat_grad, average_gradients
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Linear(3, 5), torch.nn.ReLU(), torch.nn.Linear(5, 2))
        flat = attach_flat_grad(net.parameters())
        net(scans[rank]).square().sum().backward()          # rank-dependent batch
        views_ok = all(p.grad.data_ptr() >= flat.data_ptr() for p in net.parameters())
        average_gradients(flat, world)
        q.put((rank, ok_gather, cd.numpy(), cdm.numpy(), feats.numpy(), cd1.numpy(), cdm1.numpy(), cde.numpy(),
               (lo, hi), kd.numpy(), ki.numpy(), views_ok, flat.numpy().copy(), cdf.numpy(), cdfm.numpy()))
    finally:
        dist.destroy_process_group()


def test_query_sharded_chamfer_and_scan_sharding_world2(oracle):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    pred, target = S.uniform_cloud(0, 2, 1001), S.uniform_cloud(100, 2, 777)
    ref = oracle.chamfer_distance_chunked_optimized(pred.numpy(), target.numpy())
    refm = oracle.metrics_chamfer_distance(pred.numpy(), target.numpy())
    scans = S.uniform_cloud(5, 5, 64)
    ref_feats = _ToyEncoder()(scans).numpy()
    ref_e = oracle.chamfer_distance_chunked_optimized(pred.numpy()[:, :5], target.numpy())
    ref_kd, ref_ki = oracle.knn(pred.numpy(), target.numpy(), 3)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(3, 5), torch.nn.ReLU(), torch.nn.Linear(5, 2))
    want = 0
    for r in range(2):
        net.zero_grad()
        net(scans[r]).square().sum().backward()
        want = want + torch.cat([p.grad.reshape(-1) for p in net.parameters()]).numpy() / 2
    for rank, ok_gather, cd, cdm, feats, cd1, cdm1, cde, (lo, hi), kd, ki, views_ok, flat, cdf, cdfm in res:
        assert views_ok
        np.testing.assert_allclose(cdf, ref, rtol=1e-6)     # fused: all-gather of payloads, min / sums in one finish step
        np.testing.assert_allclose(cdfm, refm, rtol=1e-6)
        np.testing.assert_allclose(flat, want, rtol=1e-6, atol=1e-7)   # both ranks hold the average of the two gradients
        np.testing.assert_array_equal(ki, ref_ki[:, lo:hi])   # indices refer to the gathered (rank-ordered) references
        np.testing.assert_array_equal(kd, ref_kd[:, lo:hi])
        assert ok_gather
        np.testing.assert_allclose(cd, ref, rtol=1e-6)
        np.testing.assert_allclose(cdm, refm, rtol=1e-6)
        np.testing.assert_allclose(cd1, ref, rtol=1e-6)     # one sweep + MIN all-reduce of the column minima
        np.testing.assert_allclose(cdm1, refm, rtol=1e-6)
        np.testing.assert_allclose(cde, ref_e, rtol=1e-6)   # empty slice on one rank
        np.testing.assert_array_equal(feats, ref_feats)


def test_partition_helpers():
    assert D.scans_of_rank(8, 8, 3) == [3]
    assert D.scans_of_rank(5, 2, 1) == [1, 3]
    covered = []
    for r in range(8):
        lo, hi = D.slice_of_rank(120001, 8, r)
        covered += list(range(lo, hi))
    assert covered == list(range(120001))
    assert D.slice_of_rank(3, 8, 7) == (3, 3)  # more ranks than points: empty slice

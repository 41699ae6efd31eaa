"""Multi-GPU sharding of the hot path: one process per GPU, ``torch.distributed`` (NCCL over NVLink).

The path shards two ways, both with tiny exchanges:

* batch of scans -- scan ``b`` lives on rank ``b mod G``; the encoder (FPS, ball query, grouping, MLP in
  eval mode) needs no data-path collective at all; only per-scan results are gathered.
* query points of one scan (Chamfer / kNN sweep) -- each rank owns a contiguous slice of the queries of
  both clouds, all-gathers the candidate sets (2 x 1.44 MB for 120k-point scans), runs the local NN-min
  and all-reduces ``2*B`` partial sums.

The local compute is injected (``nn_min_fn`` / ``knn_fn``): the product default is the CUDA op; the
world_size-2 gloo tests on CPU inject the oracle to exercise exactly this host logic.
"""
from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist


def scans_of_rank(num_scans: int, world: int, rank: int) -> List[int]:
    """Round-robin ownership: scan b -> rank b mod world."""
    return list(range(rank, num_scans, world))


def slice_of_rank(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi) slice of n query points for ``rank`` (ragged when world does not divide n)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_ragged(x: torch.Tensor, group=None, total: Optional[int] = None) -> torch.Tensor:
    """All-gather [B, n_r, C] shards with different n_r along dim 1 -> [B, sum n_r, C] in rank order.
    ``total`` (optional): the global point count when the shards are the ``slice_of_rank`` partition of it --
    the shard sizes are then known without a size exchange (no host synchronisation)."""
    world = dist.get_world_size(group)
    if world == 1:
        return x
    if total is not None:
        sizes = [hi - lo for lo, hi in (slice_of_rank(total, world, r) for r in range(world))]
        if sizes[dist.get_rank(group)] != x.shape[1]:
            raise ValueError("all_gather_ragged: the local shard is not the slice_of_rank partition of `total`")
    else:
        n = torch.tensor([x.shape[1]], dtype=torch.long, device=x.device)
        sizes = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(sizes, n, group=group)
        sizes = [int(s.item()) for s in sizes]
    if min(sizes) == max(sizes):  # equal shards: one collective into one buffer, no padding, no per-rank slicing
        buf = x.new_empty(world * x.shape[0], x.shape[1], x.shape[2])   # ranks concatenated along dim 0
        dist.all_gather_into_tensor(buf, x.contiguous(), group=group)
        buf = buf.view(world, x.shape[0], x.shape[1], x.shape[2])
        return buf.permute(1, 0, 2, 3).reshape(x.shape[0], world * x.shape[1], x.shape[2])
    nmax = max(sizes)
    pad = x.new_zeros(x.shape[0], nmax, x.shape[2])
    pad[:, : x.shape[1]] = x
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad.contiguous(), group=group)
    return torch.cat([b[:, :s] for b, s in zip(bufs, sizes)], dim=1).contiguous()


def _default_nn_min(a, b, form):
    from . import ops
    return ops.nn_min(a, b, form, False)[0]


def chamfer_query_sharded(pred_local: torch.Tensor, target_local: torch.Tensor, group=None,
                          nn_min_fn: Optional[Callable] = None, form: int = 0) -> torch.Tensor:
    """Bidirectional Chamfer of clouds whose points are sharded over the ranks of ``group``.

    pred_local [B,n_r,3], target_local [B,m_r,3] -> [B] (identical on every rank):
    ``form=0``: models/losses.py:61 (sum of the two means of clamped squared distances);
    ``form=1``: evaluation/metrics.py:42 ((mean + mean) / 2 of Euclidean distances)."""
    nn_min_fn = nn_min_fn or _default_nn_min
    pred_all = all_gather_ragged(pred_local, group)
    target_all = all_gather_ragged(target_local, group)
    f1, f2 = (0, 0) if form == 0 else (1, 2)
    d1 = nn_min_fn(pred_local, target_all, f1)       # my queries of pred against ALL of target
    d2 = nn_min_fn(target_local, pred_all, f2)       # my queries of target against ALL of pred
    sums = torch.stack([d1.double().sum(dim=1), d2.double().sum(dim=1)], dim=1)  # [B,2] partial sums
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    out = sums[:, 0] / pred_all.shape[1] + sums[:, 1] / target_all.shape[1]
    if form != 0:
        out = out / 2
    return out.float()


def _default_nn_min_pair(a, b, form):
    from . import ops
    return ops.nn_min_pair(a, b, form)


def chamfer_query_sharded_one_sweep(pred_local: torch.Tensor, target_local: torch.Tensor, group=None,
                                    pair_fn: Optional[Callable] = None, form: int = 0,
                                    pred_total: Optional[int] = None, target_total: Optional[int] = None) -> torch.Tensor:
    """Same result as ``chamfer_query_sharded`` with every pair evaluated ONCE across the whole job
    (SURVEY.md §8(e) variant): rank r sweeps its [n_r x M] tile of the pair matrix, which yields complete
    row minima for its pred points and PARTIAL column minima for all M target points; one
    ``all_reduce(MIN)`` over the [B,M] column minima (480 KB for a 120k-point scan) completes them.
    Exchange: all-gather of the target cloud (if it starts sharded), MIN all-reduce of M floats, SUM
    all-reduce of B partial row sums.  An empty local slice contributes +inf column minima.
    ``pred_total`` / ``target_total``: the global point counts when the shards are ``slice_of_rank``
    partitions; with them the call issues no size exchange and no host synchronisation."""
    pair_fn = pair_fn or _default_nn_min_pair
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    target_all = all_gather_ragged(target_local, group, target_total) if world > 1 else target_local
    B, M = target_all.shape[0], target_all.shape[1]
    if pred_local.shape[1] > 0:
        rowmin, colmin = pair_fn(pred_local, target_all, 0 if form == 0 else 1)
        rowsum = rowmin.double().sum(dim=1)
    else:
        colmin = torch.full((B, M), float("inf"), dtype=torch.float32, device=target_all.device)
        rowsum = torch.zeros(B, dtype=torch.float64, device=target_all.device)
    n_total = pred_total
    if world > 1:
        colmin = colmin.contiguous()
        dist.all_reduce(colmin, op=dist.ReduceOp.MIN, group=group)
        dist.all_reduce(rowsum, op=dist.ReduceOp.SUM, group=group)
        if n_total is None:
            cnt = torch.tensor([pred_local.shape[1]], dtype=torch.long, device=target_all.device)
            dist.all_reduce(cnt, op=dist.ReduceOp.SUM, group=group)
            n_total = int(cnt.item())
    elif n_total is None:
        n_total = pred_local.shape[1]
    out = rowsum / n_total + colmin.double().sum(dim=1) / M
    if form != 0:
        out = out / 2
    return out.float()


def _default_pack(rowmin, colmin):
    from . import ops
    return ops.chamfer_shard_pack(rowmin, colmin)


def _default_finish(gathered, n_total, form):
    from . import ops
    return ops.chamfer_shard_finish(gathered, n_total, form)


def chamfer_query_sharded_fused(pred_local: torch.Tensor, target_local: torch.Tensor, pred_total: int, target_total: int,
                                group=None, form: int = 0, pair_fn: Optional[Callable] = None,
                                pack_fn: Optional[Callable] = None, finish_fn: Optional[Callable] = None) -> torch.Tensor:
    """``chamfer_query_sharded_one_sweep`` with ONE result collective and no host-side arithmetic: all-gather of the target
    cloud -> one sweep of the local [n_r x M] tile -> pack (column minima | fp64 row sum) -> all-gather of the payloads
    -> finish (min over ranks, fp64 sums, the two means) -> [B], identical on every rank.  Shards must be the
    ``slice_of_rank`` partitions of ``pred_total`` / ``target_total`` (no size exchange, no synchronisation): the whole
    call is stream-ordered and can be captured in a CUDA graph (``GraphedShardedChamfer``)."""
    pair_fn, pack_fn, finish_fn = pair_fn or _default_nn_min_pair, pack_fn or _default_pack, finish_fn or _default_finish
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    target_all = all_gather_ragged(target_local, group, target_total) if world > 1 else target_local
    B, M = target_all.shape[0], target_all.shape[1]
    if pred_local.shape[1] > 0:
        rowmin, colmin = pair_fn(pred_local, target_all, 0 if form == 0 else 1)
    else:   # an empty query slice: no row minima, column minima at +inf
        rowmin = target_all.new_zeros(B, 0)
        colmin = torch.full((B, M), float("inf"), dtype=torch.float32, device=target_all.device)
    payload = pack_fn(rowmin, colmin)                                   # [B, P]: column minima | fp64 row-sum partials
    P = payload.shape[1]
    if world > 1:
        gathered = payload.new_empty(world * B, P)                     # ranks concatenated along dim 0
        dist.all_gather_into_tensor(gathered, payload.contiguous(), group=group)
        gathered = gathered.view(world, B, P)
    else:
        gathered = payload[None]
    return finish_fn(gathered, pred_total, form)


class GraphedShardedChamfer:
    """The fused query-sharded Chamfer captured into ONE CUDA graph per shape (both collectives included): a call is two
    device-to-device copies of the local shards into the graph's static buffers and a graph launch, so that the ~20 kernel /
    collective launches of the call no longer cost host time between them."""

    def __init__(self, pred_total: int, target_total: int, group=None, form: int = 0):
        self.pred_total, self.target_total, self.group, self.form = pred_total, target_total, group, form
        self._state = {}

    def __call__(self, pred_local: torch.Tensor, target_local: torch.Tensor) -> torch.Tensor:
        key = (tuple(pred_local.shape), tuple(target_local.shape), pred_local.device)
        st = self._state.get(key)
        if st is None:
            st = {"p": torch.empty_like(pred_local), "t": torch.empty_like(target_local)}
            st["p"].copy_(pred_local)
            st["t"].copy_(target_local)
            run = lambda: chamfer_query_sharded_fused(st["p"], st["t"], self.pred_total, self.target_total, self.group, self.form)
            side = torch.cuda.Stream(device=pred_local.device)
            side.wait_stream(torch.cuda.current_stream(pred_local.device))
            with torch.cuda.stream(side), torch.no_grad():
                for _ in range(3):
                    run()
            torch.cuda.current_stream(pred_local.device).wait_stream(side)
            torch.cuda.synchronize(pred_local.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g), torch.no_grad():
                st["out"] = run()
            st["graph"] = g
            self._state[key] = st
        st["p"].copy_(pred_local, non_blocking=True)
        st["t"].copy_(target_local, non_blocking=True)
        st["graph"].replay()
        return st["out"]

    def release(self) -> None:
        """Drop the captured graphs (they hold NCCL kernels): call before ``destroy_process_group``."""
        torch.cuda.synchronize()
        self._state.clear()


def _default_knn(q, r, k):
    from . import ops
    return ops.knn(q, r, k)


def knn_query_sharded(query_local: torch.Tensor, ref_local: torch.Tensor, k: int, group=None,
                      knn_fn: Optional[Callable] = None):
    """kNN with queries sharded and references all-gathered; results stay sharded like the queries.
    Returned indices refer to the gathered (rank-ordered) reference set."""
    knn_fn = knn_fn or _default_knn
    return knn_fn(query_local, all_gather_ragged(ref_local, group), k)


def encode_scans_sharded(encoder, scans: torch.Tensor, group=None) -> torch.Tensor:
    """Batch-of-scans sharding: ``scans`` [S,N,3] is the full batch (same on every rank, e.g. from a
    shared loader); this rank encodes scans ``rank, rank+G, ...`` and the [S,F] feature matrix is
    assembled with one all-gather of the per-rank features (F floats per scan)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    mine = scans_of_rank(scans.shape[0], world, rank)
    feats = encoder(scans[mine].contiguous()) if mine else None
    if world == 1:
        return feats
    per = (scans.shape[0] + world - 1) // world
    F = feats.shape[1] if feats is not None else 0
    Fmax = torch.tensor([F], device=scans.device)
    dist.all_reduce(Fmax, op=dist.ReduceOp.MAX, group=group)
    F = int(Fmax.item())
    buf = scans.new_zeros(per, F)
    if feats is not None:
        buf[: feats.shape[0]] = feats
    bufs = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(bufs, buf, group=group)
    out = scans.new_zeros(scans.shape[0], F)
    for r in range(world):
        idx = scans_of_rank(scans.shape[0], world, r)
        out[idx] = bufs[r][: len(idx)]
    return out

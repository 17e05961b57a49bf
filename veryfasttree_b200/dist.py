"""Multi-rank plumbing of bench.py (torch.distributed: NCCL on GPUs, gloo in the CPU tests).

ONE tree is sharded over the ranks inside the C-ABI library (vft_dist_init, csrc/vft_dist.cuh): every rank runs
the same host loop on a replicated slab, the candidate axis of the all-candidate sweeps is split and their
results are all-gathered over NVLink peer memory (or NCCL).  This module only bootstraps that group from
torch.distributed -- `init_sharded` hands the NCCL unique id from rank 0 to the others; `init_sharded_host`
routes the exchange through a torch.distributed all_gather on host bytes (gloo: the CPU double of the tests) --
and aggregates measurements (max time / sum of units) for bench.py.

`sharded_one_vs_all` is the older Python-level demonstration of the same exchange pattern (kept for its test).
"""
from __future__ import annotations

import numpy as np


def aggregate_step_times(dev_ms: float, e2e_s: float, units: float, launches: float, device=None):
    """(max dev_ms, max e2e_s, sum units, sum launches) over the ranks of the default group."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return dev_ms, e2e_s, units, launches
    t = torch.tensor([dev_ms, e2e_s, units, launches], dtype=torch.float64, device=device)
    mx = t.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    sm = t.clone()
    dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    return mx[0].item(), mx[1].item(), sm[2].item(), sm[3].item()


def candidate_block(n_nodes: int, rank: int, world: int):
    """Contiguous node-id block [begin, end) owned by `rank`."""
    per = (n_nodes + world - 1) // world
    return min(n_nodes, rank * per), min(n_nodes, (rank + 1) * per)


def merge_top_k(js, dist_, weight, crit, k):
    """Merge per-rank records into the global top-k in psort order: criterion asc, node id desc."""
    j = np.concatenate(js); d = np.concatenate(dist_); w = np.concatenate(weight); c = np.concatenate(crit)
    order = np.lexsort((-j, c))[:k]
    return j[order], d[order], w[order], c[order]


def sharded_one_vs_all(ctx, query: int, n_active: int, k: int, device=None):
    """setBestHit + top-k with the candidate axis sharded over the ranks of the default group.
    Every rank holds the profiles (replicated slab); returns the same merged top-k on every rank."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    lo, hi = candidate_block(ctx.maxnode(), rank, world)
    j, d, w, c = ctx.dist_one_vs_all(query, n_active, k, j_begin=lo, j_end=hi)
    if world == 1:
        return j, d, w, c
    # fixed-size records: k x (j, dist, weight, crit) as float64 + a count, one all_gather
    rec = torch.full((k, 4), np.nan, dtype=torch.float64, device=device)
    n = len(j)
    if n:
        rec[:n] = torch.from_numpy(np.stack([j.astype(np.float64), d.astype(np.float64), w.astype(np.float64),
                                              c.astype(np.float64)], axis=1)).to(rec.device)
    out = [torch.empty_like(rec) for _ in range(world)]
    dist.all_gather(out, rec)
    js, ds, ws, cs = [], [], [], []
    for t in out:
        a = t.cpu().numpy()
        a = a[~np.isnan(a[:, 0])]
        js.append(a[:, 0].astype(np.int64)); ds.append(a[:, 1].astype(d.dtype)); ws.append(a[:, 2].astype(d.dtype)); cs.append(a[:, 3].astype(d.dtype))
    return merge_top_k(js, ds, ws, cs, k)


def init_sharded(lib, device: int):
    """Collective: make the process group of torch.distributed a vft dist group on `device` (NCCL unique id from rank 0)."""
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    box = [lib.dist_unique_id() if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(box, src=0)
    lib.dist_init(rank, world, box[0], device)
    return lib.dist_info()


def init_sharded_host(lib, device: int = 0):
    """The same group with the exchange routed through torch.distributed.all_gather on host bytes (any backend)."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()

    def allgather(payload: bytes) -> bytes:
        t = torch.frombuffer(bytearray(payload), dtype=torch.uint8)
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return b"".join(o.numpy().tobytes() for o in out)

    lib.dist_init_host(rank, world, allgather, device)
    return lib.dist_info()

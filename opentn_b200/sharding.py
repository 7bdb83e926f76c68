"""Instance-parallel sharding over the GPUs of one node.

Every optimisation instance (tau, Kraus rank, seed, noise model) is independent: instance b lives on rank b mod world and
never talks to another during the fit.  The only collective of the path is ONE all-gather at the end (final costs, status
flags and isometries), issued through torch.distributed (NCCL over NVLink on GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np


def my_instances(total: int, rank: int, world: int) -> np.ndarray:
    """Global indices owned by `rank` (round robin: b -> rank b mod world)."""
    return np.arange(rank, total, world)


def shard_sizes(total: int, world: int) -> list:
    return [len(range(r, total, world)) for r in range(world)]


def gather_instances(local, total: int, group=None):
    """All-gather per-instance results.  `local` is a torch tensor [n_local, ...] holding this rank's instances in the order of
    `my_instances`; returns [total, ...] in global instance order on every rank.  Shards may be uneven (padded to the
    largest shard for the collective)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        assert local.shape[0] == total
        return local
    world = dist.get_world_size(group)
    sizes = shard_sizes(total, world)
    cap = max(sizes)
    pad = torch.zeros((cap,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    out = torch.empty((total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    for r in range(world):
        out[r::world] = parts[r][: sizes[r]]
    return out


def sharded_fit(run_local, xs_all, group=None):
    """Run `run_local(x_shard) -> (x_final_shard, f_final_shard, status_shard)` on this rank's instances of `xs_all`
    ([total, m, n, p], identical on every rank or only meaningfully filled for the local shard) and gather everything.
    Returns (x_all, f_all, status_all) in global order."""
    import torch
    import torch.distributed as dist
    total = xs_all.shape[0]
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    idx = my_instances(total, rank, world)
    x_loc, f_loc, s_loc = run_local(xs_all[torch.as_tensor(idx, device=xs_all.device)])
    return (gather_instances(x_loc, total, group), gather_instances(f_loc, total, group),
            gather_instances(s_loc, total, group))

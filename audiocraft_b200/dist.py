"""Batch-split data parallelism for the two hot paths.

Items are independent (per-item KV cache, per-item conv / LSTM state; a CFG pair stays on one GPU), so the N>1 path is
one process per GPU, a contiguous split of the batch, replicated weights and NO collective during compute; the only
exchange is an optional all-gather of the outputs (codes [B/N,4,T] int64 or waveforms) at the end.
"""
import typing as tp

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, rank: int, world: int) -> tp.Tuple[int, int]:
    """Contiguous, balanced split: the first (n_items % world) ranks take one extra item."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_list(items: tp.Sequence, rank: tp.Optional[int] = None, world: tp.Optional[int] = None) -> tp.List:
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    a, b = shard_bounds(len(items), rank, world)
    return list(items[a:b])


def gather_batch(local: torch.Tensor, n_items: int) -> torch.Tensor:
    """All-gather per-rank outputs [b_r, ...] (ragged b_r allowed) back into the global batch order [n_items, ...]."""
    world = dist.get_world_size()
    if world == 1:
        return local
    sizes = [shard_bounds(n_items, r, world) for r in range(world)]
    biggest = max(b - a for a, b in sizes)
    pad = torch.zeros((biggest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat([o[: b - a] for o, (a, b) in zip(out, sizes)], dim=0)

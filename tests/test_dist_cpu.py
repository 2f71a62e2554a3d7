"""world_size-2 gloo test of the N>1 path's host logic: contiguous batch split + ragged all-gather of outputs."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_items):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from audiocraft_b200.dist import gather_batch, shard_list
    items = list(range(n_items))
    mine = shard_list(items)
    # stand-in for "generate codes for my items": a deterministic function of the item id
    local = torch.stack([torch.full((4, 6), i, dtype=torch.int64) + torch.arange(6) for i in mine]) if mine else \
        torch.zeros((0, 4, 6), dtype=torch.int64)
    full = gather_batch(local, n_items)
    want = torch.stack([torch.full((4, 6), i, dtype=torch.int64) + torch.arange(6) for i in items])
    assert torch.equal(full, want), (rank, full.shape)
    dist.barrier()
    dist.destroy_process_group()


def test_batch_split_and_gather_world2():
    for n_items in (8, 5):
        mp.spawn(_worker, args=(2, _free_port(), n_items), nprocs=2, join=True)

"""CPU tests (gloo, world_size 2) of the data-parallel plumbing used by the training step."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from transformertts_b200.utils.data_parallel import GradSync, global_loss, init_from_env, shard_rows
    r, w = init_from_env(backend='gloo')
    assert (r, w) == (rank, world)
    flat = torch.arange(1000, dtype=torch.float32) * (rank + 1)
    gs = GradSync(flat)
    gs.bucket_ready(600, 900)       # "decoder" bucket first, while the rest is still being written
    flat[:600] += 1.0               # late writes to the not-yet-sent part are still picked up
    scale = gs.finish()
    want = torch.arange(1000, dtype=torch.float32) * 3.0
    want[:600] += 2.0
    ok = torch.equal(flat, want) and scale == 0.5
    loss = global_loss(torch.tensor(1.0 + rank), 10 * (rank + 1))  # weighted: (1*10 + 2*20)/30
    ok = ok and abs(float(loss) - 50.0 / 30.0) < 1e-6
    rows = shard_rows(7, rank, world)
    ok = ok and (rows == (slice(0, 4) if rank == 0 else slice(4, 7)))
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_grad_sync_buckets_and_weighted_loss_world2():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}


def test_single_process_is_identity():
    from transformertts_b200.utils.data_parallel import GradSync, shard_rows
    flat = torch.ones(10)
    gs = GradSync(flat)
    gs.bucket_ready(0, 5)
    assert gs.finish() == 1.0 and torch.equal(flat, torch.ones(10))
    assert shard_rows(64, 3, 8) == slice(24, 32)

"""world_size-2 gloo test of the multi-rank host logic (shard bounds + gather) on CPU tensors."""
import os
import socket

import numpy as NP
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, nbl, q, interleave=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from prisim_b200.sharding import gather_baseline_shards, shard_rows
    full = torch.arange(nbl * 5 * 2, dtype=torch.float64).reshape(nbl, 5, 2)
    full = torch.complex(full[..., 0], full[..., 1])
    sl = shard_rows(nbl, world, rank, interleave)
    out = gather_baseline_shards(full[sl].clone(), nbl, dst=0, interleave=interleave)
    if rank == 0:
        q.put(bool(torch.equal(out, full)))
    else:
        q.put(out is None)
    dist.barrier()
    dist.destroy_process_group()


def _run(world, nbl, interleave=False):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nbl, q, interleave)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    return res


def test_gather_two_ranks_uneven():
    assert all(_run(2, 7))


def test_gather_three_ranks_with_empty_shard():
    assert all(_run(3, 2))


def test_gather_interleaved_shards_two_and_three_ranks():
    assert all(_run(2, 7, interleave=True))
    assert all(_run(3, 2, interleave=True))          # rank 2 owns no baseline


def test_shard_rows_cover_every_baseline_once():
    from prisim_b200.sharding import shard_rows
    for nbl in (1, 7, 61075):
        for world in (1, 2, 8):
            for interleave in (False, True):
                rows = NP.concatenate([NP.arange(nbl)[shard_rows(nbl, world, r, interleave)] for r in range(world)])
                assert NP.array_equal(NP.sort(rows), NP.arange(nbl))

"""Multi-GPU sharding of the hot path (replaces run_prisim.py's mpi4py chunking,
scripts/run_prisim.py:1729-1801 chunk sizing, :2165-2209 per-chunk observe, :2233-2276 rank-0
concatenation through ``_part_N.hdf5`` files).

V[b,f,t] for different baselines are independent sums, so ranks own contiguous baseline blocks
(the reference's ``pp.key='bl'``, ``eqvol=true`` mode) and compute them with no data-path
collective; one gather of the finished shards to the writing rank replaces the file exchange.
Sources are never sharded (that would need a sum-reduce, run_prisim.py:1846-1856).

Works with any ``torch.distributed`` backend: NCCL with CUDA tensors on the GPU box (grouped
ncclSend/ncclRecv over NVLink), gloo with CPU tensors in the host-logic tests.
"""
from __future__ import annotations

import numpy as NP
import torch
import torch.distributed as dist


def shard_bounds(nbl, world_size):
    """Contiguous, near-equal baseline blocks: bounds[r]:bounds[r+1] belongs to rank r."""
    base, rem = divmod(int(nbl), int(world_size))
    sizes = [base + (1 if r < rem else 0) for r in range(world_size)]
    return NP.concatenate(([0], NP.cumsum(sizes))).astype(int)


def shard_slice(nbl, world_size, rank):
    b = shard_bounds(nbl, world_size)
    return slice(int(b[rank]), int(b[rank + 1]))


def gather_baseline_shards(local, nbl_total, dst=0, group=None):
    """Gather [nbl_local, ...] shards (baseline axis first) into [nbl_total, ...] on rank `dst`.
    One batched point-to-point exchange; other ranks return None."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    bounds = shard_bounds(nbl_total, world)
    local = local.contiguous()
    if rank == dst:
        full = torch.empty((nbl_total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        full[bounds[dst]:bounds[dst + 1]] = local
        ops = []
        views = []
        for r in range(world):
            if r == dst or bounds[r + 1] == bounds[r]:
                continue
            v = full[bounds[r]:bounds[r + 1]]
            buf = v if v.is_contiguous() else torch.empty_like(v)
            views.append((v, buf))
            ops.append(dist.P2POp(dist.irecv, buf, r, group=group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        for v, buf in views:
            if buf.data_ptr() != v.data_ptr():
                v.copy_(buf)
        return full
    if local.shape[0] > 0:
        for req in dist.batch_isend_irecv([dist.P2POp(dist.isend, local, dst, group=group)]):
            req.wait()
    return None


def make_sharded_array(cls, labels, baselines, channels, rank=None, world_size=None, **kwargs):
    """Construct the rank-local ``InterferometerArray`` over this rank's baseline block; noise is
    keyed by the global baseline index (bl_offset/nbl_total) so any world size gives the same run."""
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_initialized() else 1
    baselines = NP.asarray(baselines)
    sl = shard_slice(baselines.shape[0], world_size, rank)
    labels = NP.asarray(labels)[sl] if not isinstance(labels, list) else labels[sl]
    return cls(labels, baselines[sl], channels, bl_offset=sl.start, nbl_total=baselines.shape[0], **kwargs)

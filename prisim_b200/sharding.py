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


def shard_rows(nbl, world_size, rank, interleave=False):
    """Rows of the full baseline list owned by `rank`: a contiguous block (the reference's pp.key='bl' chunks,
    scripts/run_prisim.py:1775-1791) or, interleaved, every world_size-th baseline starting at `rank`.  Interleaving
    deals every kind of baseline (PRISim sorts them by length; the short, strongly cancelling ones that the precision
    control recomputes in fp64 all sit at the front) evenly to the ranks."""
    if interleave:
        return slice(int(rank), int(nbl), int(world_size))
    return shard_slice(nbl, world_size, rank)


def gather_baseline_shards(local, nbl_total, dst=0, group=None, interleave=False):
    """Gather [nbl_local, ...] shards (baseline axis first) into [nbl_total, ...] on rank `dst`.
    One batched point-to-point exchange; other ranks return None."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    rows = [shard_rows(nbl_total, world, r, interleave) for r in range(world)]
    nrows = [len(range(*sl.indices(nbl_total))) for sl in rows]
    local = local.contiguous()
    if rank == dst:
        full = torch.empty((nbl_total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        full[rows[dst]] = local
        ops = []
        views = []
        for r in range(world):
            if r == dst or nrows[r] == 0:
                continue
            v = full[rows[r]]
            buf = v if v.is_contiguous() else torch.empty_like(v, memory_format=torch.contiguous_format)
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


def make_sharded_array(cls, labels, baselines, channels, rank=None, world_size=None, interleave=False, **kwargs):
    """Construct the rank-local ``InterferometerArray`` over this rank's baselines (``shard_rows``); noise is
    keyed by the global baseline index (bl_offset / bl_step / nbl_total) so any world size gives the same run."""
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_initialized() else 1
    baselines = NP.asarray(baselines)
    if kwargs.get("noise_seed", None) is None and dist.is_initialized() and world_size > 1:
        # one noise seed for the whole run: rank 0 draws it, every shard uses it with its global baseline offset
        from .interferometry import fresh_noise_seed
        box = [fresh_noise_seed() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        kwargs["noise_seed"] = box[0]
    sl = shard_rows(baselines.shape[0], world_size, rank, interleave)
    labels = NP.asarray(labels)[sl] if not isinstance(labels, list) else labels[sl]
    return cls(labels, baselines[sl], channels, bl_offset=sl.start, bl_step=sl.step or 1, nbl_total=baselines.shape[0], **kwargs)


class _RawCudaBuffer(object):
    """Minimal ``__cuda_array_interface__`` carrier so torch can view memory this package allocated or mapped."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class PeerGatherBuffer(object):
    """Gather through peer memory: rank `dst` owns the output buffer; every other rank maps it over NVLink (CUDA IPC)
    and lets its kernels write the result slice directly (pass ``local`` as ``out=`` to ``engine.skyvis``).
    ``wait()`` = stream sync + barrier; afterwards ``full`` on `dst` holds all slices.

      * ``row_bounds=None`` (snapshot sharding): the buffer is [world, *shape], rank r owns ``full[r]``;
      * ``row_bounds=shard_bounds(nbl, world)`` (baseline sharding, the reference's pp.key='bl' mode): the buffer
        is ``shape`` = [nbl, ...] itself and rank r owns rows row_bounds[r]:row_bounds[r+1] -- rank `dst` ends up
        with exactly the array an unsharded run would have produced, with no concatenation step
        (scripts/run_prisim.py:2233-2242);
      * ``interleave=True``: the same [nbl, ...] buffer, rank r owns rows r, r + world, ... (``local`` is a strided
        view; ``pb200_skyvis`` takes the row stride).

    Falls back to an NCCL point-to-point gather when peer mapping is unavailable (``mode == 'nccl'``;
    ``force_nccl=True`` selects it for A/B tests)."""

    def __init__(self, shape, device, dst=0, group=None, row_bounds=None, force_nccl=False, interleave=False):
        import ctypes as C
        from . import _lib
        self.group, self.dst, self.shape = group, dst, tuple(shape)
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.device = int(device)
        self.ctx = _lib.get_context(self.device)
        self.bounds = None if row_bounds is None else [int(b) for b in row_bounds]
        self.interleave = bool(interleave)
        dev = "cuda:{0}".format(self.device)
        nel = int(NP.prod(self.shape))
        row_el = nel // self.shape[0] if self.shape[0] else 0
        whole = self.bounds is not None or self.interleave          # the buffer is the [nbl, ...] array itself
        full_shape = self.shape if whole else (self.world,) + self.shape
        full_el = int(NP.prod(full_shape))
        self._base = C.c_void_p()
        self._mapped = False
        self._owned = False
        handle = torch.zeros(65, dtype=torch.uint8, device=dev)               # 64-byte handle + ok flag
        if self.rank == dst and not force_nccl:
            ok = self.ctx.lib.pb200_device_alloc(self.ctx.handle, max(full_el, 1) * 16, C.byref(self._base)) == 0
            self._owned = ok
            hb = (C.c_ubyte * 64)()
            ok = ok and self.ctx.lib.pb200_peer_export(self.ctx.handle, self._base, hb) == 0
            handle[:64] = torch.tensor(list(hb), dtype=torch.uint8)
            handle[64] = 1 if ok else 0
        dist.broadcast(handle, src=dst, group=group)
        ok = bool(handle[64].item())
        if ok and self.rank != dst:
            hb = (C.c_ubyte * 64)(*handle[:64].cpu().tolist())
            ok = self.ctx.lib.pb200_peer_open(self.ctx.handle, hb, C.byref(self._base)) == 0
            self._mapped = ok
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        self.mode = "peer" if bool(flag.item()) else "nccl"
        if self.interleave:
            nloc = len(range(self.rank, self.shape[0], self.world))
            lo, n_local, local_shape = 0, nloc * row_el, (nloc,) + self.shape[1:]
        elif self.bounds is None:
            lo, n_local = self.rank * nel, nel
            local_shape = self.shape
        else:
            lo, n_local = self.bounds[self.rank] * row_el, (self.bounds[self.rank + 1] - self.bounds[self.rank]) * row_el
            local_shape = (self.bounds[self.rank + 1] - self.bounds[self.rank],) + self.shape[1:]
        if self.mode == "peer":
            base = self._base.value
            if self.interleave and n_local > 0:
                self.local = torch.as_tensor(_RawCudaBuffer(base, full_shape, "<c16"), device=dev)[self.rank::self.world]
            elif n_local > 0:
                self.local = torch.as_tensor(_RawCudaBuffer(base + lo * 16, local_shape, "<c16"), device=dev)
            else:
                self.local = torch.empty(local_shape, dtype=torch.complex128, device=dev)
            self.full = torch.as_tensor(_RawCudaBuffer(base, full_shape, "<c16"), device=dev) if self.rank == dst else None
        else:
            if self._owned:                                  # exported but somebody could not map it
                self.ctx.lib.pb200_device_free(self.ctx.handle, self._base)
                self._owned = False
            elif self._mapped:
                self.ctx.lib.pb200_peer_close(self.ctx.handle, self._base)
                self._mapped = False
            self._base.value = None
            self.full = torch.empty(full_shape, dtype=torch.complex128, device=dev) if self.rank == dst else None
            if self.rank == dst:
                self.local = self._slice_of(dst)
            else:
                self.local = torch.empty(local_shape, dtype=torch.complex128, device=dev)

    def _slice_of(self, r):
        if self.interleave:
            return self.full[r::self.world]
        return self.full[r] if self.bounds is None else self.full[self.bounds[r]:self.bounds[r + 1]]

    def wait(self):
        """Make every rank's slice visible on `dst` (peer mode: stores are complete when the writers' streams are;
        nccl mode: one batched point-to-point gather)."""
        if self.mode == "nccl":
            staged = []
            if self.rank == self.dst:
                ops = []
                for r in range(self.world):
                    v = self._slice_of(r)
                    if r == self.dst or v.numel() == 0:
                        continue
                    buf = v if v.is_contiguous() else torch.empty(v.shape, dtype=v.dtype, device=v.device)
                    staged.append((v, buf))
                    ops.append(dist.P2POp(dist.irecv, buf, r, group=self.group))
            else:
                ops = [dist.P2POp(dist.isend, self.local, self.dst, group=self.group)] if self.local.numel() > 0 else []
            for req in (dist.batch_isend_irecv(ops) if ops else []):
                req.wait()
            for v, buf in staged:
                if buf.data_ptr() != v.data_ptr():
                    v.copy_(buf)
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)

    def close(self):
        if self.mode == "peer" and self._base.value:
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.group)
            self.local = self.full = None
            if self.rank == self.dst:
                self.ctx.lib.pb200_device_free(self.ctx.handle, self._base)
            elif self._mapped:
                self.ctx.lib.pb200_peer_close(self.ctx.handle, self._base)
            self._base.value = None


class ShardedObserver(object):
    """One snapshot at a time over all ranks, sharded over baselines (the reference's ``pp.key='bl'`` equal-volume
    mode, scripts/run_prisim.py:1775-1791 / :2165-2209) with the concatenation on the writing rank
    (:2233-2242) fused into the kernel epilogue: every rank owns an ``InterferometerArray`` over its share of the
    baselines (interleaved by default, ``shard_rows``; ``interleave=False`` gives the reference's contiguous chunks)
    whose phase-sum kernel stores straight into its rows of rank `dst`'s [nbl_total, nchan] buffer.

        so = ShardedObserver(InterferometerArray, labels, baselines, channels, device=local_rank, ...)
        full = so.observe(timeobj, Tsysinfo, bandpass, pointing, skymodel, t_acc)    # rank dst: [nbl_total, nchan] CUDA tensor

    ``observe`` returns the gathered snapshot on `dst` (a view of one of `nbuf` gather buffers used in turn, so it stays
    valid for the next nbuf - 1 calls: with the default 2 the device->host copy of snapshot j can overlap the kernels of
    snapshot j + 1) and None elsewhere; ``so.ia`` is the rank-local array (noise / delay transforms stay local: channels
    are unsplit)."""

    def __init__(self, cls, labels, baselines, channels, dst=0, group=None, force_nccl=False, interleave=True, nbuf=2, **kwargs):
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.dst = dst
        baselines = NP.asarray(baselines)
        self.nbl_total = baselines.shape[0]
        self.interleave = bool(interleave)
        self.bounds = shard_bounds(self.nbl_total, self.world)
        self.rows = shard_rows(self.nbl_total, self.world, self.rank, self.interleave)
        self.ia = make_sharded_array(cls, labels, baselines, channels, rank=self.rank, world_size=self.world, interleave=self.interleave,
                                     **kwargs)
        # the sampled fp64 audit of the precision control is a property of the snapshot, not of a shard: spread its baselines over the ranks
        if self.world > 1 and hasattr(self.ia, "audit_baselines"):
            self.ia.audit_baselines = max(4, -(-self.ia.audit_baselines // self.world))
        self.gbufs, self._turn = [], 0
        if self.world > 1:
            for _ in range(max(1, int(nbuf))):
                self.gbufs.append(PeerGatherBuffer((self.nbl_total, self.ia.channels.size), self.ia.device, dst=dst, group=group,
                                                   row_bounds=None if self.interleave else self.bounds, interleave=self.interleave,
                                                   force_nccl=force_nccl))

    @property
    def gbuf(self):
        """the gather buffer the next observe() will fill"""
        return self.gbufs[self._turn % len(self.gbufs)] if self.gbufs else None

    def observe(self, *args, **kwargs):
        if not self.gbufs:
            self.ia.observe(*args, **kwargs)
            return self.ia.skyvis_freq_device(-1)
        g = self.gbuf
        self._turn += 1
        self.ia.next_skyvis_out = g.local                    # the kernel epilogue writes into the writing rank's buffer
        self.ia.observe(*args, **kwargs)
        g.wait()
        return g.full if self.rank == self.dst else None

    def close(self):
        for g in self.gbufs:
            g.close()
        self.gbufs = []

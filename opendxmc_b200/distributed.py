"""One process per GPU: histories are sharded by the C ABI (dxb_set_history_range), every rank holds a
full replica of the voxel grid and tables, and the per-beam 64-bit fixed-point tallies are summed with ONE
torch.distributed reduce (NCCL over NVLink 5 / NVSwitch on the GPU box; gloo in the CPU tests).  Integer sums are
order independent, so the dose is bitwise identical for any GPU count.

torch is plumbing here (process group + a tensor view of the tally buffer), never the compute path.
"""
import ctypes as C

from . import _capi as K


class _DevView:
    """exposes a raw device pointer through __cuda_array_interface__ (int64 words)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 3, "strides": None}


def shard_local_count(n_total, rank, world):
    return int(K.load().dxb_shard_local_count(int(n_total), int(rank), int(world)))


def shard_history_id(local_index, rank, world):
    return int(K.load().dxb_shard_history_id(int(local_index), int(rank), int(world)))


def tally_tensor(world, device_index):
    """torch int64 view (no copy) of the raw tally buffer of `world`'s first device."""
    import torch
    ptr, n = C.c_void_p(), C.c_uint64()
    rc = K.load().dxb_tally_buffer(world.ctx(), C.byref(ptr), C.byref(n))
    if rc != K.DXB_OK:
        raise K.DxbError(rc, "dxb_tally_buffer")
    return torch.as_tensor(_DevView(ptr.value, n.value), device=f"cuda:{device_index}")


def reduce_tallies(tally, dst=0, group=None):
    """the single exchange step of the path: integer sum of the per-rank tallies onto rank `dst`."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.reduce(tally, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return tally


def run_beam(world, beam, tally, rank, use_beam_calibration=True, progress=None, group=None):
    """Transport::operator() across ranks: shard tallies -> reduce -> (rank 0) calibration + energy->dose.
    Returns the calibration factor on rank 0, None elsewhere."""
    import torch
    from .api import Transport
    tr = Transport()
    tr.run_transport(world, beam, progress)  # returns after the transport kernels have finished (library stream)
    reduce_tallies(tally, 0, group)
    # the reduce is enqueued on torch's current stream, energy->dose runs on the library's own stream: nothing else
    # orders the two, so wait for the reduce before the tallies are converted
    if tally.is_cuda:
        torch.cuda.current_stream(tally.device).synchronize()
    if rank == 0:
        return tr.finish_beam(world, beam, use_beam_calibration)
    return None


def slab(n_voxels, rank, size):
    """voxel slab [begin, end) of `rank`: the partition used by the fused exchange and the sharded upload / read-out."""
    return n_voxels * rank // size, n_voxels * (rank + 1) // size


def set_grid_sharded(world, dim, spacing, density, material, device_index, group=None, stream=None):
    """AAVoxelGrid::setData + World::build for one process per GPU (include/dxb.h: dxb_set_grid_sharded): every rank
    uploads and packs only ITS slab of the caller's arrays, then the packed 4-byte slabs are broadcast and the
    per-material density maxima max-reduced over NVLink, and every rank finishes with the majorant.  The result is the
    grid dxb_set_grid builds, for 1/N of the host-to-device bytes per rank.  torch is plumbing (NCCL calls on views of
    the library's buffers); `density` / `material` are the full host arrays (pinned for an asynchronous copy)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    lib = K.load()
    group = group if group is not None else dist.group.WORLD
    rank, size = dist.get_rank(group), dist.get_world_size(group)
    n = int(dim[0]) * int(dim[1]) * int(dim[2])
    b, e = slab(n, rank, size)
    ctx = world.ctx()
    cdim = (C.c_uint64 * 3)(*[int(v) for v in dim])
    csp = (C.c_double * 3)(*[float(v) for v in spacing])
    density = np.ascontiguousarray(density, dtype=np.float64)
    material = np.ascontiguousarray(material, dtype=np.uint8)
    # a context with a library-managed exchange: no peer may still be pulling from the tally buffers this call clears
    lib.dxb_flush(ctx)
    if size > 1:
        dist.barrier(group=group)
    rc = lib.dxb_set_grid_sharded(ctx, cdim, csp, density.ctypes.data_as(K.c_double_p), material.ctypes.data_as(K.c_u8_p), b, e)
    if rc != K.DXB_OK:
        raise K.DxbError(rc, "dxb_set_grid_sharded", (lib.dxb_last_error(ctx) or b"").decode())
    vox, mx, nv = C.c_void_p(), C.c_void_p(), C.c_uint64()
    rc = lib.dxb_grid_buffers(ctx, C.byref(vox), C.byref(mx), C.byref(nv))
    if rc != K.DXB_OK:
        raise K.DxbError(rc, "dxb_grid_buffers")

    def view(ptr, count):
        iface = {"shape": (count,), "typestr": "<i4", "data": (ptr, False), "version": 3, "strides": None}
        return torch.as_tensor(type("_V", (), {"__cuda_array_interface__": iface})(), device=f"cuda:{device_index}")
    voxels, maxbits = view(vox.value, n), view(mx.value, 257)
    torch.cuda.synchronize(device_index)  # the slab is packed (the library's stream need not be torch's)
    ctxmgr = torch.cuda.stream(stream) if stream is not None else _nullcontext()
    with ctxmgr:
        if size > 1:
            if n % size == 0:
                dist.all_gather_into_tensor(voxels, voxels[b:e].clone(), group=group)
            else:
                for r in range(size):
                    rb, re = slab(n, r, size)
                    if re > rb:
                        dist.broadcast(voxels[rb:re], src=dist.get_global_rank(group, r), group=group)
            # non-negative floats and material indices order like their bit patterns read as int32
            dist.all_reduce(maxbits, op=dist.ReduceOp.MAX, group=group)
        (stream if stream is not None else torch.cuda.current_stream(device_index)).synchronize()
    rc = lib.dxb_finish_grid(ctx)
    if rc != K.DXB_OK:
        raise K.DxbError(rc, "dxb_finish_grid", (lib.dxb_last_error(ctx) or b"").decode())


class _nullcontext:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


class FusedExchange:
    """The exchange step fused with energy->dose over NVLink / NVSwitch (include/dxb.h: dxb_finish_beam_sharded).

    The tally buffer of every rank is re-homed into torch symmetric memory, so that each rank sees its peers' buffers
    (P2P mappings) and, on NVSwitch systems, one multicast address whose loads return the in-switch SUM of all ranks'
    words (multimem.ld_reduce).  After the transport kernels, rank r converts ITS slab of voxels reading that sum
    directly: no reduced copy of the 32 B/voxel tallies is ever written, and the 1/N slabs run in parallel.  The dose
    score of a slab stays on its rank until `gather_dose` refreshes rank 0's copy (read-out time, not per beam).
    torch supplies the memory mapping and the barrier; the arithmetic is libdxmc_b200's kernel.
    """

    def __init__(self, world, local_rank, group=None, multicast=True):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self._torch, self._dist = torch, dist
        self.world = world
        self.group = group if group is not None else dist.group.WORLD
        self.rank = dist.get_rank(self.group)
        self.size = dist.get_world_size(self.group)
        self.device = torch.device("cuda", local_rank)
        lib = K.load()
        ptr, n = C.c_void_p(), C.c_uint64()
        rc = lib.dxb_tally_buffer(world.ctx(), C.byref(ptr), C.byref(n))
        if rc != K.DXB_OK:
            raise K.DxbError(rc, "dxb_tally_buffer")
        self.n_words = int(n.value)
        self.n_voxels = self.n_words // 4
        self.tally = symm.empty(self.n_words, dtype=torch.int64, device=self.device)
        self.handle = symm.rendezvous(self.tally, self.group)
        mc = int(self.handle.multicast_ptr or 0) if multicast else 0   # 0: no NVSwitch multicast object on this system
        self.multicast_ptr = mc
        self.peer_ptrs = [int(p) for r, p in enumerate(self.handle.buffer_ptrs) if r != self.rank]
        self.begin = self.n_voxels * self.rank // self.size
        self.end = self.n_voxels * (self.rank + 1) // self.size
        self.kind = "nvls-multicast" if mc else "p2p-pull"
        # adopt last: nothing above may leave the context pointing at memory that is about to be freed
        rc = lib.dxb_set_tally_storage(world.ctx(), C.c_void_p(self.tally.data_ptr()), self.n_words)
        if rc != K.DXB_OK:
            raise K.DxbError(rc, "dxb_set_tally_storage", (lib.dxb_last_error(world.ctx()) or b"").decode())

    def slab(self, rank):
        return self.n_voxels * rank // self.size, self.n_voxels * (rank + 1) // self.size

    def barrier(self):
        """all ranks' work enqueued so far is complete (device-side barrier over the symmetric signal pads)."""
        self.handle.barrier(channel=0)
        self._torch.cuda.current_stream(self.device).synchronize()

    def finish_beam(self, beam, physics_mode, use_beam_calibration=True):
        """call after dxb_run_transport on every rank; returns the calibration factor (identical on all ranks)."""
        lib = K.load()
        self.barrier()  # every rank's transport kernels have finished: the tallies are final
        f = C.c_double()
        peers = (K.VP * max(1, len(self.peer_ptrs)))(*self.peer_ptrs)
        rc = lib.dxb_finish_beam_sharded(self.world.ctx(), C.byref(beam.desc()), int(physics_mode), 1 if use_beam_calibration else 0,
                                         C.c_void_p(self.multicast_ptr) if self.multicast_ptr else None, peers, len(self.peer_ptrs),
                                         self.begin, self.end, C.byref(f))
        if rc != K.DXB_OK:
            raise K.DxbError(rc, "dxb_finish_beam_sharded", (lib.dxb_last_error(self.world.ctx()) or b"").decode())
        self.barrier()  # nobody clears its tallies for the next beam while a peer still reads them
        return f.value

    def _dose_tensors(self):
        torch = self._torch
        lib = K.load()
        d, v, e, n = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_uint64()
        rc = lib.dxb_dose_buffers(self.world.ctx(), C.byref(d), C.byref(v), C.byref(e), C.byref(n))
        if rc != K.DXB_OK:
            raise K.DxbError(rc, "dxb_dose_buffers")

        def view(ptr, typestr):
            iface = {"shape": (int(n.value),), "typestr": typestr, "data": (ptr.value, False), "version": 3, "strides": None}
            holder = type("_V", (), {"__cuda_array_interface__": iface})()
            return torch.as_tensor(holder, device=self.device)
        return view(d, "<f8"), view(v, "<f8"), view(e, "<i8")

    def gather_dose(self, dst=0):
        """refresh rank `dst`'s copy of the remote slabs of the accumulated dose score (dose, variance, events)."""
        dist = self._dist
        ops = []
        for t in self._dose_tensors():
            if self.rank == dst:
                for r in range(self.size):
                    if r != dst:
                        b, e = self.slab(r)
                        if e > b:
                            ops.append(dist.P2POp(dist.irecv, t[b:e], r, self.group))
            elif self.end > self.begin:
                ops.append(dist.P2POp(dist.isend, t[self.begin:self.end], dst, self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        self._torch.cuda.current_stream(self.device).synchronize()

    def close(self):
        """hand the tally storage back to the library before the symmetric allocation goes away."""
        K.load().dxb_set_tally_storage(self.world.ctx(), None, 0)
        self.tally = None
        self.handle = None


def run_beam_fused(world, beam, exchange, use_beam_calibration=True, progress=None):
    """Transport::operator() across ranks with the fused exchange: shard tallies -> (barrier) -> every rank converts
    its slab of the multicast-summed tallies.  Returns the calibration factor."""
    from .api import Transport
    Transport().run_transport(world, beam, progress)
    return exchange.finish_beam(beam, world._item.lowEnergyCorrection, use_beam_calibration)


class PipelinedExchange:
    """The library-managed exchange for one process per GPU (include/dxb.h: dxb_exchange_export / dxb_exchange_import).

    The two tally buffers of every rank are mapped into every other rank through CUDA IPC.  Per beam the caller runs
    dxb_run_transport, a barrier over the ranks (the one synchronisation the library cannot do across processes) and
    dxb_finish_beam, which only ENQUEUES the exchange: copy-engine pulls of this rank's voxel slab from every peer,
    slab reduce -> dose, clear of the previous buffer.  They run underneath the next beam's transport kernels; the dose
    score of slab r stays on rank r (dxb_get_dose_range) and read-out calls wait for pending exchanges.
    torch.distributed moves 128 bytes of handles per rank and provides the barrier; nothing else."""

    kind = "ipc-copy-engine-pipelined"

    def __init__(self, world, local_rank, group=None):
        import torch
        import torch.distributed as dist
        self._torch, self._dist = torch, dist
        self.world = world
        self.group = group if group is not None else dist.group.WORLD
        self.rank = dist.get_rank(self.group)
        self.size = dist.get_world_size(self.group)
        self.device = torch.device("cuda", local_rank)
        lib = K.load()
        ctx = world.ctx()
        mine = C.create_string_buffer(K.DXB_EXCHANGE_HANDLE_BYTES)
        rc = lib.dxb_exchange_export(ctx, mine)
        if rc != K.DXB_OK:
            raise K.DxbError(rc, "dxb_exchange_export", (lib.dxb_last_error(ctx) or b"").decode())
        gathered = [None] * self.size
        dist.all_gather_object(gathered, bytes(mine.raw), group=self.group)
        blob = b"".join(gathered)
        rc = lib.dxb_exchange_import(ctx, self.rank, self.size, blob)
        if rc != K.DXB_OK:
            raise K.DxbError(rc, "dxb_exchange_import", (lib.dxb_last_error(ctx) or b"").decode())
        n = world._item.size()
        self.n_voxels = n
        self.begin, self.end = slab(n, self.rank, self.size)

    def barrier(self):
        self._dist.barrier(group=self.group)

    def finish_beam(self, beam, physics_mode, use_beam_calibration=True):
        """after dxb_run_transport on every rank: barrier, then enqueue the exchange; returns the calibration factor."""
        lib = K.load()
        self.barrier()
        f = C.c_double()
        rc = lib.dxb_finish_beam(self.world.ctx(), C.byref(beam.desc()), int(physics_mode), 1 if use_beam_calibration else 0, C.byref(f))
        if rc != K.DXB_OK:
            raise K.DxbError(rc, "dxb_finish_beam", (lib.dxb_last_error(self.world.ctx()) or b"").decode())
        return f.value

    def flush(self):
        rc = K.load().dxb_flush(self.world.ctx())
        if rc != K.DXB_OK:
            raise K.DxbError(rc, "dxb_flush")

    def times_ms(self):
        """device time of the last flushed exchange on this rank: (pulls, slab reduce -> dose, clear)."""
        t = (C.c_double * 3)()
        K.load().dxb_exchange_times(self.world.ctx(), t)
        return tuple(t)

    def close(self):
        """every rank's pending pulls must have finished before any rank frees the buffers they read"""
        self.flush()
        self.barrier()
        K.load().dxb_exchange_close(self.world.ctx())
        self.barrier()


def run_beam_pipelined(world, beam, exchange, use_beam_calibration=True, progress=None):
    """Transport::operator() across ranks with the library-managed exchange.  Returns the calibration factor."""
    from .api import Transport
    Transport().run_transport(world, beam, progress)
    return exchange.finish_beam(beam, world._item.lowEnergyCorrection, use_beam_calibration)

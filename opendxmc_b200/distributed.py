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
    from .api import Transport
    tr = Transport()
    tr.run_transport(world, beam, progress)
    reduce_tallies(tally, 0, group)
    if rank == 0:
        return tr.finish_beam(world, beam, use_beam_calibration)
    return None

"""Manual multi-rank timing of the pieces of one bench step (torchrun, one rank per GPU): where does the time outside
the transport kernel go?  Usage: torchrun ... profiles/time_exchange.py [multicast|p2p] [histories]"""
import ctypes as C
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import opendxmc_b200 as dx  # noqa: E402
from opendxmc_b200 import _capi as K, distributed as D  # noqa: E402


def main():
    rank, world_size, local_rank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    kind = sys.argv[1] if len(sys.argv) > 1 else "multicast"
    nh = int(float(sys.argv[2])) if len(sys.argv) > 2 else 1_000_000_000
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = K.load()
    wl = dx.workloads.ct_spiral_patient(scale=1, histories=nh)
    world = wl.build_world(1, [local_rank])
    world.set_history_range(rank, world_size)
    ctx = world.ctx()
    ex = D.FusedExchange(world, local_rank, multicast=(kind == "multicast"))
    desc = wl.beam.desc()

    def sync():
        torch.cuda.synchronize()

    rows = []
    for it in range(6):
        lib.dxb_set_seed(ctx, 1000 + it)
        dist.barrier()
        sync()
        t = [time.perf_counter()]
        rc = lib.dxb_run_transport(ctx, C.byref(desc), 1, None)
        assert rc == 0
        t.append(time.perf_counter())
        ex.barrier()
        t.append(time.perf_counter())
        f = C.c_double()
        peers = (K.VP * max(1, len(ex.peer_ptrs)))(*ex.peer_ptrs)
        rc = lib.dxb_finish_beam_sharded(ctx, C.byref(desc), 1, 0, C.c_void_p(ex.multicast_ptr) if ex.multicast_ptr else None, peers,
                                         len(ex.peer_ptrs), ex.begin, ex.end, C.byref(f))
        assert rc == 0
        t.append(time.perf_counter())
        ex.barrier()
        t.append(time.perf_counter())
        st = world.run_stats()
        rows.append([1e3 * (b - a) for a, b in zip(t[:-1], t[1:])] + [st["transport_ms"]])
    allrows = [None] * world_size
    dist.all_gather_object(allrows, rows[-1])
    if rank == 0:
        print("last iteration, per rank: " + " | ".join("r%d run %.2f b1 %.2f fin %.2f b2 %.2f kern %.2f" % (i, *r) for i, r in enumerate(allrows)))
        print(f"{kind} ranks={world_size}: ms  run_transport(host)  barrier1  finish_sharded  barrier2 | transport kernel(events)")
        for r in rows:
            print("   " + "  ".join(f"{v:9.3f}" for v in r), flush=True)
    ex.close()
    world.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

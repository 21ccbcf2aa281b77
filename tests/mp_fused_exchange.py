"""Multi-process GPU check (launched by tests/test_gpu_beams_and_outputs.py with torchrun, one rank per GPU):
the fused exchange (symmetric-memory tallies, multimem.ld_reduce / P2P pull + energy->dose in ONE kernel per slab)
must reproduce the single-GPU dose score bit for bit, for both read paths, over two accumulated beams."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import opendxmc_b200 as dx  # noqa: E402
from opendxmc_b200 import distributed as D  # noqa: E402


def main():
    rank, world_size, local_rank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    wl = dx.workloads.ct_spiral_patient(scale=4, histories=2_000_000, step_deg=5.0)
    kinds = []
    for multicast in (True, False):
        world = wl.build_world(1, [local_rank])
        # rebuild the grid through the sharded upload (slab per rank + NVLink exchange): must give the same voxels
        D.set_grid_sharded(world, wl.dim, wl.spacing, wl.density, wl.material, local_rank)
        world.set_history_range(rank, world_size)
        world.set_calibration_histories(720_000)
        ex = D.FusedExchange(world, local_rank, multicast=multicast)
        kinds.append(ex.kind)
        f1 = D.run_beam_fused(world, wl.beam, ex, use_beam_calibration=True)
        world.set_seed(1234)
        f2 = D.run_beam_fused(world, wl.beam, ex, use_beam_calibration=False)   # a second beam accumulates
        # sharded read-out: every rank reads its own slab; together they must tile the gathered dose score
        mine = world.fetch_dose_range(ex.begin, ex.end)
        ex.gather_dose(0)
        got = world.fetch_dose() if rank == 0 else None
        parts = [None] * world_size if rank == 0 else None
        dist.gather_object([a[ex.begin:ex.end] for a in mine], parts, dst=0)
        if rank == 0:
            for k in range(3):
                tiled = np.concatenate([p[k] for p in parts])
                assert np.array_equal(tiled, got[k]), "sharded read-out differs from the gathered dose score"
        ex.close()
        world.close()
        if rank == 0:
            ref_world = wl.build_world(1, [local_rank])
            ref_world.set_calibration_histories(720_000)
            tr = dx.Transport()
            assert tr(ref_world, wl.beam, None, True)
            ref_world.set_seed(1234)
            assert tr(ref_world, wl.beam, None, False)
            ref = ref_world.fetch_dose()
            g1 = ref_world.run_stats()["calibration_factor"]
            ref_world.close()
            for a, b, name in zip(got, ref, ("dose", "variance", "events")):
                assert np.array_equal(a, b), f"{name} differs (multicast={multicast}): max |d| = {np.abs(a - b).max()}"
            assert ref[2].sum() > 0 and f1 > 0 and f2 > 0 and g1 > 0
        dist.barrier()
    if rank == 0:
        print("FUSED_OK kinds=%s ranks=%d" % (",".join(kinds), world_size), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

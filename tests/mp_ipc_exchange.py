"""Multi-process GPU check (launched by tests/test_gpu_beams_and_outputs.py with torchrun, one rank per GPU): the
library-managed pipelined exchange (dxb_exchange_export / _import: CUDA IPC peers, copy-engine slab pulls, slab reduce
-> dose, double-buffered tallies) must reproduce the single-GPU dose score bit for bit over four beams run back to back
without a read-out in between, the first one CT-calibrated."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import opendxmc_b200 as dx  # noqa: E402
from opendxmc_b200 import distributed as D  # noqa: E402

SEED = 4321


def main():
    rank, world_size, local_rank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    wl = dx.workloads.ct_spiral_patient(scale=4, histories=2_000_000, step_deg=5.0)
    world = wl.build_world(1, [local_rank])
    D.set_grid_sharded(world, wl.dim, wl.spacing, wl.density, wl.material, local_rank)
    world.set_calibration_histories(720_000)
    ex = D.PipelinedExchange(world, local_rank)
    world.set_seed(SEED)
    factors = [D.run_beam_pipelined(world, wl.beam, ex, use_beam_calibration=(k == 0)) for k in range(4)]
    # a new grid upload on the exchanging context (what an e2e step does), then one more beam
    D.set_grid_sharded(world, wl.dim, wl.spacing, wl.density, wl.material, local_rank)
    mine_before = world.fetch_dose_range(ex.begin, ex.end)
    assert mine_before[2][ex.begin:ex.end].sum() == 0, "set_grid must clear the dose score"
    world.set_seed(SEED)
    for k in range(4):
        D.run_beam_pipelined(world, wl.beam, ex, use_beam_calibration=(k == 0))
    mine = world.fetch_dose_range(ex.begin, ex.end)   # waits for the pending exchanges of this rank
    times = ex.times_ms()
    parts = [None] * world_size if rank == 0 else None
    dist.gather_object([a[ex.begin:ex.end] for a in mine], parts, dst=0)
    ex.close()
    world.close()
    if rank == 0:
        got = [np.concatenate([p[k] for p in parts]) for k in range(3)]
        ref_world = wl.build_world(1, [local_rank])
        ref_world.set_calibration_histories(720_000)
        ref_world.set_seed(SEED)
        tr = dx.Transport()
        ref_f = []
        for k in range(4):
            assert tr(ref_world, wl.beam, None, k == 0)
            ref_f.append(ref_world.run_stats()["calibration_factor"])
        ref = ref_world.fetch_dose()
        ref_world.close()
        assert ref_f == factors, (ref_f, factors)
        for a, b, name in zip(got, ref, ("dose", "variance", "events")):
            assert np.array_equal(a, b), f"{name} differs: {np.count_nonzero(a != b)} voxels, max |d| = {np.abs(a.astype(np.float64) - b.astype(np.float64)).max()}"
        assert ref[2].sum() > 0
        print("IPC_OK ranks=%d exchange_ms(pulls, reduce, clear)=%s" % (world_size, ["%.3f" % t for t in times]), flush=True)
    dist.barrier()
    # the baseline exchange: distributed.run_beam = shard tallies -> NCCL int64 reduce to rank 0 -> calibration + energy->dose
    # there (the reduce runs on torch's stream, energy->dose on the library's: run_beam orders the two)
    world = wl.build_world(1, [local_rank])
    world.set_history_range(rank, world_size)
    world.set_calibration_histories(720_000)
    world.set_seed(SEED)
    tally = D.tally_tensor(world, local_rank)
    f = [D.run_beam(world, wl.beam, tally, rank, use_beam_calibration=(k == 0)) for k in range(2)]
    if rank == 0:
        got = world.fetch_dose()
        ref_world = wl.build_world(1, [local_rank])
        ref_world.set_calibration_histories(720_000)
        ref_world.set_seed(SEED)
        tr = dx.Transport()
        for k in range(2):
            assert tr(ref_world, wl.beam, None, k == 0)
        ref = ref_world.fetch_dose()
        ref_world.close()
        for a, b, name in zip(got, ref, ("dose", "variance", "events")):
            assert np.array_equal(a, b), f"run_beam (NCCL reduce): {name} differs"
        assert f[0] > 0
        print("NCCL_REDUCE_OK", flush=True)
    world.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Manual GPU probe: are the kernel builds bit-identical on the full C2 volume, run to run and against each other?"""
import sys
sys.path.insert(0, ".")
import numpy as np
import opendxmc_b200 as dx

wl = dx.workloads.ct_spiral_patient(scale=1, histories=int(float(sys.argv[1])) if len(sys.argv) > 1 else 4_000_000)


def run(opts):
    world = wl.build_world(1, [0])
    for k, v in opts.items():
        world.set_option(k, v)
    dx.Transport().run_transport(world, wl.beam)
    got = [a.copy() for a in world.energy_scored()]
    st = world.run_stats()
    world.close()
    return got, st


ref, rst = run({"pool_slots": 0, "slots_per_lane": 0})
for opts in ({"pool_slots": 16}, {"pool_slots": 12}, {"pool_slots": 12, "pool_min_blocks": 5}, {"pool_slots": 12, "pool_min_blocks": 5},
             {"pool_slots": 16, "pool_min_blocks": 5}, {"pool_slots": 12, "pool_min_blocks": 6}, {"pool_slots": 8, "pool_min_blocks": 5}):
    got, st = run(opts)
    same = [bool(np.array_equal(a, b)) for a, b in zip(ref, got)]
    ndiff = int((ref[2] != got[2]).sum())
    ediff = int((ref[0] != got[0]).sum())
    print(opts, "identical", same, "voxels with different counts", ndiff, "different energy", ediff,
          "counters", {k: int(st[k]) - int(rst[k]) for k in ("histories", "steps", "interactions", "deposits")},
          "sumE rel", float((got[0].sum() - ref[0].sum()) / ref[0].sum()), flush=True)


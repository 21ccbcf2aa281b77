"""Manual GPU tuning run (not a test): the lane-multiplexed transport kernel against the register kernel.

1. bit-exactness: both kernels follow the same random-number protocol and sum fixed-point tallies, so their
   tallies must be identical word for word (C2 at 1/4 resolution, 2e6 histories);
2. throughput of option sets on the full C2 volume.

Usage: python profiles/tune_mux.py [histories] [sets]
"""
import sys
sys.path.insert(0, ".")
import numpy as np
import opendxmc_b200 as dx


def tallies(wl, opts, mode=1):
    world = wl.build_world(mode, [0])
    for k, v in opts.items():
        world.set_option(k, v)
    tr = dx.Transport()
    tr.run_transport(world, wl.beam)
    st = world.run_stats()
    e, e2, cnt = world.energy_scored()
    world.close()
    return np.asarray(e), np.asarray(e2), np.asarray(cnt), st


def check_exact():
    for name, wl in (("c2/4", dx.workloads.ct_spiral_patient(scale=4, histories=2_000_000)),
                     ("c1", dx.workloads.ctdi_body_phantom(n=32, histories=1_000_000, step_deg=5.0))):
        ref = tallies(wl, {"pool_slots": 0, "slots_per_lane": 0})
        for slots, pairs in ((4, 1), (4, 2), (3, 1), (2, 1), (6, 2)):
            got = tallies(wl, {"pool_slots": 0, "slots_per_lane": slots, "step_pairs": pairs})
            same = all(np.array_equal(a, b) for a, b in zip(ref[:3], got[:3]))
            keys = ("histories", "steps", "interactions", "deposits")
            print(f"exact {name} slots={slots} pairs={pairs}: tallies identical={same} "
                  f"counters identical={all(ref[3][k] == got[3][k] for k in keys)} sumE={got[0].sum():.6e}", flush=True)


def timing(wl, opts, tag, reps=2):
    world = wl.build_world(1, [0])
    for k, v in opts.items():
        world.set_option(k, v)
    tr = dx.Transport()
    best = None
    for _ in range(reps):
        tr.run_transport(world, wl.beam)
        st = world.run_stats()
        best = st if best is None or st["transport_ms"] < best["transport_ms"] else best
    st = best
    print(f"{tag:60s} ms={st['transport_ms']:8.2f} hist/s={st['histories'] / st['transport_ms'] * 1e3:.3e}", flush=True)
    world.close()


if __name__ == "__main__":
    nh = int(float(sys.argv[1])) if len(sys.argv) > 1 else 200_000_000
    which = sys.argv[2] if len(sys.argv) > 2 else "exact,base"
    if "exact" in which:
        check_exact()
    wl = dx.workloads.ct_spiral_patient(scale=1, histories=nh)
    if "base" in which:
        timing(wl, {"pool_slots": 0, "slots_per_lane": 0}, "register kernel (v4)")
        for slots in (2, 3, 4, 6):
            for pairs in (1, 2, 3):
                timing(wl, {"pool_slots": 0, "slots_per_lane": slots, "step_pairs": pairs}, f"mux slots={slots} pairs={pairs}")
    if "poolexact" in which:
        for name, w2 in (("c2/4", dx.workloads.ct_spiral_patient(scale=4, histories=2_000_000)),
                         ("c1", dx.workloads.ctdi_body_phantom(n=32, histories=1_000_000, step_deg=5.0))):
            ref = tallies(w2, {"pool_slots": 0, "slots_per_lane": 0})
            for slots, pairs in ((12, 1), (12, 2), (8, 1), (16, 2), (6, 1)):
                got = tallies(w2, {"pool_slots": slots, "step_pairs": pairs})
                same = all(np.array_equal(a, b) for a, b in zip(ref[:3], got[:3]))
                keys = ("histories", "steps", "interactions", "deposits")
                print(f"pool exact {name} slots={slots} pairs={pairs}: tallies identical={same} "
                      f"counters identical={all(ref[3][k] == got[3][k] for k in keys)}", flush=True)
    if "pooltime" in which:
        timing(wl, {"pool_slots": 0, "slots_per_lane": 0}, "register kernel (v4)")
        for slots in (6, 8, 12, 16):
            for pairs in (1, 2):
                timing(wl, {"pool_slots": slots, "step_pairs": pairs}, f"pool slots={slots} pairs={pairs}")
    if "poolpolicy" in which:
        for th, sw in ((128, 2), (192, 3), (256, 4), (384, 6), (512, 8), (512, 7)):
            for slots in (12, 16):
                timing(wl, {"pool_slots": slots, "pool_threads": th, "service_warps": sw},
                       f"pool{slots} threads={th} service={sw}", reps=2)
    if "poolocc" in which:
        # register budget / residency: 64 registers x 4 blocks vs 48 x 5 vs 40 x 6 (256 threads per block)
        w2 = dx.workloads.ct_spiral_patient(scale=4, histories=2_000_000)
        ref = tallies(w2, {"pool_slots": 0, "slots_per_lane": 0})
        for mb in (5, 6):
            got = tallies(w2, {"pool_slots": 12, "pool_min_blocks": mb})
            print(f"pool exact c2/4 min_blocks={mb}: tallies identical={all(np.array_equal(a, b) for a, b in zip(ref[:3], got[:3]))}", flush=True)
        for mb in (0, 5, 6):
            for slots in (8, 12, 16):
                for sw in (3, 4):
                    timing(wl, {"pool_slots": slots, "pool_min_blocks": mb, "service_warps": sw},
                           f"pool{slots} min_blocks={mb} service={sw}", reps=2)
    if which.startswith("psingle"):
        f = which.split(":")
        opts = {"pool_slots": int(f[1]), "step_pairs": int(f[2])}
        for kv in f[3:]:
            k, v = kv.split("=")
            opts[k] = float(v)
        timing(wl, opts, which, reps=1)
    if which.startswith("single"):
        # single:<slots>:<pairs>[:key=value...] — one configuration, one repetition (for ncu captures)
        f = which.split(":")
        opts = {"pool_slots": 0, "slots_per_lane": int(f[1]), "step_pairs": int(f[2])}
        for kv in f[3:]:
            k, v = kv.split("=")
            opts[k] = float(v)
        timing(wl, opts, which, reps=1)
    if "muxpolicy" in which:
        for bias in (-8, -4, 0, 4, 8):
            for rf in (4, 8, 16):
                timing(wl, {"slots_per_lane": 4, "step_pairs": 2, "interact_bias": bias, "refill_threshold": rf},
                       f"mux slots=4 pairs=2 bias={bias} refill={rf}")
        for rt in (4, 8, 16):
            timing(wl, {"slots_per_lane": 4, "step_pairs": 2, "rayleigh_threshold": rt}, f"mux slots=4 pairs=2 ray={rt}")
    if "blocks" in which:
        for slots, th, bps in ((4, 128, 6), (4, 128, 5), (4, 256, 2), (3, 128, 8), (3, 256, 3), (6, 128, 4), (6, 128, 3), (2, 256, 5)):
            timing(wl, {"slots_per_lane": slots, "step_pairs": 2, "threads_per_block": th, "blocks_per_sm": bps},
                   f"mux slots={slots} pairs=2 threads={th} blocks/SM={bps}")

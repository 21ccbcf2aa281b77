"""Quick manual GPU probe (not a test): timings and counters of the transport kernel on the bench workloads."""
import sys, time, itertools
sys.path.insert(0, ".")
import numpy as np
import opendxmc_b200 as dx


def run(wl, mode=1, reps=2, opts=None, tag=""):
    world = wl.build_world(mode, [0])
    for k, v in (opts or {}).items():
        world.set_option(k, v)
    tr = dx.Transport()
    best = None
    for r in range(reps):
        tr.run_transport(world, wl.beam)
        st = world.run_stats()
        best = st if best is None or st["transport_ms"] < best["transport_ms"] else best
    st = best
    print(f"{wl.name} {tag} transport_ms={st['transport_ms']:.2f} hist={st['histories']:.3e} "
          f"hist/s={st['histories']/st['transport_ms']*1e3:.3e} S={st['steps']/st['histories']:.2f} I={st['interactions']/st['histories']:.2f} "
          f"D={st['deposits']/st['histories']:.2f} launches={st['kernel_launches']}", flush=True)
    world.close()


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "c1"):
        run(dx.workloads.ctdi_body_phantom(n=64, histories=10_000_000))
    if which in ("all", "c2s"):
        run(dx.workloads.ct_spiral_patient(scale=4, histories=20_000_000))
    if which in ("all", "c2"):
        run(dx.workloads.ct_spiral_patient(scale=1, histories=100_000_000))
    if which == "sweep":
        wl = dx.workloads.ct_spiral_patient(scale=1, histories=100_000_000)
        for td, ti, trr in itertools.product((2, 4, 8), (8, 12, 16), (4, 8)):
            run(wl, opts={"refill_threshold": td, "interact_threshold": ti, "rayleigh_threshold": trr}, tag=f"TD={td} TI={ti} TR={trr}")
        for bps in (3, 4):
            for th in (128, 256):
                run(wl, opts={"blocks_per_sm": bps * (256 // th), "threads_per_block": th}, tag=f"bps={bps * (256 // th)} threads={th}")

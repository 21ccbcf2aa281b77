"""Quick manual GPU probe (not a test): C1 + C2-small timings and counters."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import opendxmc_b200 as dx

def run(wl, mode=1, reps=2):
    world = wl.build_world(mode, [0])
    tr = dx.Transport()
    for r in range(reps):
        t = time.time()
        tr.run_transport(world, wl.beam)
        st = world.run_stats()
        print(wl.name, f"rep{r} wall={time.time()-t:.3f}s transport_ms={st['transport_ms']:.2f} hist={st['histories']:.3e} "
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

"""Manual GPU run (not a test): transport throughput of the five BASELINE.json configurations (C1-C5), 1 GPU.
Usage: python profiles/bench_configs.py [histories]"""
import sys
sys.path.insert(0, ".")
import opendxmc_b200 as dx

nh = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
W = dx.workloads
cases = [("C1 CTDI body phantom 64^3, axial", lambda: W.ctdi_body_phantom(n=64, histories=nh), 1),
         ("C2 CT patient 512x512x300, spiral", lambda: W.ct_spiral_patient(scale=1, histories=nh), 1),
         ("C3 ICRP AM shape 254x127x222, chest spiral", lambda: W.icrp_phantom("AM", histories=nh), 1),
         ("C4 thorax 512x512x400, dual source + AEC", lambda: W.ct_dual_source_thorax(scale=1, histories=nh), 1),
         ("C5 ICRP 10y shape 419x226x576, DX 80 kV", lambda: W.icrp_phantom("10M", histories=nh, beam_kind="dx"), 1),
         ("C2 physics mode 0", lambda: W.ct_spiral_patient(scale=1, histories=nh), 0),
         ("C2 physics mode 2", lambda: W.ct_spiral_patient(scale=1, histories=nh), 2)]
for name, make, mode in cases:
    wl = make()
    world = wl.build_world(mode, [0])
    tr = dx.Transport()
    best = None
    for _ in range(2):
        tr.run_transport(world, wl.beam)
        st = world.run_stats()
        best = st if best is None or st["transport_ms"] < best["transport_ms"] else best
    st = best
    h = st["histories"]
    print(f"{name:48s} materials={len(wl.materials):3d} hist={h:.2e} ms={st['transport_ms']:8.2f} hist/s={h / st['transport_ms'] * 1e3:.3e} "
          f"S={st['steps'] / h:6.2f} I={st['interactions'] / h:5.2f} D={st['deposits'] / h:5.2f}", flush=True)
    world.close()

"""Manual GPU run (not a test): the dense-box build of the pool kernel on one GPU, transport kernel time only (CUDA events
inside the library).  Every BASELINE.json configuration with the default options (dense box / slab table in auto mode) and with
dense_box = 0.  Usage: python profiles/sweep_r02_densebox.py [histories]"""
import sys
sys.path.insert(0, ".")
import opendxmc_b200 as dx

nh = int(float(sys.argv[1])) if len(sys.argv) > 1 else 200_000_000
W = dx.workloads
cases = [("C2 CT patient 512x512x300, spiral", lambda: W.ct_spiral_patient(scale=1, histories=nh), 1, [{}, {"dense_box": 0}]),
         ("C4 thorax 512x512x400, dual source + AEC", lambda: W.ct_dual_source_thorax(scale=1, histories=nh), 1, [{}, {"dense_box": 0}]),
         ("C3 ICRP AM shape 254x127x222, chest spiral", lambda: W.icrp_phantom("AM", histories=nh), 1, [{}, {"local_majorant": 0}, {"local_majorant": 0, "dense_box": 0}]),
         ("C5 ICRP 10y shape 419x226x576, DX 80 kV", lambda: W.icrp_phantom("10M", histories=nh, beam_kind="dx"), 1, [{}, {"local_majorant": 0}]),
         ("C1 CTDI body phantom 64^3, axial", lambda: W.ctdi_body_phantom(n=64, histories=nh), 1, [{}]),
         ("C2 physics mode 0", lambda: W.ct_spiral_patient(scale=1, histories=nh), 0, [{}]),
         ("C2 physics mode 2", lambda: W.ct_spiral_patient(scale=1, histories=nh), 2, [{}])]
only = sys.argv[2] if len(sys.argv) > 2 else ""
for name, make, mode, variants in cases:
    if only and only not in name:
        continue
    wl = make()
    for opts in variants:
        world = wl.build_world(mode, [0])
        for k, v in opts.items():
            world.set_option(k, v)
        tr = dx.Transport()
        best = None
        for _ in range(3):
            tr.run_transport(world, wl.beam)
            st = world.run_stats()
            best = st if best is None or st["transport_ms"] < best["transport_ms"] else best
        st = best
        h = st["histories"]
        box = world.dense_box()
        part = 1.0
        for a in range(3):
            part *= (box["box"][a + 3] - box["box"][a]) / wl.dim[a]
        print(f"{name:46s} {str(opts):40s} hist={h:.2e} ms={st['transport_ms']:8.2f} hist/s={h / st['transport_ms'] * 1e3:.3e} "
              f"S={st['steps'] / h:6.2f} flights={st['hops'] / h:5.2f} I={st['interactions'] / h:5.2f} D={st['deposits'] / h:5.2f} "
              f"db={st['dense_box']} lm={st['local_majorant']} (box {part:.2f} of the grid, ratio@60keV {box['ratio'][378 >> 5]:.2e}, useful {box['useful']})", flush=True)
        world.close()

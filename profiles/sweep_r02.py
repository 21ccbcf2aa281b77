"""Manual GPU run (not a test): round-2 kernel experiments on one GPU, transport kernel time only (CUDA events inside the
library).  Usage: python profiles/sweep_r02.py [histories]
  * every BASELINE.json configuration with the default options (slab-local majorants in auto mode) and with
    local_majorant = 0 (global Woodcock majorant everywhere, the round-1 behaviour);
  * C2 with the prefetch variant of the quad step (step_quad = 2: no speculative load outstanding at the release fence);
  * the brick pre-filter (skips the gathers of certainly-virtual collisions; bit-identical results) off / 8 / 16 / 32 voxels."""
import sys
sys.path.insert(0, ".")
import opendxmc_b200 as dx

nh = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
W = dx.workloads
cases = [("C2 CT patient 512x512x300, spiral", lambda: W.ct_spiral_patient(scale=1, histories=nh), 1,
          [{}, {"brick_filter": 1, "brick_voxels": 16}, {"brick_filter": 1, "brick_voxels": 32}, {"step_quad": 2}, {"local_majorant": 1}]),
         ("C3 ICRP AM shape 254x127x222, chest spiral", lambda: W.icrp_phantom("AM", histories=nh), 1, [{}, {"local_majorant": 0}, {"slab_cm": 4.0}, {"slab_cm": 8.0}, {"slab_cm": 16.0}]),
         ("C5 ICRP 10y shape 419x226x576, DX 80 kV", lambda: W.icrp_phantom("10M", histories=nh, beam_kind="dx"), 1, [{}, {"local_majorant": 0}, {"slab_cm": 4.0}, {"slab_cm": 8.0}, {"slab_cm": 16.0}]),
         ("C4 thorax 512x512x400, dual source + AEC", lambda: W.ct_dual_source_thorax(scale=1, histories=nh), 1, [{}, {"brick_filter": 1, "brick_voxels": 16}]),
         ("C1 CTDI body phantom 64^3, axial", lambda: W.ctdi_body_phantom(n=64, histories=nh), 1, [{}]),
         ("C2 physics mode 0", lambda: W.ct_spiral_patient(scale=1, histories=nh), 0, [{}]),
         ("C2 physics mode 2", lambda: W.ct_spiral_patient(scale=1, histories=nh), 2, [{}])]
for name, make, mode, variants in cases:
    wl = make()
    for opts in variants:
        world = wl.build_world(mode, [0])
        for k, v in opts.items():
            world.set_option(k, v)
        if "slab_cm" in opts or "brick_voxels" in opts:
            world.build()  # the slab / brick tables are built with the grid
        tr = dx.Transport()
        best = None
        for _ in range(3):
            tr.run_transport(world, wl.beam)
            st = world.run_stats()
            best = st if best is None or st["transport_ms"] < best["transport_ms"] else best
        st = best
        h = st["histories"]
        n, shift, useful, _ = world.local_majorant()
        print(f"{name:46s} {str(opts):44s} mats={len(wl.materials):3d} hist={h:.2e} ms={st['transport_ms']:8.2f} hist/s={h / st['transport_ms'] * 1e3:.3e} "
              f"S={st['steps'] / h:6.2f} fetch={st['voxel_fetches'] / h:6.2f} hops={st['hops'] / h:5.2f} I={st['interactions'] / h:5.2f} D={st['deposits'] / h:5.2f} "
              f"lm={st['local_majorant']} (slabs {n}, 2^{shift} layers, useful {useful})", flush=True)
        world.close()

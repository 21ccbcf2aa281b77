"""Target of the compute-sanitizer runs (profiles/run_sanitizer_r02.sh): small transports that walk the device code paths -
the pool kernel in the three physics modes with the nested CTDI calibration (hole sums), the brick pre-filter build, the
slab-local-majorant build on an ICRP-shaped phantom, energy -> dose, post-processing and the per-organ dose."""
import sys
sys.path.insert(0, ".")
import numpy as np
import opendxmc_b200 as dx

nh = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000
tr = dx.Transport()
for mode in (0, 1, 2):
    wl = dx.workloads.ctdi_body_phantom(n=32, histories=nh, step_deg=10.0)
    world = wl.build_world(mode, [0])
    world.set_calibration_histories(100_000)
    tr(world, wl.beam, useBeamCalibration=True)
    d, v, n = world._item.doseArrays()
    print("C1 mode", mode, "dose sum", float(np.sum(d)), "events", int(np.sum(n)), flush=True)
    world.close()

wl = dx.workloads.ct_spiral_patient(scale=8, histories=nh, step_deg=10.0)
world = dx.World([0])
grid = world.addItem(dx.AAVoxelGrid(1))
assert grid.setData(wl.dim, wl.density, wl.material, wl.materials)
grid.setSpacing(wl.spacing)
world.build()
world.set_option("brick_filter", 1)
world.set_option("brick_voxels", 8)
world.set_option("local_majorant", 1)
world.set_option("slab_cm", 4.0)
world.build()
tr.run_transport(world, wl.beam)
print("C2/8 brick filter off (slab-local build wins):", world.run_stats()["local_majorant"], world.run_stats()["hops"], flush=True)
world.set_option("local_majorant", 0)
world.build()
tr(world, wl.beam, useBeamCalibration=False)
st = world.run_stats()
print("C2/8 brick filter: fetches", st["voxel_fetches"], "of", st["steps"], flush=True)
pp = world.dose_postprocessed(True)
print("postprocessed", float(np.sum(pp[0])), flush=True)
world.close()

wl = dx.workloads.icrp_phantom("AM", scale=4, histories=nh)
world = wl.build_world(1, [0])
tr(world, wl.beam, useBeamCalibration=False)
st = world.run_stats()
od = world.organ_dose(wl.organ, len(wl.organ_names))
print("C3/4 slab-local majorants", st["local_majorant"], "hops", st["hops"], "organ dose max", float(np.max(od[0])), flush=True)
world.close()

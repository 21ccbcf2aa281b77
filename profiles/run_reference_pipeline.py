"""Manual GPU run (opt-in test: DXB_RUN_REF_PIPELINE=1 pytest -m gpu -k reference_pipeline): OpenDXMC's own
SimulationPipeline / worker<CORRECTION>() (R:src/libopendxmc/simulationpipeline.cpp:124-235), compiled unmodified into
oracle/_ref/opendxmc_ref, runs a CT sequential beam on the reference's PMMA cylinder through the C++ shims; the same
world is then rebuilt through the Python mirror and must give the same dose, variance and event count after the
reference's post-processing (air mask, uGy rule) - bit for bit, since both sides drive the same library with the same
seed.  Usage: python profiles/run_reference_pipeline.py [mode] [delete_air] [histories per exposure]"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(mode=1, delete_air=1, per_exposure=20000):
    import opendxmc_b200 as dx
    exe = os.path.join(ROOT, "oracle", "_ref", "opendxmc_ref")
    with tempfile.TemporaryDirectory() as tmp:
        prefix = os.path.join(tmp, "ref")
        r = subprocess.run([exe, "run", str(mode), str(delete_air), str(per_exposure), prefix], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
        meta = json.load(open(prefix + ".json"))
        n = int(np.prod(meta["dim"]))
        dens = np.fromfile(prefix + ".density.bin", dtype=np.float64)
        mat = np.fromfile(prefix + ".material.bin", dtype=np.uint8)
        ref = [np.fromfile(prefix + f".{k}.bin", dtype=np.float64) for k in ("dose", "variance", "count")]
    assert dens.size == mat.size == n and all(a.size == n for a in ref)
    mats = [dx.Material.byWeight(dx.NISTMaterials.Composition(nm)) for nm in ("Air, Dry (near sea level)", "Polymethyl Methacralate (Lucite, Perspex)")]
    world = dx.World([0])
    grid = world.addItem(dx.AAVoxelGrid(mode))
    grid.setData(meta["dim"], dens, mat, mats)
    grid.setSpacing(meta["spacing"])
    world.build()
    beam = dx.CTSequentialBeam((0, 0, 0), (0, 0, 1), {13: 9.0})
    beam.setStepAngleDeg(10.0)
    beam.setNumberOfParticlesPerExposure(per_exposure)
    assert beam.numberOfExposures() == meta["exposures"]
    assert dx.Transport()(world, beam, None, True)
    d, v, c, units = world.dose_postprocessed(bool(delete_air))
    world.close()
    same = [bool(np.array_equal(a, b)) for a, b in zip((d, v, c), ref)]
    print(f"reference pipeline vs Python mirror: identical dose/variance/count = {same}, units {units} / {meta['dose_units']}, "
          f"sum dose {d.sum():.6e} vs {ref[0].sum():.6e}, events {int(c.sum())} vs {int(ref[2].sum())}", flush=True)
    return same, units, meta["dose_units"], (d, v, c), ref


if __name__ == "__main__":
    a = [int(x) for x in sys.argv[1:4]]
    run(*a)

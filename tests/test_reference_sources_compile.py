"""The strongest drop-in check available offline: OpenDXMC's OWN translation units for this boundary are compiled
UNMODIFIED from /root/reference against include/dxmc/ and linked with libdxmc_b200.so (oracle/Makefile.ref ->
oracle/_ref/opendxmc_ref, driver oracle/ref_driver.cpp; Qt / VTK replaced by tests/stubs/):

    dxmc_specialization.cpp   the `Beam` variant over all six beam types, the app's DXBeam subclass
    beamactorcontainer.cpp    exposure(i).position()/directionCosines()/collimationHalfAngles() of every beam type
    datacontainer.cpp         CTAECFilter, the water-equivalent-diameter AEC profile
    otherphantomimportpipeline.cpp  NISTMaterials::Composition / density, the PMMA cylinder
    ctsegmentationpipeline.cpp  Tube, Material::byNistName / attenuationValues: HU -> (material, density)
    simulationpipeline.cpp    worker<CORRECTION>(): World / AAVoxelGrid / Material / Transport / TransportProgress / doseScored
    basepipeline.cpp
  + syntax-only: icrpphantomimportpipeline.cpp.

The host-side parts are then RUN here (no GPU) and their numbers compared with the Python mirror (opendxmc_b200/api.py)
that the GPU tests and bench.py use.  Reads /root/reference, so it only runs where the reference tree is mounted."""
import json
import math
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/src/libopendxmc"
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "dxmc_specialization.cpp")), reason="reference tree not mounted")


@pytest.fixture(scope="module")
def ref_rows():
    r = subprocess.run(["make", "-f", os.path.join(ROOT, "oracle", "Makefile.ref")], cwd=ROOT, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    r = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "opendxmc_ref"), "host"], capture_output=True, text=True)
    assert r.returncode == 0, (r.stdout, r.stderr)
    return [json.loads(ln) for ln in r.stdout.strip().splitlines()]


@pytest.mark.parametrize("unit", ["icrpphantomimportpipeline.cpp"])
def test_reference_unit_compiles_unmodified(unit):
    r = subprocess.run(["g++", "-std=c++20", "-fsyntax-only", f"-I{ROOT}/tests/stubs", f"-I{ROOT}/include", f"-I{REF}", os.path.join(REF, unit)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_reference_dxbeam_matches_the_python_mirror(dx, ref_rows):
    rows = {d["tag"]: d for d in ref_rows if d["kind"] == "dxbeam"}

    def check(tag, b):
        d = rows[tag]
        c0, c1 = b.directionCosines()
        assert np.allclose(b.position(), d["pos"], rtol=0, atol=1e-12), (tag, b.position(), d["pos"])
        assert np.allclose(c0, d["c0"], rtol=0, atol=1e-14) and np.allclose(c1, d["c1"], rtol=0, atol=1e-14), tag
        assert np.allclose(b.collimationHalfAngles(), d["half"], rtol=1e-15, atol=0), tag
        assert np.allclose(b.collimation(), d["coll"], rtol=1e-14, atol=0), tag
        assert b.primaryAngleDeg() == pytest.approx(d["prim"], abs=1e-12) and b.secondaryAngleDeg() == pytest.approx(d["sec"], abs=1e-12)

    b = dx.DXBeam()
    check("default", b)
    b.setRotationCenter([1.0, 2.0, 3.0])
    b.setSourcePatientDistance(80.0)
    check("moved", b)
    b.setPrimaryAngleDeg(35.0)
    b.setSecondaryAngleDeg(-20.0)
    check("rotated", b)
    b.setSourceDetectorDistance(120.0)
    b.setCollimation([35.0, 43.0])
    check("collimated", b)
    b.setPrimaryAngleDeg(400.0)
    b.setSecondaryAngleDeg(-120.0)
    check("clamped", b)


def _corners(pos, cosines, half, scale):
    """R:src/libopendxmc/beamactorcontainer.cpp:44-69, restated"""
    c0, c1 = np.array(cosines[0]), np.array(cosines[1])
    d = np.cross(c0, c1)
    out = []
    for i, ys in enumerate((1, -1, -1, 1)):
        sx = math.sin(half[0]) * (1 if i < 2 else -1)
        sy = math.sin(half[1]) * ys
        sz = math.sqrt(1 - sx * sx - sy * sy)
        out.append(np.array(pos) + scale * (c0 * sx + c1 * sy + d * sz))
    return np.array(out)


def test_reference_beam_outlines_match_the_python_exposures(dx, ref_rows):
    """BeamActorContainer::update (the reference's code) walks exposure(i) of every beam type through the C++ shims;
    the Python mirror must give the same source positions and field corners."""
    rows = {d["tag"]: d for d in ref_rows if d["kind"] == "outline"}
    al9 = {13: 9.0}

    spiral = dx.CTSpiralBeam((0, 0, -4), (0, 0, 4), al9)
    spiral.setStepAngleDeg(30.0)
    cbct = dx.CBCTBeam((1, 2, 3), (0, 0, 1), {13: 2.0})
    cbct.setStepAngleDeg(20.0)
    seq = dx.CTSequentialBeam((0, 0, -2), (0, 0, 1), al9)
    seq.setStepAngleDeg(45.0)
    seq.setNumberOfSlices(2)
    for tag, b in (("spiral", spiral), ("cbct", cbct), ("sequential", seq)):
        pts = np.array(rows[tag]["points"])
        n = b.numberOfExposures()
        assert len(pts) == n + 4 and len(rows[tag]["cells"][0]) == n, tag
        mine = np.array([b.exposure(i).position() for i in range(n)])
        assert np.allclose(pts[:n], mine, rtol=0, atol=1e-11), tag
        e0 = b.exposure(0)
        assert np.allclose(pts[n:], _corners(e0.position(), e0.directionCosines(), e0.collimationHalfAngles(), b.sourceDetectorDistance()),
                           rtol=0, atol=1e-10), tag

    dual = dx.CTSpiralDualEnergyBeam((0, 0, -4), (0, 0, 4), al9)
    dual.setStepAngleDeg(30.0)
    pts = np.array(rows["dual"]["points"])
    nd = dual.numberOfExposures()
    n = nd // 2
    a = np.array([dual.exposure(2 * i).position() for i in range(n)])       # tube A = even exposures
    bb = np.array([dual.exposure(2 * i + 1).position() for i in range(n)])  # tube B = odd  (R:beamactorcontainer.cpp:134-146)
    assert np.allclose(pts[:n], a, rtol=0, atol=1e-11) and np.allclose(pts[n:2 * n], bb, rtol=0, atol=1e-11)

    pencil = np.array(rows["pencil"]["points"])
    assert np.allclose(pencil, [[0, -30, 0], [0, -10, 0]])

    b = dx.DXBeam()
    b.setRotationCenter([1.0, 2.0, 3.0])
    b.setSourcePatientDistance(80.0)
    b.setSourceDetectorDistance(120.0)
    b.setCollimation([35.0, 43.0])
    b.setPrimaryAngleDeg(400.0)
    b.setSecondaryAngleDeg(-120.0)
    pts = np.array(rows["dx"]["points"])
    assert np.allclose(pts[0], b.position(), atol=1e-12)
    assert np.allclose(pts[1:], _corners(b.position(), b.directionCosines(), b.collimationHalfAngles(), b.sourceDetectorDistance()), rtol=0, atol=1e-10)


def test_reference_wed_aec_profile_matches_the_python_mirror(dx, ref_rows):
    """DataContainer::calculateAECfilterFromWaterEquivalentDiameter (R:src/libopendxmc/datacontainer.cpp:42-100) run on
    the reference's own PMMA cylinder, against workloads.wed_aec_profile + CTAECFilter of the Python mirror."""
    d = [r for r in ref_rows if r["kind"] == "wed"][0]
    nx, ny, nz = d["dim"]
    sp = d["spacing"]
    # the reference's cylinder rule (R:src/libopendxmc/otherphantomimportpipeline.cpp:32-50)
    i = np.arange(nx) - nx / 2.0
    j = np.arange(ny) - ny / 2.0
    inside = (i[None, :] ** 2 + j[:, None] ** 2) <= min(nx / 2.0, ny / 2.0) ** 2
    assert int(inside.sum()) == d["pmma_voxels_per_slice"]
    air, pmma = dx.NISTMaterials.density("Air, Dry (near sea level)"), dx.NISTMaterials.density("Polymethyl Methacralate (Lucite, Perspex)")
    assert [air, pmma] == pytest.approx(d["density"], rel=1e-15)
    dens = np.where(inside, pmma, air)[None, :, :].repeat(nz, axis=0).reshape(-1)
    w = dx.workloads.wed_aec_profile(dens, (nx, ny, nz), sp)
    assert np.allclose(np.log(w) / 0.2, d["wed"], rtol=1e-13)
    half = sp[2] * nz / 2.0
    f = dx.CTAECFilter((0, 0, -half), (0, 0, half), list(w))
    assert np.allclose(f.weights(), d["aec_weights"], rtol=1e-12)
    assert f.start()[2] == pytest.approx(d["aec_start_z"]) and f.stop()[2] == pytest.approx(d["aec_stop_z"]) and f.isEmpty() == d["aec_empty"]


def test_reference_ct_segmentation_matches_the_restatement_and_the_oracle(dx, orc, ref_rows):
    """CTSegmentationPipeline::updateImageData (the reference's code, R:src/libopendxmc/ctsegmentationpipeline.cpp:61-169)
    on a HU ramp against the numpy restatement the GPU test uses for dxb_segment_ct and against the oracle's rule."""
    import ctypes as C
    d = [r for r in ref_rows if r["kind"] == "segmentation"][0]
    hu = d["hu0"] + d["hu_step"] * np.arange(d["n"]) + d["hu_offset"]
    ref_mat = np.array([int(ch) for ch in d["material"]], dtype=np.uint8)
    ref_dens = np.array(d["density"])
    names = ["Air, Dry (near sea level)", "Adipose Tissue (ICRP)", "Tissue, Soft (ICRP)", "Muscle, Skeletal", "Bone, Cortical (ICRP)"]
    assert d["materials"] == names
    tube = dx.Tube(120.0)
    tube.setAlFiltration(9.0)
    mats = [dx.Material.byNistName(nm) for nm in names]
    dn = [dx.NISTMaterials.density(nm) for nm in names]
    dn[-1] = 1.09
    e = tube.getEnergy()
    w = tube.getSpecter(e, True)
    water, air = dx.Material.byNistName("Water, Liquid"), mats[0]
    wd, ad = dx.NISTMaterials.density("Water, Liquid"), dx.NISTMaterials.density(names[0])
    uw = np.array([water.attenuationValues(x).sum() for x in e])
    ua = np.array([air.attenuationValues(x).sum() for x in e])
    um = [np.array([m.attenuationValues(x).sum() for x in e]) for m in mats]
    HU = [1000 * np.sum(w * (um[i] * dn[i] - uw * wd) / (uw * wd - ua * ad)) for i in range(5)]
    sep = np.array([(HU[i] + HU[i + 1]) / 2 for i in range(4)])
    att = np.array([np.sum(w * um[i]) for i in range(5)])
    omat = np.zeros(hu.size, dtype=np.uint8)
    odens = np.zeros(hu.size)
    dp = C.POINTER(C.c_double)
    orc.load().orc_segment(hu.ctypes.data_as(dp), hu.size, sep.ctypes.data_as(dp), 4, att.ctypes.data_as(dp),
                           float(np.sum(w * uw) * wd), float(np.sum(w * ua) * ad), omat.ctypes.data_as(C.POINTER(C.c_uint8)), odens.ctypes.data_as(dp))
    assert set(np.unique(ref_mat)) == {0, 1, 2, 3, 4}
    assert np.array_equal(ref_mat, omat)
    assert np.allclose(ref_dens, odens, rtol=1e-11, atol=1e-14)

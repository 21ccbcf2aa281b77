"""The strongest drop-in check available offline: OpenDXMC's OWN translation units for this boundary are compiled
UNMODIFIED from /root/reference against include/dxmc/ and linked with libdxmc_b200.so (oracle/Makefile.ref ->
oracle/_ref/opendxmc_ref, driver oracle/ref_driver.cpp; Qt / VTK replaced by tests/stubs/):

    dxmc_specialization.cpp   the `Beam` variant over all six beam types, the app's DXBeam subclass
    beamactorcontainer.cpp    exposure(i).position()/directionCosines()/collimationHalfAngles() of every beam type
    datacontainer.cpp         CTAECFilter, the water-equivalent-diameter AEC profile
    otherphantomimportpipeline.cpp  NISTMaterials::Composition / density, the PMMA cylinder
    icrpphantomimportpipeline.cpp   organ / media tables -> organ, material, density arrays
    ctsegmentationpipeline.cpp  Tube, Material::byNistName / attenuationValues: HU -> (material, density)
    dosetablepipeline.cpp     per-organ voxels / volume / mass / dose
    beamsettingsmodel.cpp     EVERY getter and setter of the six beam types, tube, bowtie, AEC, organ AEC (1800 lines)
    bowtiefilterreader.cpp    the 41 bowtie filters of data/bowtiefilters/bowtiefilters.json (over a small JSON parser)
    hdf5wrapper.cpp           save / load of the scene (radian accessors, filters, parseCompoundStr) over an in-memory HDF5 stand-in
    simulationpipeline.cpp    worker<CORRECTION>(): World / AAVoxelGrid / Material / Transport / TransportProgress / doseScored
    basepipeline.cpp

The host-side parts are then RUN here (no GPU) and their numbers compared with the Python mirror (opendxmc_b200/api.py)
that the GPU tests and bench.py use; the simulation pipeline itself runs end to end over a CPU test double of the
context-level C ABI (oracle/cpu_double.cpp).  Reads /root/reference, so it only runs where the reference tree is mounted."""
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


@pytest.mark.parametrize("case", ["all_organs", "sparse", "remove_arms"])
def test_reference_icrp_import_matches_the_python_mirror(dx, ref_rows, tmp_path, case):
    """ICRPPhantomImportPipeline::importPhantom (the reference's code, R:src/libopendxmc/icrpphantomimportpipeline.cpp:258-351)
    with the reference's real AM organ / media tables on a small organ array, against workloads.import_icrp_tables."""
    data = "/root/reference/data/phantoms/icrp/AM"
    if case == "all_organs":
        raw = np.concatenate([np.arange(141, dtype=np.uint8), np.arange(141, dtype=np.uint8)[::-1], np.zeros(18, dtype=np.uint8)])
    else:
        rng = np.random.default_rng(11)
        raw = rng.choice(np.array([0, 3, 9, 17, 29, 30, 61, 88, 95, 96, 97, 120, 139, 140], dtype=np.uint8), 300).astype(np.uint8)
    path = tmp_path / "organs.bin"
    raw.tofile(path)
    remove = case == "remove_arms"
    r = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "opendxmc_ref"), "icrp", str(path), f"{data}/AM_organs.dat", f"{data}/AM_media.dat",
                        "30", "5", "2", "1" if remove else "0"], capture_output=True, text=True)
    assert r.returncode == 0, (r.stdout[:200], r.stderr)
    ref = json.loads(r.stdout)
    mine_raw = raw.copy()
    if remove:
        # the reference blanks organs whose name contains arm / hand / Humeri / Ulnae (:274-295)
        t = dx.workloads.icrp_tables()["AM"]
        for o in t["organs"]:
            if any(k in o["name"] for k in ("arm", "hand", "Humeri", "Ulnae")):
                mine_raw[mine_raw == o["id"]] = 0
    organ, names, material, density, media_names, comps = dx.workloads.import_icrp_tables("AM", mine_raw)
    assert np.array_equal(organ, ref["organ"])
    assert names == ref["organ_names"]
    assert np.array_equal(material, ref["material"])
    assert np.array_equal(density, np.array(ref["density"]))
    assert media_names == [m["name"] for m in ref["materials"]]
    for c, m in zip(comps, ref["materials"]):
        rz = {int(z): w for z, w in m["Z"].items() if w > 0}
        assert set(c) == set(rz) and all(c[z] == pytest.approx(rz[z], rel=1e-12) for z in c), m["name"]
    assert ref["spacing"] == pytest.approx([0.1, 0.1, 0.1])  # mm -> cm (setSpacingInmm)


def test_reference_dose_table_matches_the_oracle(orc, ref_rows, tmp_path):
    """DoseTablePipeline::updateImageData (the reference's code, R:src/libopendxmc/dosetablepipeline.cpp:36-95) against the
    oracle's per-organ dose (the GPU test compares dxb_organ_dose with the same oracle function)."""
    rng = np.random.default_rng(5)
    dim, sp = (12, 10, 8), (0.11, 0.07, 0.2)
    n = dim[0] * dim[1] * dim[2]
    organ = rng.integers(0, 9, n).astype(np.uint8)
    organ[organ == 4] = 3                                   # organ 4 has no voxels: no table row
    dose = rng.random(n) * 3.0
    dens = rng.random(n) + 0.05
    prefix = str(tmp_path / "t")
    organ.tofile(prefix + ".organ.bin")
    dose.tofile(prefix + ".dose.bin")
    dens.tofile(prefix + ".density.bin")
    r = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "opendxmc_ref"), "dosetable", prefix, *map(str, dim), *map(repr, sp), "10"],
                       capture_output=True, text=True)
    assert r.returncode == 0, (r.stdout[:200], r.stderr)
    t = json.loads(r.stdout)
    assert t["header"][:4] == ["Name", "# Voxels", "Volume cm3", "Mass g"] and t["header"][4].startswith("Dose ")
    vol = sp[0] * sp[1] * sp[2]
    d, m, c = orc.organ_dose(dose, dens, organ, vol, 10)
    rows = {row["organ"]: row for row in t["rows"]}
    assert sorted(rows) == [o for o in range(10) if c[o] > 0] and 4 not in rows and 9 not in rows
    for o, row in rows.items():
        assert row["name"] == f"organ {o}" and row["voxels"] == c[o]
        assert row["volume"] == pytest.approx(c[o] * vol, rel=1e-13)
        assert row["mass"] == pytest.approx(m[o], rel=1e-12) and row["dose"] == pytest.approx(d[o], rel=1e-12)


def _beam_rows(dx, edited):
    """the Python mirror of what BeamSettingsModel shows: {row path: value}; the same defaults the model applies to a new
    beam (R:src/libopendxmc/beamsettingsmodel.cpp:470-474, 729-735, 929-933, 1166-1170, 1412-1416) and, if `edited`, the
    same edits oracle/ref_driver.cpp makes through the model's setters."""
    rows = {}

    def tube_rows(prefix, b, names=("Al", "Cu", "Sn", "Ag")):
        t = b.tube()
        rows[f"{prefix}/Tube potential [kV]"] = t.voltage()
        rows[f"{prefix}/Tube anode angle [deg]"] = t.anodeAngleDeg()
        for nm, z in (("Al", 13), ("Cu", 29), ("Sn", 50), ("Ag", 47)):
            if nm in names:
                rows[f"{prefix}/Tube {nm} filtration [mm]"] = t.filtration(z)
        rows[f"{prefix}/Tube half value layer [mmAl]"] = b.tubeAlHalfValueLayer()
        rows[f"{prefix}/Tube mean energy [keV]"] = b.tubeMeanSpecterEnergy()

    def organ_rows(prefix, b):
        o = b.organAECFilter()
        rows[f"{prefix}/Organ AEC/Use Organ AEC"] = o.useFilter()
        rows[f"{prefix}/Organ AEC/Start angle [deg]"] = o.startAngleDeg()
        rows[f"{prefix}/Organ AEC/Stop angle [deg]"] = o.stopAngleDeg()
        rows[f"{prefix}/Organ AEC/Ramp angle [deg]"] = o.rampAngleDeg()
        rows[f"{prefix}/Organ AEC/Low weight [0-1]"] = o.lowWeight()
        rows[f"{prefix}/Organ AEC/High weight"] = o.maxWeight()
        rows[f"{prefix}/Organ AEC/Compensate outside beam"] = o.compensateOutside()

    def count_rows(prefix, b, jobs=False):
        rows[f"{prefix}/Number of {'jobs' if jobs else 'exposures'}"] = b.numberOfExposures()
        rows[f"{prefix}/Particles per {'job' if jobs else 'exposure'}"] = b.numberOfParticlesPerExposure()
        rows[f"{prefix}/Total number of particles"] = b.numberOfParticles()

    b = dx.DXBeam(filtration={13: 2.0, 29: 0.1})
    b.setNumberOfParticlesPerExposure(1_000_000)
    if edited:
        b.setTubeVoltage(80.0)
        b.addTubeFiltrationMaterial(13, 3.5)
        b.setCollimation([30.0, 25.0])
        b.setPrimaryAngleDeg(40.0)
        b.setSourcePatientDistance(75.0)
    p = "DX Beam"
    rows[f"{p}/Source rotation center distance [cm]"] = b.sourcePatientDistance()
    rows[f"{p}/Primary angle [deg]"] = b.primaryAngleDeg()
    rows[f"{p}/Secondary angle [deg]"] = b.secondaryAngleDeg()
    rows[f"{p}/Source detector distance [cm]"] = b.sourceDetectorDistance()
    rows[f"{p}/Collimation [cm x cm]"] = list(b.collimation())
    rows[f"{p}/Collimation angles [deg]"] = [2 * a for a in b.collimationHalfAnglesDeg()]
    tube_rows(f"{p}/Tube", b)
    rows[f"{p}/Dose area product [mGycm^2]"] = b.DAPvalue()
    count_rows(p, b, jobs=True)

    b = dx.CBCTBeam((0, 0, 0), (0, 0, 1), {13: 2.0, 29: 0.1})
    b.setCollimationHalfAnglesDeg(10, 10)
    b.setNumberOfParticlesPerExposure(1_000_000)
    if edited:
        b.setStepAngleDeg(2.0)
        b.setStopAngleDeg(200.0)
    p = "CBCT Beam"
    rows[f"{p}/Isocenter [cm]"] = list(b.isocenter())
    rows[f"{p}/Rotation axis"] = list(b.rotationAxis())
    rows[f"{p}/Source detector distance [cm]"] = b.sourceDetectorDistance()
    rows[f"{p}/Set start angle [deg]"] = b.startAngleDeg()
    rows[f"{p}/Set stop angle [deg]"] = b.stopAngleDeg()
    rows[f"{p}/Set angle step [deg]"] = b.stepAngleDeg()
    rows[f"{p}/Collimation angles [deg]"] = [2 * a for a in b.collimationHalfAnglesDeg()]
    rows[f"{p}/Collimation [cm x cm]"] = [2 * b.sourceDetectorDistance() * math.tan(a) for a in b.collimationHalfAngles()]  # :842-851
    tube_rows(f"{p}/Tube", b)
    rows[f"{p}/Dose area product [mGycm^2]"] = b.DAPvalue()
    count_rows(p, b)

    b = dx.PencilBeam()
    if edited:
        b.setEnergy(75.0)
        b.setDirection([0, 1, 0])
    p = "Pencil Beam"
    rows[f"{p}/Position (x, y, z) [cm]"] = list(b.position())
    rows[f"{p}/Direction normal (x, y, z)"] = list(b.direction())
    rows[f"{p}/Photon energy [keV]"] = b.energy()
    rows[f"{p}/Air KERMA [mGy]"] = b.airKerma()
    count_rows(p, b, jobs=True)

    def ct_common(b):
        b.setSourceDetectorDistance(119)
        b.setCollimation(3.84)
        b.setStepAngleDeg(5)
        b.setNumberOfParticlesPerExposure(1_000_000)

    b = dx.CTSpiralBeam((0, 0, -20), (0, 0, 20), {13: 9.0})
    ct_common(b)
    if edited:
        b.setPitch(1.4)
        b.setStepAngleDeg(10.0)
        b.setCollimation(2.0)
        b.setCTDIvol(7.5)
        b.setStopPosition([0, 0, 10])
        b.addTubeFiltrationMaterial(50, 0.4)
        b.organAECFilter().setLowWeight(0.3)
        b.organAECFilter().setStopAngleDeg(120.0)
    p = "CT Spiral Beam"
    rows[f"{p}/Start position [cm]"] = list(b.startPosition())
    rows[f"{p}/Stop position [cm]"] = list(b.stopPosition())
    rows[f"{p}/Scan FOV [cm]"] = b.scanFieldOfView()
    rows[f"{p}/Source detector distance [cm]"] = b.sourceDetectorDistance()
    rows[f"{p}/Total collimation [cm]"] = b.collimation()
    rows[f"{p}/Set start angle [deg]"] = b.startAngleDeg()
    rows[f"{p}/Set angle step [deg]"] = b.stepAngleDeg()
    rows[f"{p}/Pitch"] = b.pitch()
    rows[f"{p}/CTDIvol [mGy]"] = b.CTDIvol()
    rows[f"{p}/CTDI phantom diameter [cm]"] = b.CTDIdiameter()
    tube_rows(f"{p}/Tube", b)
    organ_rows(p, b)
    count_rows(p, b)

    b = dx.CTSequentialBeam((0, 0, 0), (0, 0, 1), {13: 9.0})
    ct_common(b)
    if edited:
        b.setNumberOfSlices(7)
        b.setSliceSpacing(1.5)
        b.setCTDIw(3.0)
    p = "CT Sequential Beam"
    rows[f"{p}/Start position [cm]"] = list(b.position())
    rows[f"{p}/Direction vector"] = list(b.scanNormal())
    rows[f"{p}/Number of slices"] = b.numberOfSlices()
    rows[f"{p}/Slice spacing [cm]"] = b.sliceSpacing()
    rows[f"{p}/Scan FOV [cm]"] = b.scanFieldOfView()
    rows[f"{p}/Source detector distance [cm]"] = b.sourceDetectorDistance()
    rows[f"{p}/Total collimation [cm]"] = b.collimation()
    rows[f"{p}/Set start angle [deg]"] = b.startAngleDeg()
    rows[f"{p}/Set angle step [deg]"] = b.stepAngleDeg()
    rows[f"{p}/CTDIw [mGy]"] = b.CTDIw()
    rows[f"{p}/CTDI phantom diameter [cm]"] = b.CTDIdiameter()
    tube_rows(f"{p}/Tube", b)
    organ_rows(p, b)
    count_rows(p, b)

    b = dx.CTSpiralDualEnergyBeam((0, 0, -20), (0, 0, 20), {13: 9.0})
    ct_common(b)
    if edited:
        b.setTubeBVoltage(140.0)
        b.setRelativeMasTubeB(2.5)
        b.setPitch(3.0)
        b.setTubeBoffsetAngleDeg(95.0)
        b.setScanFieldOfViewB(30.0)
    p = "CT Spiral Dual Energy Beam"
    rows[f"{p}/Start position [cm]"] = list(b.startPosition())
    rows[f"{p}/Stop position [cm]"] = list(b.stopPosition())
    rows[f"{p}/Scan FOV Tube A [cm]"] = b.scanFieldOfViewA()
    rows[f"{p}/Scan FOV Tube B [cm]"] = b.scanFieldOfViewB()
    rows[f"{p}/Source detector distance [cm]"] = b.sourceDetectorDistance()
    rows[f"{p}/Total collimation [cm]"] = b.collimation()
    rows[f"{p}/Set start angle [deg]"] = b.startAngleDeg()
    rows[f"{p}/Set angle step [deg]"] = b.stepAngleDeg()
    rows[f"{p}/Tube B offset angle [deg]"] = b.tubeBoffsetAngleDeg()
    rows[f"{p}/Pitch"] = b.pitch()
    rows[f"{p}/CTDIvol [mGy]"] = b.CTDIvol()
    rows[f"{p}/CTDI phantom diameter [cm]"] = b.CTDIdiameter()
    for ab, tube, mas, hvl, wgt, mean in (("A", b.tubeA(), b.relativeMasTubeA(), b.tubeAAlHalfValueLayer(), b.tubeRelativeWeightA(), b.tubeAMeanSpecterEnergy()),
                                          ("B", b.tubeB(), b.relativeMasTubeB(), b.tubeBAlHalfValueLayer(), b.tubeRelativeWeightB(), b.tubeBMeanSpecterEnergy())):
        q = f"{p}/Tube {ab}"
        rows[f"{q}/Tube potential [kV]"] = tube.voltage()
        rows[f"{q}/Relative tube current"] = mas
        rows[f"{q}/Tubes anode angle [deg]"] = tube.anodeAngleDeg()
        for nm, z in (("Al", 13), ("Cu", 29), ("Sn", 50)):
            rows[f"{q}/Tube {nm} filtration [mm]"] = tube.filtration(z)
        rows[f"{q}/Tube half value layer [mmAl]"] = hvl
        rows[f"{q}/Relative photon count weight"] = wgt
        rows[f"{q}/Tube mean energy [keV]"] = mean
    organ_rows(p, b)
    count_rows(p, b)
    return rows


@pytest.mark.parametrize("tag", ["defaults", "edited"])
def test_reference_beam_settings_model_matches_the_python_mirror(dx, ref_rows, tag):
    """BeamSettingsModel (the reference's code, R:src/libopendxmc/beamsettingsmodel.cpp, all 1800 lines of getters and
    setters over the six beam types) runs on the C++ shims; every row it shows must equal the Python mirror."""
    r = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "opendxmc_ref"), "beammodel"], capture_output=True, text=True, cwd="/root/reference")
    assert r.returncode == 0 and "25 of 25 edits applied" in r.stderr, r.stderr
    dumps = {}
    for block in r.stdout.split("{\"kind\": \"beammodel\"")[1:]:
        d = json.loads("{\"kind\": \"beammodel\"" + block)
        dumps[d["tag"]] = d
    ref = dict(map(tuple, dumps[tag]["rows"]))
    assert dumps[tag]["beams"] == 6 and len(ref) >= 140
    mine = _beam_rows(dx, tag == "edited")
    missing = [k for k in mine if k not in ref]
    assert not missing, missing
    for k, v in mine.items():
        rv = ref[k]
        if isinstance(v, (list, tuple)):
            assert [float(x) for x in rv.split(",")] == pytest.approx([float(x) for x in v], abs=2e-6), k   # QString::number: 6 decimals
        elif isinstance(v, bool):
            assert rv is v, k
        else:
            assert rv == pytest.approx(float(v), rel=1e-12, abs=1e-12), (k, rv, v)
    # rows the mirror does not model are only the bowtie / AEC selectors
    extra = sorted(k for k in ref if k not in mine)
    assert all(("Bowtie filter" in k) or ("Use current AEC profile" in k) or k.endswith("Rotation center (x, y, z) [cm]") for k in extra), extra


def test_reference_hdf5_wrapper_round_trip_over_the_shims(ref_rows, tmp_path):
    """HDF5Wrapper (the reference's code, R:src/libopendxmc/hdf5wrapper.cpp, compiled unmodified) saves the scene to a REAL
    file in the HDF5 format and loads it back - tests/stubs/H5Cpp.h maps the HDF5 C++ API onto libdxmc_b200's own reader /
    writer (include/dxb.h: dxb_h5_*): every setting of the loadable beam types survives (radian accessors, tube
    filtration, organ AEC ...), and so do the grid and the materials (AtomHandler::toSymbol -> parseCompoundStr).  The
    file is then inspected independently of the reference's code."""
    import ctypes as C
    from opendxmc_b200 import _capi as K
    path = tmp_path / "scene.h5"
    r = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "opendxmc_ref"), "h5roundtrip", str(path)], capture_output=True, text=True, cwd="/root/reference")
    assert r.returncode == 0, r.stderr
    head = json.loads(r.stdout.splitlines()[0])
    # the file the reference's save() produced: HDF5 signature, arrays in z-y-x order (:121-124), one deflated chunk for
    # the big arrays (:145-151), variable-length strings, beam groups numbered from 1 with attributes
    raw = path.read_bytes()
    assert raw[:8] == b"\x89HDF\r\n\x1a\n"
    lib = K.load()
    h = K.VP()
    assert lib.dxb_h5_open(C.byref(h), str(path).encode()) == K.DXB_OK
    t, rk, d, z = C.c_int(), C.c_int(), (C.c_uint64 * 8)(), C.c_int()
    dims = np.zeros(3, dtype=np.uint64)
    assert lib.dxb_h5_dataset_read(h, b"/dimensions", dims.ctypes.data_as(K.VP), 24) == K.DXB_OK
    assert list(dims) == [24, 24, 5]
    for name, code in (("densityarray", 1), ("materialarray", 3), ("organarray", 3)):
        assert lib.dxb_h5_dataset_info(h, ("/" + name).encode(), C.byref(t), C.byref(rk), d, C.byref(z)) == K.DXB_OK
        assert (t.value, rk.value, list(d)[:3], z.value) == (code, 3, [5, 24, 24], 1), name
    assert lib.dxb_h5_dataset_info(h, b"/materialnames", C.byref(t), C.byref(rk), d, C.byref(z)) == K.DXB_OK and t.value == 4
    assert lib.dxb_h5_dataset_string(h, b"/materialnames", 0) == b"Air, Dry (near sea level)"
    assert lib.dxb_h5_dataset_string(h, b"/materialcomposition", 1).startswith(b"H")   # PMMA: H, C, O
    groups = lib.dxb_h5_list(h, b"/beams").decode().split()
    assert "CTSpiralBeams" in groups and "DXBeams" in groups and "CTSpiralDualEnergyBeams" in groups
    attrs = lib.dxb_h5_list(h, b"/beams/CTSpiralBeams/1").decode()
    assert "a pitch" in attrs or "a " in attrs
    lib.dxb_h5_close(h)
    assert head["same_grid"] and head["same_materials"] and head["worst_composition_rel"] < 1e-12
    # the pencil beam has no save() overload (R:hdf5wrapper.cpp:1050-1068)
    assert head["beams_saved"] == 5
    # two defects of the reference itself, reproduced as they are: the dual-energy beam is saved under
    # "CTSpiralDualEnergyBeams" (:588) but looked for under "CTDualEnergySpiralBeams" (:1159), and the AEC weights are
    # saved as "aecweights" (:448) but read back as "weights" (:1141)
    assert head["beams_loaded"] == 4 and head["same_aec"] is False
    dumps = {}
    for block in r.stdout.split("{\"kind\": \"beammodel\"")[1:]:
        d = json.loads("{\"kind\": \"beammodel\"" + block)
        dumps[d["tag"]] = dict(map(tuple, d["rows"]))
    saved, loaded = dumps["saved"], dumps["loaded"]
    assert {k.split("/")[0] for k in loaded} == {"DX Beam", "CBCT Beam", "CT Spiral Beam", "CT Sequential Beam"}
    assert len(loaded) >= 95
    for k, v in loaded.items():
        assert saved[k] == v, (k, saved[k], v)


def test_reference_bowtie_reader_matches_the_python_reader(dx, ref_rows):
    """BowtieFilterReader::read (the reference's code, R:src/libopendxmc/bowtiefilterreader.cpp:34-110) on the reference's
    data/bowtiefilters/bowtiefilters.json - 41 filters with unsorted points - against workloads.read_bowtie_filters, and
    the shim's BowtieFilter::operator() against the Python mirror at sample fan angles."""
    r = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "opendxmc_ref"), "bowtie"], capture_output=True, text=True, cwd="/root/reference")
    assert r.returncode == 0, r.stderr
    ref = {f["name"]: f for f in json.loads(r.stdout)["filters"]}
    mine = dx.workloads.read_bowtie_filters("/root/reference/data/bowtiefilters/bowtiefilters.json")
    assert len(ref) == 41 and set(ref) == set(mine) and "Siemens Definition Flash W1 120kV" in ref
    packaged = dx.workloads.read_bowtie_filters()          # the copy shipped with the package must be the same table
    assert set(packaged) == set(mine)
    for name, f in ref.items():
        b = mine[name]
        assert np.array_equal(np.array(f["data"]), np.stack([b.angle, b.weight], axis=1)), name
        assert np.array_equal(packaged[name].angle, b.angle) and np.array_equal(packaged[name].weight, b.weight), name
        assert [b(0.05 * k) for k in range(10)] == pytest.approx(f["weights"], rel=1e-13), name


def _run_beam(dx, kind, per_exposure, level, wed_weights, half_z):
    """the Python mirror of oracle/ref_driver.cpp: makeRunBeam"""
    if kind == "sequential":
        b = dx.CTSequentialBeam((0, 0, 0), (0, 0, 1), {13: 9.0})
        b.setStepAngleDeg(10.0)
        b.setCTDIw(level)
    elif kind == "spiral":
        b = dx.CTSpiralBeam((0, 0, -6), (0, 0, 6), {13: 7.0, 29: 0.05})
        b.setStepAngleDeg(15.0)
        b.setPitch(1.2)
        b.setCollimation(2.4)
        b.setScanFieldOfView(40.0)
        b.setStartAngleDeg(30.0)
        b.setTubeVoltage(100.0)
        b.setCTDIvol(level)
        b.setBowtieFilter(dx.BowtieFilter([(0.0, 1.0), (0.1, 0.8), (0.2, 0.5), (0.3, 0.25), (0.39, 0.1)]))
        b.setAECFilter(dx.CTAECFilter((0, 0, -half_z), (0, 0, half_z), list(wed_weights)))
        o = b.organAECFilter()
        o.setUseFilter(True)
        o.setStartAngleDeg(300.0)
        o.setStopAngleDeg(60.0)
        o.setRampAngleDeg(25.0)
        o.setLowWeightFactor(0.4)
    elif kind == "dual":
        b = dx.CTSpiralDualEnergyBeam((0, 0, -5), (0, 0, 5), {13: 9.0})
        b.setStepAngleDeg(20.0)
        b.setPitch(2.0)
        b.setCollimation(3.0)
        b.setTubeAVoltage(80.0)
        b.setTubeBVoltage(140.0)
        b.addTubeBFiltrationMaterial(50, 0.4)
        b.setRelativeMasTubeB(0.6)
        b.setTubeBoffsetAngleDeg(95.0)
        b.setScanFieldOfViewB(33.0)
        b.setCTDIvol(level)
    elif kind == "dx":
        b = dx.DXBeam(filtration={13: 2.0, 29: 0.1})
        b.setRotationCenter([0, 0, 1])
        b.setSourcePatientDistance(60.0)
        b.setSourceDetectorDistance(110.0)
        b.setCollimation([20.0, 12.0])
        b.setPrimaryAngleDeg(25.0)
        b.setSecondaryAngleDeg(-10.0)
        b.setTubeVoltage(80.0)
        b.setDAPvalue(level)
        b.setNumberOfExposures(12)
    elif kind == "cbct":
        b = dx.CBCTBeam((0, 0, 0.5), (0, 0, 1), {13: 2.5})
        b.setSourceDetectorDistance(90.0)
        b.setStartAngleDeg(10.0)
        b.setStopAngleDeg(250.0)
        b.setStepAngleDeg(12.0)
        b.setCollimationHalfAnglesDeg(8.0, 5.0)
        b.setDAPvalue(level)
    else:
        b = dx.PencilBeam()
        b.setPosition([0.3, -30.0, 0.2])
        b.setDirection([0, 1, 0])
        b.setEnergy(70.0)
        b.setAirKerma(level)
        b.setNumberOfExposures(10)
    b.setNumberOfParticlesPerExposure(per_exposure)
    return b


@pytest.mark.parametrize("kind,mode,delete_air,level", [
    ("sequential", 1, 1, 1.0), ("sequential", 1, 0, 1.0), ("sequential", 0, 1, 0.004), ("sequential", 2, 1, 1.0),
    ("spiral", 1, 1, 5.0), ("dual", 1, 1, 5.0), ("dx", 1, 0, 1.0), ("cbct", 2, 1, 200.0), ("pencil", 1, 1, 1.0)])
def test_reference_simulation_pipeline_end_to_end_on_the_cpu_double(dx, orc, ref_rows, tmp_path, kind, mode, delete_air, level):
    """OpenDXMC's own SimulationPipeline - worker<CORRECTION>() with its World / AAVoxelGrid / Material / Transport calls and
    its post-processing (air mask, uGy rule; R:src/libopendxmc/simulationpipeline.cpp:124-235) - compiled unmodified, runs
    end to end on each of the six beam types (bowtie, WED AEC, organ AEC, dual source, DAP / air-kerma / CTDI calibrations)
    over the reference's own PMMA cylinder.  oracle/_ref/opendxmc_ref_cpu links a CPU test double of the nine
    context-level dxb_* calls (oracle/cpu_double.cpp -> the oracle) ahead of the library, so no GPU is needed; the
    Python mirror + oracle + orc_postprocess on the same inputs must give the same result, which pins the C++ shims'
    plumbing (materials, grid, every beam descriptor) and the post-processing the GPU path is tested against."""
    from opendxmc_b200 import _capi as K
    prefix = str(tmp_path / "ref")
    env = dict(os.environ, DXB_DOUBLE_CALIB="360000")
    r = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "opendxmc_ref_cpu"), "run", str(mode), str(delete_air), "1200", prefix, repr(level), kind],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, (r.stdout, r.stderr)
    meta = json.load(open(prefix + ".json"))
    n = int(np.prod(meta["dim"]))
    dens = np.fromfile(prefix + ".density.bin", dtype=np.float64)
    mat = np.fromfile(prefix + ".material.bin", dtype=np.uint8)
    ref = [np.fromfile(prefix + f".{k}.bin", dtype=np.float64) for k in ("dose", "variance", "count")]
    assert dens.size == mat.size == n
    names = ("Air, Dry (near sea level)", "Polymethyl Methacralate (Lucite, Perspex)")
    mats = [dx.Material.byWeight(dx.NISTMaterials.Composition(nm)) for nm in names]
    ow = orc.OracleWorld(meta["dim"], meta["spacing"], dens, mat, mats)
    wed = dx.workloads.wed_aec_profile(dens, meta["dim"], meta["spacing"])
    beam = _run_beam(dx, kind, 1200, level, wed, meta["spacing"][2] * meta["dim"][2] / 2.0)
    assert beam.numberOfExposures() == meta["exposures"]
    d, v, c = ow.transport(beam, mode, True, 0x0DDC0FFEE, 360000)[:3]
    dd, vv, cc = d.copy(), v.copy(), c.astype(np.float64)
    micro = orc.load().orc_postprocess(dd.ctypes.data_as(K.c_double_p), vv.ctypes.data_as(K.c_double_p), cc.ctypes.data_as(K.c_double_p),
                                       mat.ctypes.data_as(K.c_u8_p), n, delete_air)
    assert meta["dose_units"] == ("uGy" if micro else "mGy")
    if kind == "sequential":
        assert (meta["dose_units"] == "uGy") == (level < 0.1)          # the uGy rule fires for the weak beam
    assert np.array_equal(cc, ref[2]) and cc.sum() > 2000
    # thread order moves last bits only; the variance is a difference of sums, so its smallest values carry cancellation noise
    assert np.allclose(dd, ref[0], rtol=1e-12, atol=0)
    assert np.allclose(vv, ref[1], rtol=1e-9, atol=1e-13 * float(vv.max()))
    air = mat == 0
    assert (ref[0][air].sum() == 0) == bool(delete_air) and ref[0][~air].sum() > 0


def test_second_start_simulation_on_the_same_pipeline(tmp_path):
    """The reference's worker raises the stop flag of the pipeline's single TransportProgress at the end of every run
    (R:src/libopendxmc/simulationpipeline.cpp:234) and nothing in OpenDXMC clears it: dxmc::Transport::operator() must
    start the progress object itself, or the second startSimulation() is 'cancelled' before it begins and the 3 s timer
    re-publishes the stale first result."""
    prefix = str(tmp_path / "twice")
    env = dict(os.environ, DXB_DOUBLE_CALIB="180000")
    r = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "opendxmc_ref_cpu"), "run", "1", "1", "600", prefix, "1.0", "sequential", "1"],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, (r.stdout, r.stderr)
    meta = json.load(open(prefix + ".json"))
    assert meta["second_run_identical"] == 1


def test_world_shim_rejects_a_malformed_options_variable(tmp_path):
    """DXMC_B200_OPTIONS="key=value,..." lets an application that cannot be recompiled reach dxb_set_option (e.g. the reference's
    tracking rule: dense_box=0,local_majorant=0); a malformed entry must stop the run loudly, not be ignored."""
    env = dict(os.environ, DXMC_B200_OPTIONS="nonsense")
    r = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "opendxmc_ref_cpu"), "run", "1", "1", "600", str(tmp_path / "opt"), "1.0", "sequential"],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode != 0
    assert "DXMC_B200_OPTIONS: expected key=value" in r.stderr

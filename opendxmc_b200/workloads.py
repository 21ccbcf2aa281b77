"""Synthetic inputs for the BASELINE.json configurations.

Everything here produces *inputs* in the layouts the reference feeds to AAVoxelGrid::setData
(R:src/libopendxmc/simulationpipeline.cpp:145-150): f64 density [g/cm3], u8 material index, x fastest
(R:src/libopendxmc/otherphantomimportpipeline.cpp:44), spacing in cm, plus a beam object.  The
geometry is analytic and deterministic (no RNG); sizes can be scaled down for parity tests.
"""
import json
import math
import os

import numpy as np

from . import api

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")
DEFAULT_BOWTIE = "Siemens Definition Flash W1 120kV"  # R:src/libopendxmc/beamsettingsmodel.cpp:1303-1311


def read_bowtie_filters(path=None):
    """BowtieFilterReader::read — R:src/libopendxmc/bowtiefilterreader.cpp:51-97 (same JSON schema)."""
    doc = json.load(open(path or os.path.join(_DATA, "bowtiefilters.json")))
    res = {}
    if not isinstance(doc, dict) or not isinstance(doc.get("filters"), list):
        return res
    for f in doc["filters"]:
        if not isinstance(f, dict):
            continue
        name = f.get("name", "")
        data = [(d["angle"], d["weight"]) for d in f.get("filterdata", [])
                if isinstance(d, dict) and isinstance(d.get("angle"), (int, float)) and isinstance(d.get("weight"), (int, float))]
        if data and name:
            res[name] = api.BowtieFilter(data)
    return res


def icrp_tables():
    return json.load(open(os.path.join(_DATA, "icrp_tables.json")))


def icrp_shapes():
    return json.load(open(os.path.join(_DATA, "icrp_shapes.json")))


class Workload:
    def __init__(self, name, dim, spacing, density, material, materials, material_names, beam, organ=None, organ_names=None):
        self.name = name
        self.dim = [int(v) for v in dim]
        self.spacing = [float(v) for v in spacing]
        self.density = density
        self.material = material
        self.materials = materials
        self.material_names = material_names
        self.beam = beam
        self.organ = organ
        self.organ_names = organ_names

    @property
    def n_voxels(self):
        return self.dim[0] * self.dim[1] * self.dim[2]

    def build_world(self, correction=1, devices=None):
        world = api.World(devices)
        grid = world.addItem(api.AAVoxelGrid(correction))
        if not grid.setData(self.dim, self.density, self.material, self.materials):
            raise ValueError("setData rejected the workload arrays")
        grid.setSpacing(self.spacing)
        world.build()
        return world


def _coords(n, d):
    # voxel-centre coordinates of a grid centred on the origin (R:src/libopendxmc/datacontainer.cpp:174-178)
    return (np.arange(n, dtype=np.float32) + 0.5) * np.float32(d) - np.float32(0.5 * n * d)


def _hash_jitter(nx, ny, nz, amplitude):
    """deterministic +-amplitude multiplicative jitter from a hash of the flat voxel index (no RNG)."""
    out = np.empty((nz, ny, nx), dtype=np.float32)
    plane = (np.arange(ny, dtype=np.uint64)[:, None] * np.uint64(nx) + np.arange(nx, dtype=np.uint64)[None, :])
    for k in range(nz):
        h = plane + np.uint64(k) * np.uint64(nx * ny)
        h = (h ^ (h >> np.uint64(15))) * np.uint64(0x2C1B3C6D) & np.uint64(0xFFFFFFFF)
        h = (h ^ (h >> np.uint64(12))) * np.uint64(0x297A2D39) & np.uint64(0xFFFFFFFF)
        h = h ^ (h >> np.uint64(15))
        out[k] = (h & np.uint64(0xFFFF)).astype(np.float32) * np.float32(2.0 / 65535.0) - np.float32(1.0)
    return np.float32(1.0) + np.float32(amplitude) * out


def ctdi_body_phantom(n=64, histories=10_000_000, step_deg=1.0):
    """C1: CTDI 32 cm PMMA body phantom on an n^3 grid (36 cm cube), 120 kV axial beam, 1e7 histories."""
    side = 36.0
    d = side / n
    x = _coords(n, d)
    r2 = x[None, :] ** 2 + x[:, None] ** 2
    inside = (r2 <= 16.0 ** 2)
    mat2d = inside.astype(np.uint8)
    material = np.broadcast_to(mat2d[None, :, :], (n, n, n)).copy().reshape(-1)
    names = ["Air, Dry (near sea level)", "Polymethyl Methacralate (Lucite, Perspex)"]
    rho = [api.NISTMaterials.density(nm) for nm in names]
    density = np.where(material == 1, rho[1], rho[0]).astype(np.float64)
    mats = [api.Material.byWeight(api.NISTMaterials.Composition(nm)) for nm in names]

    beam = api.CTSequentialBeam([0, 0, 0], [0, 0, 1], {13: 9.0})
    beam.setTubeVoltage(120.0)
    beam.setSourceDetectorDistance(119.0)
    beam.setCollimation(3.84)
    beam.setScanFieldOfView(50.0)
    beam.setNumberOfSlices(1)
    beam.setStepAngleDeg(step_deg)
    beam.setBowtieFilter(read_bowtie_filters()[DEFAULT_BOWTIE])
    nexp = beam.numberOfExposures()
    beam.setNumberOfParticlesPerExposure(max(1, int(math.ceil(histories / nexp))))
    return Workload(f"C1 CTDI body phantom {n}^3 axial 120kV", [n, n, n], [d, d, d], density, material, mats, names, beam,
                    organ=material.copy(), organ_names=names)


def _patient_volume(nx, ny, nz, dx, dy, dz, heart=False):
    """air / lung / soft tissue / cortical bone anatomy of configs C2 and C4."""
    x = _coords(nx, dx)[None, None, :]
    y = _coords(ny, dy)[None, :, None]
    z = _coords(nz, dz)[:, None, None]
    zh = 0.5 * nz * dz
    material = np.zeros((nz, ny, nx), dtype=np.uint8)
    density = np.full((nz, ny, nx), 0.0012, dtype=np.float32)
    body = (x / 17.0) ** 2 + (y / 12.0) ** 2 <= 1.0
    body = np.broadcast_to(body, material.shape)
    material[body] = 2
    density[body] = 1.03
    # two lung ellipsoids
    for cx in (-7.5, 7.5):
        lung = ((x - cx) / 5.5) ** 2 + ((y + 0.5) / 7.5) ** 2 + (z / (0.85 * zh)) ** 2 <= 1.0
        material[lung] = 1
        density[lung] = 0.26
    if heart:
        h = ((x + 1.5) / 4.5) ** 2 + ((y + 2.0) / 4.0) ** 2 + (z / (0.35 * zh)) ** 2 <= 1.0
        material[h] = 2
        density[h] = 1.05
    # rib shell: elliptical annulus, present in 1.2 cm bands every 2.4 cm along z
    rr = (x / 15.0) ** 2 + (y / 10.0) ** 2
    ro = (x / 16.0) ** 2 + (y / 11.0) ** 2
    band = (np.mod(z + 100.0, 2.4) < 1.2)
    ribs = (rr >= 1.0) & (ro <= 1.0) & band
    material[ribs] = 3
    density[ribs] = 1.6
    # spine: cylinder of 4 cm diameter behind the lungs
    spine = (x ** 2 + (y - 7.5) ** 2 <= 4.0)
    spine = np.broadcast_to(spine, material.shape)
    material[spine] = 3
    density[spine] = 1.6
    density *= _hash_jitter(nx, ny, nz, 0.02)
    return density.astype(np.float64).reshape(-1), material.reshape(-1)


def _patient_materials():
    names = ["Air, Dry (near sea level)", "Lung (soft tissue composition)", "Tissue, Soft (ICRP)", "Bone, Cortical (ICRP)"]
    comp = [api.NISTMaterials.Composition("Air, Dry (near sea level)"), api.NISTMaterials.Composition("Tissue, Soft (ICRP)"),
            api.NISTMaterials.Composition("Tissue, Soft (ICRP)"), api.NISTMaterials.Composition("Bone, Cortical (ICRP)")]
    return names, [api.Material.byWeight(c) for c in comp]


def ct_spiral_patient(scale=1, histories=1_000_000_000, step_deg=1.0):
    """C2: 512x512x300 synthetic CT patient, 0.08x0.08x0.1 cm, 120 kV spiral, pitch 1, default bowtie.
    scale > 1 coarsens the grid by that factor (same physical extent) for parity tests."""
    nx, ny, nz = 512 // scale, 512 // scale, 300 // scale
    dx, dy, dz = 0.08 * scale, 0.08 * scale, 0.1 * scale
    density, material = _patient_volume(nx, ny, nz, dx, dy, dz)
    names, mats = _patient_materials()
    zh = 0.5 * nz * dz
    beam = api.CTSpiralBeam([0, 0, -zh], [0, 0, zh], {13: 9.0})
    beam.setTubeVoltage(120.0)
    beam.setSourceDetectorDistance(119.0)
    beam.setCollimation(3.84)
    beam.setPitch(1.0)
    beam.setScanFieldOfView(50.0)
    beam.setStartAngleDeg(0.0)
    beam.setStepAngleDeg(step_deg)
    beam.setBowtieFilter(read_bowtie_filters()[DEFAULT_BOWTIE])
    nexp = beam.numberOfExposures()
    beam.setNumberOfParticlesPerExposure(max(1, int(math.ceil(histories / nexp))))
    organ = material.copy()
    return Workload(f"C2 CT patient {nx}x{ny}x{nz} spiral 120kV bowtie", [nx, ny, nz], [dx, dy, dz], density, material, mats, names,
                    beam, organ=organ, organ_names=["air", "lung", "soft tissue", "bone"])


def wed_aec_profile(density, dim, spacing):
    """DataContainer::calculateAECfilterFromWaterEquivalentDiameter — R:src/libopendxmc/datacontainer.cpp:42-100:
    per-slice water-equivalent diameter D_w = 2 sqrt(sum(rho) dx dy / pi), weight exp(0.2 D_w) (then the filter normalises)."""
    nx, ny, nz = dim
    d = density.reshape(nz, ny, nx)
    area = d.sum(axis=(1, 2)) * spacing[0] * spacing[1]
    dw = 2.0 * np.sqrt(area / math.pi)
    return np.exp(0.2 * dw)


def ct_dual_source_thorax(scale=1, histories=10_000_000_000, step_deg=1.0):
    """C4: 512x512x400 thorax, dual-source spiral (Flash): B offset 95 deg, FOV 50/33, pitch 3.2, WED AEC."""
    nx, ny, nz = 512 // scale, 512 // scale, 400 // scale
    dx, dy, dz = 0.08 * scale, 0.08 * scale, 0.1 * scale
    density, material = _patient_volume(nx, ny, nz, dx, dy, dz, heart=True)
    names, mats = _patient_materials()
    zh = 0.5 * nz * dz
    beam = api.CTSpiralDualEnergyBeam([0, 0, -zh], [0, 0, zh], {13: 9.0})
    beam.setTubeAVoltage(120.0)
    beam.setTubeBVoltage(120.0)
    beam.setTubeBoffsetAngleDeg(95.0)
    beam.setScanFieldOfViewA(50.0)
    beam.setScanFieldOfViewB(33.0)
    beam.setSourceDetectorDistance(119.0)
    beam.setCollimation(3.84)
    beam.setPitch(3.2)
    beam.setStepAngleDeg(step_deg)
    bt = read_bowtie_filters()[DEFAULT_BOWTIE]
    beam.setBowtieFilterA(bt)
    beam.setBowtieFilterB(bt)
    w = wed_aec_profile(density, (nx, ny, nz), (dx, dy, dz))
    beam.setAECFilter([0, 0, -zh], [0, 0, zh], w)
    nexp = beam.numberOfExposures()
    beam.setNumberOfParticlesPerExposure(max(1, int(math.ceil(histories / nexp))))
    return Workload(f"C4 thorax {nx}x{ny}x{nz} dual-source spiral AEC", [nx, ny, nz], [dx, dy, dz], density, material, mats, names,
                    beam, organ=material.copy(), organ_names=["air", "lung", "soft tissue", "bone"])


def icrp_dat_text(phantom):
    """The phantom's `<phantom>_organs.dat` / `<phantom>_media.dat` tables as TEXT in the layout of the ICRP files (id at
    column 0, name from column 6, tissue number and density behind a name field at least 50 characters wide; media:
    id, name, 13 mass-% columns for Z = 1,6,7,8,11,12,15,16,17,19,20,26,53), regenerated from the packaged
    icrp_tables.json - the input of dxb_icrp_import / dxb_icrp_plan (include/dxb.h).  The reference's own parser reads
    this text to the same tables as the original files (tests/test_reference_sources_compile.py)."""
    t = icrp_tables()[phantom]
    organs = ["Organs and tissues of the %s reference computational phantom" % phantom, "", "Organ Organ" + " " * 42 + "Tissue Density",
              "ID" + " " * 51 + "number"]
    for o in t["organs"]:
        d = "%.3f" % o["density"]
        if float(d) != o["density"]:
            d = repr(float(o["density"]))
        organs.append("%-6d%-49s%2d%9s" % (o["id"], o["name"], o["medium"], d))
    zs = [1, 6, 7, 8, 11, 12, 15, 16, 17, 19, 20, 26, 53]
    media = [" " * 80 + "".join("%6d" % z for z in zs), "No." + " " * 100 + "(% by mass)"]
    for m in t["media"]:
        cols = []
        for z in zs:
            w = m["composition"][str(z)]
            c = "%.1f" % w
            cols.append(c if float(c) == w else repr(float(w)))
        media.append("%-6d%-72s%s" % (m["id"], m["name"], "".join("%6s" % c for c in cols)))
    return "\n".join(organs) + "\n", "\n".join(media) + "\n"


def icrp_import(organ_array, organs_text, media_text, remove_arms=False, world=None):
    """ICRPPhantomImportPipeline::importPhantom through the C ABI (R:src/libopendxmc/icrpphantomimportpipeline.cpp:258-351).
    With a `world` (a built api.World: any context with a device) the O(N) passes run on the GPU (dxb_icrp_import);
    without one the host-side rules alone are evaluated (dxb_icrp_plan) and the three look-up tables applied with numpy.
    Returns (organ, organ_names, material, density, media_names, compositions) like import_icrp_tables."""
    import ctypes as C
    from . import _capi as K
    lib = K.load()
    organ_array = np.ascontiguousarray(organ_array, dtype=np.uint8).reshape(-1)
    n = organ_array.size
    plan = K.VP()
    if world is not None:
        organ, material, density = np.zeros(n, dtype=np.uint8), np.zeros(n, dtype=np.uint8), np.zeros(n)
        rc = lib.dxb_icrp_import(world.ctx(), organ_array.ctypes.data_as(K.c_u8_p), n, organs_text.encode(), media_text.encode(),
                                 1 if remove_arms else 0, organ.ctypes.data_as(K.c_u8_p), material.ctypes.data_as(K.c_u8_p),
                                 density.ctypes.data_as(K.c_double_p), C.byref(plan))
        if rc != K.DXB_OK:
            raise K.DxbError(rc, "dxb_icrp_import", (lib.dxb_last_error(world.ctx()) or b"").decode())
    else:
        present = np.zeros(256, dtype=np.uint8)
        present[np.unique(organ_array)] = 1
        rc = lib.dxb_icrp_plan(C.byref(plan), organs_text.encode(), media_text.encode(), 1 if remove_arms else 0, present.ctypes.data_as(K.c_u8_p))
        if rc != K.DXB_OK:
            raise K.DxbError(rc, "dxb_icrp_plan")
        lo, lm, ld = np.zeros(256, dtype=np.uint8), np.zeros(256, dtype=np.uint8), np.zeros(256)
        lib.dxb_icrp_luts(plan, lo.ctypes.data_as(K.c_u8_p), lm.ctypes.data_as(K.c_u8_p), ld.ctypes.data_as(K.c_double_p))
        organ, material, density = lo[organ_array], lm[organ_array], ld[organ_array]
    try:
        names = [lib.dxb_icrp_organ_name(plan, i).decode() for i in range(lib.dxb_icrp_n_organs(plan))]
        media_names, comps = [], []
        for i in range(lib.dxb_icrp_n_media(plan)):
            media_names.append(lib.dxb_icrp_medium_name(plan, i).decode())
            z, w = np.zeros(16, dtype=np.uint32), np.zeros(16)
            k = lib.dxb_icrp_medium_composition(plan, i, z.ctypes.data_as(K.c_u32_p), w.ctypes.data_as(K.c_double_p), 16)
            comps.append({int(z[j]): float(w[j]) for j in range(k)})
    finally:
        lib.dxb_icrp_destroy(plan)
    return organ, names, material, density, media_names, comps


def import_icrp_tables(phantom, organ_array):
    """ICRPPhantomImportPipeline::importPhantom remap rules — R:src/libopendxmc/icrpphantomimportpipeline.cpp:209-351:
    air appended as organ 0 / medium 0 (rho 0.001, {N:0.8,O:0.2}); organs absent from the array pruned and ids made
    consecutive; media not referenced pruned and ids made consecutive; material/density arrays by organ lookup."""
    t = icrp_tables()[phantom]
    organs = [dict(o) for o in t["organs"]] + [{"id": 0, "name": "Air", "medium": 0, "density": 0.001}]
    media = [dict(m) for m in t["media"]] + [{"id": 0, "name": "Air", "composition": {"7": 0.8, "8": 0.2}}]
    organ_array = np.ascontiguousarray(organ_array, dtype=np.uint8).reshape(-1)
    present = np.zeros(256, dtype=bool)
    present[np.unique(organ_array)] = True
    organs = sorted([o for o in organs if present[o["id"]]], key=lambda o: o["id"])
    lut = np.zeros(256, dtype=np.uint8)
    for i, o in enumerate(organs):
        lut[o["id"]] = i
        o["id"] = i
    organ_new = lut[organ_array]
    used = {o["medium"] for o in organs}
    media = sorted([m for m in media if m["id"] in used], key=lambda m: m["id"])
    mlut = {m["id"]: i for i, m in enumerate(media)}
    for o in organs:
        o["medium"] = mlut[o["medium"]]
    o2m = np.array([o["medium"] for o in organs], dtype=np.uint8)
    o2d = np.array([o["density"] for o in organs], dtype=np.float64)
    material = o2m[organ_new]
    density = o2d[organ_new]
    comps = [{int(z): w for z, w in m["composition"].items() if w > 0} for m in media]
    return organ_new, [o["name"] for o in organs], material, density, [m["name"] for m in media], comps


def _nested_ellipsoid_organs(nx, ny, nz, dx, dy, dz, organ_ids):
    """synthetic organ map of an ICRP phantom's shape: a body ellipsoid-cylinder filled with nested ellipsoidal
    organs laid out along z by id (the real voxel arrays are missing: R:.MISSING_LARGE_BLOBS)."""
    x = _coords(nx, dx)[None, None, :]
    y = _coords(ny, dy)[None, :, None]
    z = _coords(nz, dz)[:, None, None]
    hx, hy, hz = 0.5 * nx * dx, 0.5 * ny * dy, 0.5 * nz * dz
    organ = np.zeros((nz, ny, nx), dtype=np.uint8)
    body = np.broadcast_to((x / (0.85 * hx)) ** 2 + (y / (0.8 * hy)) ** 2 <= 1.0, organ.shape) & np.broadcast_to(np.abs(z) <= 0.97 * hz, organ.shape)
    ids = list(organ_ids)
    background = ids[0]
    organ[body] = background
    rest = ids[1:]
    n = len(rest)
    for k, oid in enumerate(rest):
        # centres spiral down the body; sizes cycle so that small and large organs both exist
        zc = -0.9 * hz + 1.8 * hz * (k + 0.5) / n
        ang = 2.399963 * k
        rad = 0.45 * (0.3 + 0.7 * ((k * 7) % 10) / 10.0)
        xc, yc = rad * hx * math.cos(ang), rad * hy * math.sin(ang)
        ax = hx * (0.10 + 0.12 * ((k * 3) % 5) / 5.0)
        ay = hy * (0.12 + 0.14 * ((k * 5) % 7) / 7.0)
        az = max(hz * 3.0 / n, 1.5 * dz)
        e = ((x - xc) / ax) ** 2 + ((y - yc) / ay) ** 2 + ((z - zc) / az) ** 2 <= 1.0
        organ[e & body] = oid
    return organ.reshape(-1)


def icrp_phantom(phantom="AM", scale=1, histories=100_000_000, beam_kind="ct_chest"):
    """C3 (AM, chest CT, per-organ dose) and C5 (10M child, 80 kV DX) on synthetic organ maps of the real shapes,
    with the real organ -> medium -> density tables."""
    shp = icrp_shapes()[phantom]
    nx, ny, nz = [max(4, v // scale) for v in shp["dimensions"]]
    dx, dy, dz = [s * 0.1 * scale for s in shp["spacing_mm"]]  # mm -> cm (R:src/libopendxmc/datacontainer.cpp:247-253)
    t = icrp_tables()[phantom]
    ids = [o["id"] for o in t["organs"]]
    # use the residual/soft-tissue-like organ as body background if present
    bg = next((o["id"] for o in t["organs"] if "Residual" in o["name"] or "Muscle" in o["name"]), ids[0])
    ids = [bg] + [i for i in ids if i != bg]
    organ_raw = _nested_ellipsoid_organs(nx, ny, nz, dx, dy, dz, ids)
    organ, organ_names, material, density, media_names, comps = import_icrp_tables(phantom, organ_raw)
    mats = [api.Material.byWeight(c) for c in comps]
    if any(m is None for m in mats):
        raise ValueError("Material.byWeight failed for an ICRP medium")
    zh = 0.5 * nz * dz
    if beam_kind == "ct_chest":
        half = min(17.5, zh)
        zc = 0.25 * zh
        beam = api.CTSpiralBeam([0, 0, zc - half], [0, 0, zc + half], {13: 9.0})
        beam.setTubeVoltage(120.0)
        beam.setSourceDetectorDistance(119.0)
        beam.setCollimation(3.84)
        beam.setPitch(1.0)
        beam.setScanFieldOfView(50.0)
        beam.setStepAngleDeg(1.0)
        beam.setBowtieFilter(read_bowtie_filters()[DEFAULT_BOWTIE])
        nexp = beam.numberOfExposures()
        beam.setNumberOfParticlesPerExposure(max(1, int(math.ceil(histories / nexp))))
    else:
        # AP chest radiograph, 80 kV, 2 mm Al + 0.1 mm Cu, SPD 100, SDD 100, 35 x 43 cm field
        beam = api.DXBeam(filtration={13: 2.0, 29: 0.1})
        beam.setTubeVoltage(80.0)
        beam.setSourceDetectorDistance(100.0)
        beam.setRotationCenter([0.0, 0.0, 0.25 * zh])
        beam.setSourcePatientDistance(100.0)
        beam.setPrimaryAngleDeg(0.0)
        beam.setSecondaryAngleDeg(0.0)
        beam.setCollimation([35.0, 43.0])
        beam.setDAPvalue(1.0)
        nexp = 64
        beam.setNumberOfExposures(nexp)
        beam.setNumberOfParticlesPerExposure(max(1, int(math.ceil(histories / nexp))))
    return Workload(f"ICRP {phantom} shape {nx}x{ny}x{nz} {beam_kind}", [nx, ny, nz], [dx, dy, dz], density, material, mats,
                    media_names, beam, organ=organ, organ_names=organ_names)


CONFIGS = {
    "C1": lambda **kw: ctdi_body_phantom(**kw),
    "C2": lambda **kw: ct_spiral_patient(**kw),
    "C3": lambda **kw: icrp_phantom("AM", beam_kind="ct_chest", **kw),
    "C4": lambda **kw: ct_dual_source_thorax(**kw),
    "C5": lambda **kw: icrp_phantom("10M", beam_kind="dx", **{"histories": 1_000_000_000, **kw}),
}

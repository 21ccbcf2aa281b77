"""ctypes view of oracle/liboracle.so — TEST INFRASTRUCTURE (see oracle/oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
PARITY UNPINNED: the oracle restates the recalled DXMClib algorithm; it is pinned only against analytic
known answers and the Random123 Philox vectors (tests/test_oracle_*.py).
"""
import ctypes as C
import os
import subprocess

import numpy as np

from opendxmc_b200 import _capi as K

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")


class orc_stats(C.Structure):
    _fields_ = [
        ("histories", C.c_uint64), ("steps", C.c_uint64), ("interactions", C.c_uint64), ("deposits", C.c_uint64),
        ("energy_emitted_kev", C.c_double), ("energy_deposited_kev", C.c_double), ("calibration_factor", C.c_double),
        ("seconds", C.c_double), ("threads", C.c_int), ("hops", C.c_uint64),
    ]


_lib = None
VP = C.c_void_p
BD = C.POINTER(K.dxb_beam_desc)
MT = C.POINTER(K.dxb_material_tables)


def build():
    subprocess.check_call(["make", "-s", "oracle"], cwd=os.path.dirname(_HERE))


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        build()
    lib = C.CDLL(LIB_PATH)
    sig = {
        "orc_world_create": (VP, [K.c_u64_p, K.c_double_p, K.c_double_p, K.c_u8_p, C.c_uint32, MT]),
        "orc_world_destroy": (None, [VP]),
        "orc_world_set_local_majorant": (None, [VP, C.c_int, C.c_int, K.c_float_p]),
        "orc_world_build_local_majorant": (C.c_int, [VP, C.c_int]),
        "orc_world_set_dense_box": (None, [VP, K.c_float_p, K.c_float_p]),
        "orc_world_build_dense_box": (C.c_int, [VP, C.c_double]),
        "orc_set_device_mirroring": (None, [C.c_int]),
        "orc_get_device_mirroring": (C.c_int, []),
        "orc_world_set_reference_materials": (None, [VP, MT, MT, C.c_double, C.c_double]),
        "orc_run": (C.c_int, [VP, BD, C.c_int, C.c_uint64, C.c_int, C.c_uint64, C.c_uint64, K.c_double_p, K.c_double_p, K.c_u64_p,
                              C.POINTER(orc_stats)]),
        "orc_transport": (C.c_int, [VP, BD, C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_int, K.c_double_p, K.c_double_p, K.c_u64_p,
                                    C.POINTER(orc_stats)]),
        "orc_ct_calibration": (C.c_double, [VP, BD, C.c_int, C.c_uint64, C.c_uint64, C.c_int]),
        "orc_philox4x32_10": (None, [K.c_u32_p, K.c_u32_p, K.c_u32_p]),
        "orc_beam_number_of_exposures": (C.c_uint64, [BD]),
        "orc_beam_exposure": (C.c_int, [BD, C.c_uint64, C.POINTER(K.dxb_exposure)]),
        "orc_bowtie_weight": (C.c_double, [C.POINTER(K.dxb_bowtie), C.c_double]),
        "orc_aec_weight": (C.c_double, [C.POINTER(K.dxb_aec), K.c_double_p]),
        "orc_organ_aec_weight": (C.c_double, [C.POINTER(K.dxb_organ_aec), C.c_double]),
        "orc_attenuation": (None, [MT, C.c_double, K.c_double_p]),
        "orc_majorant": (C.c_double, [VP, C.c_double]),
        "orc_sample_compton": (None, [MT, C.c_int, C.c_double, C.c_uint64, C.c_uint64, K.c_double_p, K.c_double_p]),
        "orc_sample_rayleigh": (None, [MT, C.c_int, C.c_double, C.c_uint64, C.c_uint64, K.c_double_p]),
        "orc_sample_source": (None, [BD, C.c_uint64, C.c_uint64, C.c_uint64, K.c_double_p, K.c_double_p, K.c_double_p, K.c_double_p]),
        "orc_organ_dose": (None, [K.c_double_p, K.c_double_p, K.c_u8_p, C.c_uint64, C.c_double, C.c_uint32, K.c_double_p, K.c_double_p,
                                  K.c_u64_p]),
        "orc_postprocess": (C.c_int, [K.c_double_p, K.c_double_p, K.c_double_p, K.c_u8_p, C.c_uint64, C.c_int]),
        "orc_segment": (None, [K.c_double_p, C.c_uint64, K.c_double_p, C.c_int, K.c_double_p, C.c_double, C.c_double, K.c_u8_p,
                               K.c_double_p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _dp(a):
    return a.ctypes.data_as(K.c_double_p)


class unmirrored:
    """context manager: oracle worlds created inside run on the caller's f64 data as handed over, WITHOUT the
    quantisations that mirror the device's storage formats (oracle.h: orc_set_device_mirroring)."""

    def __enter__(self):
        load().orc_set_device_mirroring(0)
        return self

    def __exit__(self, *a):
        load().orc_set_device_mirroring(1)
        return False


def philox(key, ctr):
    k = np.array(key, dtype=np.uint32)
    c = np.array(ctr, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    load().orc_philox4x32_10(k.ctypes.data_as(K.c_u32_p), c.ctypes.data_as(K.c_u32_p), out.ctypes.data_as(K.c_u32_p))
    return out


class OracleWorld:
    """World<AAVoxelGrid> of the oracle, fed with the product's master (f64) tables so both sides run on identical data."""

    def __init__(self, dim, spacing, density, material, materials):
        import opendxmc_b200 as dx
        lib = load()
        self._materials = list(materials)  # keep table owners alive
        tabs = (K.dxb_material_tables * len(materials))(*[m.tables() for m in materials])
        dimc = (C.c_uint64 * 3)(*[int(v) for v in dim])
        spc = (C.c_double * 3)(*[float(v) for v in spacing])
        density = np.ascontiguousarray(density, dtype=np.float64).reshape(-1)
        material = np.ascontiguousarray(material, dtype=np.uint8).reshape(-1)
        self.n = density.size
        self._h = lib.orc_world_create(dimc, spc, _dp(density), material.ctypes.data_as(K.c_u8_p), len(materials), tabs)
        self._air = dx.Material.byNistName("Air, Dry (near sea level)")
        self._pmma = dx.Material.byNistName("Polymethyl Methacralate (Lucite, Perspex)")
        ta, tp = self._air.tables(), self._pmma.tables()
        lib.orc_world_set_reference_materials(self._h, C.byref(ta), C.byref(tp), dx.NISTMaterials.density("Air, Dry (near sea level)"),
                                              dx.NISTMaterials.density("Polymethyl Methacralate (Lucite, Perspex)"))

    @classmethod
    def from_workload(cls, wl):
        return cls(wl.dim, wl.spacing, wl.density, wl.material, wl.materials)

    def __del__(self):
        try:
            load().orc_world_destroy(self._h)
        except Exception:
            pass

    def build_local_majorant(self, shift):
        """slab-local majorants from the oracle's own f64 table: slabs of 2**shift voxel layers; returns the slab count"""
        return int(load().orc_world_build_local_majorant(self._h, int(shift)))

    def build_dense_box(self, theta=0.02):
        """dense-box tracking from the oracle's own f64 tables; True if the box is a proper part of the grid"""
        return bool(load().orc_world_build_dense_box(self._h, float(theta)))

    def set_dense_box(self, faces, ratio):
        """track with the box the device built (World.dense_box()); faces None switches it off"""
        if faces is None:
            load().orc_world_set_dense_box(self._h, None, None)
            return
        f = np.ascontiguousarray(faces, dtype=np.float32)
        r = np.ascontiguousarray(ratio, dtype=np.float32)
        load().orc_world_set_dense_box(self._h, f.ctypes.data_as(K.c_float_p), r.ctypes.data_as(K.c_float_p))

    def set_local_majorant(self, shift, n_slabs, ratio):
        """track with the table the device built (World.local_majorant()); n_slabs < 2 switches it off"""
        t = np.ascontiguousarray(ratio, dtype=np.float32) if ratio is not None else np.zeros(1, dtype=np.float32)
        load().orc_world_set_local_majorant(self._h, int(shift), int(n_slabs), t.ctypes.data_as(K.c_float_p))

    def run(self, beam, physics_mode=1, seed=0x0DDC0FFEE, threads=0, rank=0, world=1):
        """energy tallies of one beam: (energy[keV], energy_sq, n_events, stats dict)."""
        e, e2 = np.zeros(self.n), np.zeros(self.n)
        cnt = np.zeros(self.n, dtype=np.uint64)
        st = orc_stats()
        rc = load().orc_run(self._h, C.byref(beam.desc()), physics_mode, seed, threads, rank, world, _dp(e), _dp(e2),
                            cnt.ctypes.data_as(K.c_u64_p), C.byref(st))
        if rc != 0:
            raise RuntimeError(f"orc_run failed: {rc}")
        return e, e2, cnt, {f: getattr(st, f) for f, _ in orc_stats._fields_}

    def transport(self, beam, physics_mode=1, use_calibration=True, seed=0x0DDC0FFEE, calibration_histories=3_600_000, threads=0):
        d, v = np.zeros(self.n), np.zeros(self.n)
        cnt = np.zeros(self.n, dtype=np.uint64)
        st = orc_stats()
        rc = load().orc_transport(self._h, C.byref(beam.desc()), physics_mode, 1 if use_calibration else 0, seed, calibration_histories,
                                  threads, _dp(d), _dp(v), cnt.ctypes.data_as(K.c_u64_p), C.byref(st))
        if rc != 0:
            raise RuntimeError(f"orc_transport failed: {rc}")
        return d, v, cnt, {f: getattr(st, f) for f, _ in orc_stats._fields_}

    def ct_calibration(self, beam, physics_mode=1, seed=0x0DDC0FFEE, calibration_histories=3_600_000, threads=0):
        return load().orc_ct_calibration(self._h, C.byref(beam.desc()), physics_mode, seed, calibration_histories, threads)

    def majorant(self, e):
        return load().orc_majorant(self._h, float(e))


def attenuation(material, e):
    out = (C.c_double * 4)()
    t = material.tables()
    load().orc_attenuation(C.byref(t), float(e), out)
    return list(out)


def sample_compton(material, mode, energy, n, seed=1):
    c, r = np.zeros(n), np.zeros(n)
    t = material.tables()
    load().orc_sample_compton(C.byref(t), mode, float(energy), seed, n, _dp(c), _dp(r))
    return c, r


def sample_rayleigh(material, mode, energy, n, seed=1):
    c = np.zeros(n)
    t = material.tables()
    load().orc_sample_rayleigh(C.byref(t), mode, float(energy), seed, n, _dp(c))
    return c


def sample_source(beam, first, n, seed=0x0DDC0FFEE):
    pos, dirs, e, w = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros(n), np.zeros(n)
    load().orc_sample_source(C.byref(beam.desc()), seed, first, n, _dp(pos), _dp(dirs), _dp(e), _dp(w))
    return pos, dirs, e, w


def beam_exposure(beam, i):
    e = K.dxb_exposure()
    rc = load().orc_beam_exposure(C.byref(beam.desc()), int(i), C.byref(e))
    if rc != 0:
        raise IndexError(i)
    return e


def beam_number_of_exposures(beam):
    return int(load().orc_beam_number_of_exposures(C.byref(beam.desc())))


def organ_dose(dose, density, organ, voxel_volume, n_organs):
    dose = np.ascontiguousarray(dose, dtype=np.float64)
    density = np.ascontiguousarray(density, dtype=np.float64)
    organ = np.ascontiguousarray(organ, dtype=np.uint8)
    d, m = np.zeros(n_organs), np.zeros(n_organs)
    c = np.zeros(n_organs, dtype=np.uint64)
    load().orc_organ_dose(_dp(dose), _dp(density), organ.ctypes.data_as(K.c_u8_p), dose.size, voxel_volume, n_organs, _dp(d), _dp(m),
                          c.ctypes.data_as(K.c_u64_p))
    return d, m, c


def postprocess(dose, variance, events, material, delete_air):
    dose, variance, events = [np.array(a, dtype=np.float64) for a in (dose, variance, events)]
    material = np.ascontiguousarray(material, dtype=np.uint8)
    micro = load().orc_postprocess(_dp(dose), _dp(variance), _dp(events), material.ctypes.data_as(K.c_u8_p), dose.size, 1 if delete_air else 0)
    return dose, variance, events, ("uGy" if micro else "mGy")

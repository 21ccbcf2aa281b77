"""Host-side mirror of the DXMClib API surface that OpenDXMC consumes, in Python
over the C ABI (include/dxb.h).  Names, argument meaning and units follow the reference call sites:

    R:src/libopendxmc/simulationpipeline.cpp:124-235   World / AAVoxelGrid / Transport / doseScored
    R:src/libopendxmc/dxmc_specialization.cpp:22-90    DXBeam pose + collimation round trip
    R:src/libopendxmc/beamsettingsmodel.cpp:257-1832   every beam setter / getter
    R:src/libopendxmc/ctsegmentationpipeline.cpp:66-163 Tube, Material, NISTMaterials

The C++ shim headers under include/dxmc/ are the drop-in for the reference (C++); this module is
what tests/ and bench.py drive.  Nothing here computes physics: every number comes out of
libdxmc_b200.so, and transport requires a CUDA device (no CPU fallback).
"""
import ctypes as C
import math

import numpy as np

from . import _capi as K


def _lib():
    return K.load()


def _dp(a):
    return a.ctypes.data_as(K.c_double_p)


def _check(rc, where, ctx=None):
    if rc != K.DXB_OK:
        detail = ""
        if ctx is not None:
            detail = (_lib().dxb_last_error(ctx) or b"").decode()
        raise K.DxbError(rc, where, detail)


def DEG_TO_RAD():
    return math.pi / 180.0


def RAD_TO_DEG():
    return 180.0 / math.pi


# --------------------------------------------------------------------------- materials
class AttenuationValues:
    def __init__(self, photoelectric, incoherent, coherent):
        self.photoelectric, self.incoherent, self.coherent = photoelectric, incoherent, coherent

    def sum(self):
        return self.photoelectric + self.incoherent + self.coherent


class Material:
    """dxmc::Material<5>.  Factories return None where the reference returns std::nullopt
    (R:src/libopendxmc/simulationpipeline.cpp:136-141)."""

    def __init__(self, handle):
        self._h = handle

    def __del__(self):
        try:
            if self._h:
                _lib().dxb_material_destroy(self._h)
        except Exception:
            pass

    @staticmethod
    def byWeight(weights):
        z = np.array(list(weights.keys()), dtype=np.uint32)
        w = np.array([weights[k] for k in weights.keys()], dtype=np.float64)
        h = K.VP()
        rc = _lib().dxb_material_by_weight(C.byref(h), len(z), z.ctypes.data_as(K.c_u32_p), _dp(w))
        return Material(h) if rc == K.DXB_OK else None

    @staticmethod
    def byNistName(name):
        h = K.VP()
        rc = _lib().dxb_material_by_nist_name(C.byref(h), name.encode())
        return Material(h) if rc == K.DXB_OK else None

    @staticmethod
    def byChemicalFormula(formula):
        h = K.VP()
        rc = _lib().dxb_material_by_chemical_formula(C.byref(h), formula.encode())
        return Material(h) if rc == K.DXB_OK else None

    @staticmethod
    def fromTables(arrays, like):
        """a material made of EXTERNALLY supplied tables (include/dxb.h: dxb_material_from_tables - the drop-in route for
        EPICS-derived data).  `arrays`: dict with photo / incoh / coh / etr [n_energy] and ff_cdf / sf [n_x] on the
        library grids (see table_arrays()); shells and scalar fields are taken from the material `like`."""
        t = like.tables()
        keep = {}
        for name in ("photo", "incoh", "coh", "etr", "ff_cdf", "sf"):
            keep[name] = np.ascontiguousarray(arrays[name], dtype=np.float64)
            setattr(t, name, keep[name].ctypes.data_as(K.c_double_p))
        t.incoh_kn = t.incoh
        t.coh_thomson = t.coh
        h = K.VP()
        rc = _lib().dxb_material_from_tables(C.byref(h), C.byref(t))
        return Material(h) if rc == K.DXB_OK else None

    def attenuationValues(self, energy):
        out = (C.c_double * 3)()
        _check(_lib().dxb_material_attenuation(self._h, float(energy), out), "attenuationValues")
        return AttenuationValues(out[0], out[1], out[2])

    def massEnergyTransferAttenuation(self, energy):
        return _lib().dxb_material_mass_energy_transfer(self._h, float(energy))

    def effectiveZ(self):
        return _lib().dxb_material_effective_z(self._h)

    def formFactor(self, x):
        return _lib().dxb_material_form_factor(self._h, float(x))

    def scatterFactor(self, x):
        return _lib().dxb_material_scatter_factor(self._h, float(x))

    def tables(self):
        """dxb_material_tables (arrays owned by the material: keep `self` alive while using it)."""
        t = K.dxb_material_tables()
        _check(_lib().dxb_material_tables_get(self._h, C.byref(t)), "tables")
        return t

    def table_arrays(self):
        t = self.tables()
        ne, nx = t.n_energy, t.n_x
        g = lambda p, n: np.ctypeslib.as_array(p, shape=(n,)).copy()

        def nodes(vmin, n, per):
            i = np.arange(n)
            return vmin * np.exp2(i // per) * (1.0 + (i % per) / per)
        return {
            "energy": nodes(t.e_min_kev, ne, t.nodes_per_octave_e),
            "photo": g(t.photo, ne), "incoh": g(t.incoh, ne), "coh": g(t.coh, ne), "etr": g(t.etr, ne),
            "x": nodes(t.x_min, nx, t.nodes_per_octave_x),
            "ff_cdf": g(t.ff_cdf, nx), "sf": g(t.sf, nx),
        }


class NISTMaterials:
    @staticmethod
    def listNames():
        lib = _lib()
        return [lib.dxb_nist_name(i).decode() for i in range(lib.dxb_nist_count())]

    @staticmethod
    def density(name):
        return _lib().dxb_nist_density(name.encode())

    @staticmethod
    def Composition(name):
        z = np.zeros(32, dtype=np.uint32)
        w = np.zeros(32, dtype=np.float64)
        n = _lib().dxb_nist_composition(name.encode(), z.ctypes.data_as(K.c_u32_p), _dp(w), 32)
        return {int(z[i]): float(w[i]) for i in range(n)}


class AtomHandler:
    @staticmethod
    def toSymbol(Z):
        return _lib().dxb_atom_symbol(int(Z)).decode()


# --------------------------------------------------------------------------- tube
class Tube:
    """dxmc::Tube — R:src/libopendxmc/ctsegmentationpipeline.cpp:66-71, beamsettingsmodel.cpp:257-346."""

    def __init__(self, tubeVoltage=120.0, anodeAngleDeg=12.0, energyResolution=1.0):
        self._d = K.dxb_tube_desc()
        self._d.voltage_kv = tubeVoltage
        self._d.anode_angle_deg = anodeAngleDeg
        self._d.energy_resolution_kev = energyResolution
        self._filt = {}

    def _sync(self):
        self._d.n_filt = len(self._filt)
        for i, (z, mm) in enumerate(self._filt.items()):
            self._d.filt_Z[i] = z
            self._d.filt_mm[i] = mm
        return C.byref(self._d)

    def voltage(self):
        return self._d.voltage_kv

    def setVoltage(self, kv):
        self._d.voltage_kv = min(max(float(kv), 20.0), 150.0)

    def anodeAngleDeg(self):
        return self._d.anode_angle_deg

    def setAnodeAngleDeg(self, a):
        self._d.anode_angle_deg = float(a)

    def addFiltrationMaterial(self, Z, mm):
        if 1 <= int(Z) <= 92 and (int(Z) in self._filt or len(self._filt) < K.TUBE_MAX_FILT):
            self._filt[int(Z)] = abs(float(mm))

    def filtration(self, Z):
        return self._filt.get(int(Z), 0.0)

    def setAlFiltration(self, mm):
        self.addFiltrationMaterial(13, mm)

    def setCuFiltration(self, mm):
        self.addFiltrationMaterial(29, mm)

    def setSnFiltration(self, mm):
        self.addFiltrationMaterial(50, mm)

    def clearFiltrationMaterials(self):
        self._filt = {}

    def getEnergy(self):
        n = _lib().dxb_tube_energies(self._sync(), None, 0)
        e = np.zeros(n)
        _lib().dxb_tube_energies(self._sync(), _dp(e), n)
        return e

    def getSpecter(self, energies=None, normalize=True):
        e = self.getEnergy() if energies is None else np.ascontiguousarray(energies, dtype=np.float64)
        w = np.zeros(len(e))
        _check(_lib().dxb_tube_spectrum(self._sync(), _dp(e), len(e), 1 if normalize else 0, _dp(w)), "getSpecter")
        return w

    def mAsToHVL(self):
        return self.alHalfValueLayer()

    def alHalfValueLayer(self):
        return _lib().dxb_tube_al_half_value_layer_mm(self._sync())

    def meanSpecterEnergy(self):
        return _lib().dxb_tube_mean_energy(self._sync())


# --------------------------------------------------------------------------- filters
class BowtieFilter:
    """dxmc::BowtieFilter(vector<pair<angle_rad, weight>>) — R:src/libopendxmc/bowtiefilterreader.cpp:74-93."""

    def __init__(self, data=()):
        d = np.array(list(data), dtype=np.float64).reshape(-1, 2)
        self.angle = np.ascontiguousarray(d[:, 0])
        self.weight = np.ascontiguousarray(d[:, 1])

    def _desc(self):
        b = K.dxb_bowtie()
        b.n = len(self.angle)
        b.angle_rad = _dp(self.angle)
        b.weight = _dp(self.weight)
        return b

    def __call__(self, angle):
        b = self._desc()
        return _lib().dxb_bowtie_weight(C.byref(b), float(angle))


class CTAECFilter:
    """dxmc::CTAECFilter(start, stop, weights) — R:src/libopendxmc/datacontainer.cpp:37,59."""

    def __init__(self, start=(0, 0, 0), stop=(0, 0, 0), weights=()):
        self.setData(start, stop, weights)

    def setData(self, start, stop, weights):
        self._start = [float(v) for v in start]
        self._stop = [float(v) for v in stop]
        self._w = np.ascontiguousarray(weights, dtype=np.float64)

    def isEmpty(self):
        return len(self._w) < 2

    def weights(self):
        return self._w

    def start(self):
        return list(self._start)

    def stop(self):
        return list(self._stop)

    def size(self):
        return len(self._w)

    def length(self):
        return math.dist(self._start, self._stop)

    def _desc(self):
        a = K.dxb_aec()
        a.n = len(self._w)
        a.start[:] = self._start
        a.stop[:] = self._stop
        a.weights = _dp(self._w)
        return a

    def __call__(self, pos):
        a = self._desc()
        p = (C.c_double * 3)(*pos)
        return _lib().dxb_aec_weight(C.byref(a), p)


class CTOrganAECFilter:
    """dxmc::CTOrganAECFilter — R:src/libopendxmc/beamsettingsmodel.cpp:349-437."""

    def __init__(self):
        self._d = K.dxb_organ_aec()
        self._d.low_weight = 0.6
        self._d.ramp_angle = 20 * DEG_TO_RAD()
        self._d.stop_angle = math.pi

    def useFilter(self):
        return bool(self._d.use_filter)

    def setUseFilter(self, on):
        self._d.use_filter = 1 if on else 0

    def compensateOutside(self):
        return bool(self._d.compensate_outside)

    def setCompensateOutside(self, on):
        self._d.compensate_outside = 1 if on else 0

    def startAngle(self):
        return self._d.start_angle

    def setStartAngle(self, a):
        self._d.start_angle = float(a)

    def startAngleDeg(self):
        return self._d.start_angle * RAD_TO_DEG()

    def setStartAngleDeg(self, a):
        self._d.start_angle = float(a) * DEG_TO_RAD()

    def stopAngle(self):
        return self._d.stop_angle

    def setStopAngle(self, a):
        self._d.stop_angle = float(a)

    def stopAngleDeg(self):
        return self._d.stop_angle * RAD_TO_DEG()

    def setStopAngleDeg(self, a):
        self._d.stop_angle = float(a) * DEG_TO_RAD()

    def rampAngle(self):
        return self._d.ramp_angle

    def setRampAngle(self, a):
        self._d.ramp_angle = abs(float(a))

    def rampAngleDeg(self):
        return self._d.ramp_angle * RAD_TO_DEG()

    def setRampAngleDeg(self, a):
        self._d.ramp_angle = abs(float(a)) * DEG_TO_RAD()

    def lowWeight(self):
        return self._d.low_weight

    def setLowWeight(self, w):
        self._d.low_weight = min(max(float(w), 0.0), 1.0)

    setLowWeightFactor = setLowWeight  # the name the reference calls (R:src/libopendxmc/beamsettingsmodel.cpp:407)

    def maxWeight(self):
        return _lib().dxb_organ_aec_max_weight(C.byref(self._d))

    def __call__(self, angle):
        return _lib().dxb_organ_aec_weight(C.byref(self._d), float(angle))


# --------------------------------------------------------------------------- beams
class _Beam:
    TYPE = -1

    def __init__(self):
        self._d = K.dxb_beam_desc()
        _lib().dxb_beam_desc_init(C.byref(self._d), self.TYPE)
        self._keep = {}

    # -- C ABI view
    def desc(self):
        """Synchronise derived data (spectra, filters) and return the dxb_beam_desc."""
        self._sync()
        return self._d

    def _sync(self):
        pass

    def _set_spectrum(self, slot, tube):
        e = tube.getEnergy()
        w = tube.getSpecter(e, True)
        self._keep[("spec", slot)] = (e, w)
        s = self._d.spectrum[slot]
        s.n = len(e)
        s.energy_kev = _dp(e)
        s.weight = _dp(w)

    def _set_bowtie(self, slot, bowtie):
        self._keep[("bow", slot)] = bowtie
        b = self._d.bowtie[slot]
        if bowtie is None or len(bowtie.angle) < 2:
            b.n = 0
        else:
            b.n = len(bowtie.angle)
            b.angle_rad = _dp(bowtie.angle)
            b.weight = _dp(bowtie.weight)

    # -- common API
    def numberOfExposures(self):
        return int(_lib().dxb_beam_number_of_exposures(C.byref(self.desc())))

    def numberOfParticlesPerExposure(self):
        return int(self._d.particles_per_exposure)

    def setNumberOfParticlesPerExposure(self, n):
        self._d.particles_per_exposure = int(n)

    def numberOfParticles(self):
        return int(_lib().dxb_beam_number_of_particles(C.byref(self.desc())))

    def exposure(self, i):
        e = K.dxb_exposure()
        _check(_lib().dxb_beam_exposure(C.byref(self.desc()), int(i), C.byref(e)), "exposure")
        return Exposure(e)


class Exposure:
    def __init__(self, e):
        self._e = e

    def position(self):
        return list(self._e.position)

    def directionCosines(self):
        return [list(self._e.cosines[0]), list(self._e.cosines[1])]

    def direction(self):
        return list(self._e.direction)

    def collimationHalfAngles(self):
        return list(self._e.half_angles)

    def weight(self):
        return self._e.weight

    def numberOfParticles(self):
        return int(self._e.n_particles)

    def tube(self):
        return int(self._e.tube)


class _TubeBeam(_Beam):
    """beams that own one dxmc::Tube (R:src/libopendxmc/beamsettingsmodel.cpp:257-346)."""

    def __init__(self, filtration=None):
        super().__init__()
        self._tube = Tube()
        for z, mm in (filtration or {}).items():
            self._tube.addFiltrationMaterial(z, mm)

    def tube(self):
        return self._tube

    def setTube(self, tube):
        self._tube = tube

    def setTubeVoltage(self, kv):
        self._tube.setVoltage(kv)

    def setTubeAnodeAngleDeg(self, a):
        self._tube.setAnodeAngleDeg(a)

    def addTubeFiltrationMaterial(self, Z, mm):
        self._tube.addFiltrationMaterial(Z, mm)

    def tubeFiltration(self, Z):
        return self._tube.filtration(Z)

    def clearTubeFiltrationMaterials(self):
        self._tube.clearFiltrationMaterials()

    def tubeAlHalfValueLayer(self):
        return self._tube.alHalfValueLayer()

    def tubeMeanSpecterEnergy(self):
        return self._tube.meanSpecterEnergy()

    def _sync(self):
        self._set_spectrum(0, self._tube)


class DXBeam(_TubeBeam):
    """dxmc::DXBeam<false> + OpenDXMC's subclass (R:src/libopendxmc/dxmc_specialization.cpp:22-90):
    rotation centre / source-patient distance / primary + secondary angle drive the pose."""
    TYPE = K.BEAM_DX

    def __init__(self, pos=None, cosines=None, filtration=None):
        super().__init__(filtration if filtration is not None else {13: 2.0, 29: 0.1})
        self._center = [0.0, 0.0, 0.0]
        self._spd = 100.0  # R:src/libopendxmc/dxmc_specialization.hpp:72-73
        self._sdd = 100.0
        self._angles = [0.0, 0.0]
        if pos is None and cosines is None:
            self._updatePosition()  # the OpenDXMC subclass constructor, R:src/libopendxmc/dxmc_specialization.cpp:22-26
        else:
            # dxmc::DXBeam<false>(pos, cosines, filtration) base form
            self.setPosition(pos if pos is not None else (0, 0, 0))
            self.setDirectionCosines(cosines if cosines is not None else ((1, 0, 0), (0, -1, 0)))
        self.setCollimation([20.0, 20.0])

    def position(self):
        return list(self._d.position)

    def setPosition(self, p):
        self._d.position[:] = [float(v) for v in p]

    def directionCosines(self):
        return [list(self._d.cosines[0]), list(self._d.cosines[1])]

    def setDirectionCosines(self, c):
        for k in range(2):
            n = math.sqrt(sum(float(v) ** 2 for v in c[k])) or 1.0
            self._d.cosines[k][:] = [float(v) / n for v in c[k]]

    def collimationHalfAngles(self):
        return list(self._d.half_angles)

    def setCollimationHalfAngles(self, a, b=None):
        if b is not None:
            a = [a, b]
        self._d.half_angles[:] = [abs(float(a[0])), abs(float(a[1]))]

    def collimationHalfAnglesDeg(self):
        return [v * RAD_TO_DEG() for v in self._d.half_angles]

    def setCollimationHalfAnglesDeg(self, a, b=None):
        if b is not None:
            a = [a, b]
        self.setCollimationHalfAngles([v * DEG_TO_RAD() for v in a])

    # OpenDXMC subclass: stored value is tan(size/2/SDD) (dxmc_specialization.cpp:46-60), reproduced as is
    def setCollimation(self, size_cm):
        self.setCollimationHalfAngles([math.tan(0.5 * abs(s) / self._sdd) for s in size_cm])

    def collimation(self):
        return [2.0 * self._sdd * math.atan(a) for a in self._d.half_angles]

    def sourceDetectorDistance(self):
        return self._sdd

    def setSourceDetectorDistance(self, d):
        # R:src/libopendxmc/dxmc_specialization.cpp:40-44: only m_SDD changes; the stored half angles stay
        self._sdd = abs(float(d))
        self._d.sdd = self._sdd
        self._updatePosition()

    def sourcePatientDistance(self):
        return self._spd

    def setSourcePatientDistance(self, d):
        self._spd = abs(float(d))
        self._updatePosition()

    def rotationCenter(self):
        return list(self._center)

    def setRotationCenter(self, c):
        self._center = [float(v) for v in c]
        self._updatePosition()

    def primaryAngleDeg(self):
        return self._angles[0] * RAD_TO_DEG()

    def secondaryAngleDeg(self):
        return self._angles[1] * RAD_TO_DEG()

    def setPrimaryAngleDeg(self, a):
        self._angles[0] = min(max(float(a), -180.0), 180.0) * DEG_TO_RAD()
        self._updatePosition()

    def setSecondaryAngleDeg(self, a):
        self._angles[1] = min(max(float(a), -90.0), 90.0) * DEG_TO_RAD()
        self._updatePosition()

    def _updatePosition(self):
        # R:src/libopendxmc/dxmc_specialization.cpp:78-90: base cosines {0,0,1},{-1,0,0}; rotate by the primary angle
        # about z, then by the secondary angle about -x; position = centre - SPD * (c0 x c1)
        def rot(v, axis, ang):
            k = np.array(axis, dtype=float)
            v = np.array(v, dtype=float)
            return v * math.cos(ang) + np.cross(k, v) * math.sin(ang) + k * np.dot(k, v) * (1 - math.cos(ang))
        c0, c1 = np.array([0.0, 0.0, 1.0]), np.array([-1.0, 0.0, 0.0])
        c0 = rot(c0, [0, 0, 1], self._angles[0])
        c1 = rot(c1, [0, 0, 1], self._angles[0])
        c0 = rot(c0, [-1, 0, 0], self._angles[1])
        c1 = rot(c1, [-1, 0, 0], self._angles[1])
        d = np.cross(c0, c1)
        self.setDirectionCosines([c0, c1])
        self.setPosition(np.array(self._center) - self._spd * d)

    def DAPvalue(self):
        return self._d.dap

    def setDAPvalue(self, v):
        self._d.dap = abs(float(v))

    def setNumberOfExposures(self, n):
        self._d.n_exposures = max(1, int(n))


class PencilBeam(_Beam):
    """dxmc::PencilBeam<false> — R:src/libopendxmc/beamsettingsmodel.cpp:637-711."""
    TYPE = K.BEAM_PENCIL

    def __init__(self, pos=(0, 0, 0), direction=(0, 0, 1), energy=60.0):
        super().__init__()
        self.setPosition(pos)
        self.setDirection(direction)
        self.setEnergy(energy)

    def position(self):
        return list(self._d.position)

    def setPosition(self, p):
        self._d.position[:] = [float(v) for v in p]

    def direction(self):
        return list(self._d.direction)

    def setDirection(self, v):
        n = math.sqrt(sum(float(x) ** 2 for x in v)) or 1.0
        self._d.direction[:] = [float(x) / n for x in v]

    def energy(self):
        return self._d.energy

    def setEnergy(self, e):
        self._d.energy = min(max(float(e), 1.0), 150.0)

    def airKerma(self):
        return self._d.air_kerma

    def setAirKerma(self, k):
        self._d.air_kerma = abs(float(k))

    def setNumberOfExposures(self, n):
        self._d.n_exposures = max(1, int(n))


class CBCTBeam(_TubeBeam):
    """dxmc::CBCTBeam<false> — R:src/libopendxmc/beamsettingsmodel.cpp:730-895."""
    TYPE = K.BEAM_CBCT

    def __init__(self, isocenter=(0, 0, 0), axis=(0, 0, 1), filtration=None):
        super().__init__(filtration if filtration is not None else {13: 2.0, 29: 0.1})
        self.setIsocenter(isocenter)
        self.setRotationAxis(axis)

    def isocenter(self):
        return list(self._d.isocenter)

    def setIsocenter(self, p):
        self._d.isocenter[:] = [float(v) for v in p]

    def rotationAxis(self):
        return list(self._d.direction)

    def setRotationAxis(self, v):
        n = math.sqrt(sum(float(x) ** 2 for x in v)) or 1.0
        self._d.direction[:] = [float(x) / n for x in v]

    def sourceDetectorDistance(self):
        return self._d.sdd

    def setSourceDetectorDistance(self, d):
        self._d.sdd = max(abs(float(d)), 1.0)

    def startAngle(self):
        return self._d.start_angle

    def setStartAngle(self, a):
        self._d.start_angle = float(a)

    def stopAngle(self):
        return self._d.stop_angle

    def setStopAngle(self, a):
        self._d.stop_angle = float(a)

    def stepAngle(self):
        return self._d.step_angle

    def setStepAngle(self, a):
        self._d.step_angle = max(abs(float(a)), 0.1 * DEG_TO_RAD())

    def startAngleDeg(self):
        return self.startAngle() * RAD_TO_DEG()

    def setStartAngleDeg(self, a):
        self.setStartAngle(a * DEG_TO_RAD())

    def stopAngleDeg(self):
        return self.stopAngle() * RAD_TO_DEG()

    def setStopAngleDeg(self, a):
        self.setStopAngle(a * DEG_TO_RAD())

    def stepAngleDeg(self):
        return self.stepAngle() * RAD_TO_DEG()

    def setStepAngleDeg(self, a):
        self.setStepAngle(a * DEG_TO_RAD())

    def collimationHalfAngles(self):
        return list(self._d.half_angles)

    def setCollimationHalfAngles(self, a, b=None):
        if b is not None:
            a = [a, b]
        self._d.half_angles[:] = [abs(float(a[0])), abs(float(a[1]))]

    def collimationHalfAnglesDeg(self):
        return [v * RAD_TO_DEG() for v in self._d.half_angles]

    def setCollimationHalfAnglesDeg(self, a, b=None):
        if b is not None:
            a = [a, b]
        self.setCollimationHalfAngles([v * DEG_TO_RAD() for v in a])

    def DAPvalue(self):
        return self._d.dap

    def setDAPvalue(self, v):
        self._d.dap = abs(float(v))


class _CTBase(_TubeBeam):
    def __init__(self, filtration=None):
        super().__init__(filtration if filtration is not None else {13: 9.0})
        self._bowtie = None
        self._organ = CTOrganAECFilter()

    def scanFieldOfView(self):
        return self._d.fov

    def setScanFieldOfView(self, v):
        self._d.fov = max(abs(float(v)), 1.0)

    def sourceDetectorDistance(self):
        return self._d.sdd

    def setSourceDetectorDistance(self, d):
        self._d.sdd = max(abs(float(d)), 1.0)

    def collimation(self):
        return self._d.collimation

    def setCollimation(self, c):
        self._d.collimation = max(abs(float(c)), 0.01)

    def startAngle(self):
        return self._d.start_angle

    def setStartAngle(self, a):
        self._d.start_angle = float(a)

    def startAngleDeg(self):
        return self._d.start_angle * RAD_TO_DEG()

    def setStartAngleDeg(self, a):
        self._d.start_angle = float(a) * DEG_TO_RAD()

    def stepAngle(self):
        return self._d.step_angle

    def setStepAngle(self, a):
        self._d.step_angle = max(abs(float(a)), 0.1 * DEG_TO_RAD())

    def stepAngleDeg(self):
        return self._d.step_angle * RAD_TO_DEG()

    def setStepAngleDeg(self, a):
        self.setStepAngle(float(a) * DEG_TO_RAD())

    def CTDIdiameter(self):
        return self._d.ctdi_diameter

    def setCTDIdiameter(self, d):
        self._d.ctdi_diameter = max(abs(float(d)), 3.0)

    def setBowtieFilter(self, bowtie):
        self._bowtie = bowtie

    def bowtieFilter(self):
        return self._bowtie

    def organAECFilter(self):
        return self._organ

    def _sync(self):
        super()._sync()
        self._set_bowtie(0, self._bowtie)
        self._d.organ_aec = self._organ._d


class CTSequentialBeam(_CTBase):
    """dxmc::CTSequentialBeam<false> (the reference's axial CT beam) — R:src/libopendxmc/beamsettingsmodel.cpp:921-1131."""
    TYPE = K.BEAM_CT_SEQUENTIAL

    def __init__(self, start=(0, 0, 0), normal=(0, 0, 1), filtration=None):
        super().__init__(filtration)
        self.setPosition(start)
        self.setScanNormal(normal)

    def position(self):
        return list(self._d.position)

    def setPosition(self, p):
        self._d.position[:] = [float(v) for v in p]

    def scanNormal(self):
        return list(self._d.direction)

    def setScanNormal(self, v):
        n = math.sqrt(sum(float(x) ** 2 for x in v)) or 1.0
        self._d.direction[:] = [float(x) / n for x in v]

    def numberOfSlices(self):
        return int(self._d.n_slices)

    def setNumberOfSlices(self, n):
        self._d.n_slices = max(1, int(n))

    def sliceSpacing(self):
        return self._d.slice_spacing

    def setSliceSpacing(self, s):
        self._d.slice_spacing = abs(float(s))

    def CTDIw(self):
        return self._d.ctdi

    def setCTDIw(self, v):
        self._d.ctdi = abs(float(v))


class CTDIBeam(CTSequentialBeam):
    """DXMClib-internal axial beam used by the CT calibration; one rotation, no organ AEC."""
    TYPE = K.BEAM_CTDI


class CTSpiralBeam(_CTBase):
    """dxmc::CTSpiralBeam<false> — R:src/libopendxmc/beamsettingsmodel.cpp:1158-1378."""
    TYPE = K.BEAM_CT_SPIRAL

    def __init__(self, start=(0, 0, 0), stop=(0, 0, 1), filtration=None):
        super().__init__(filtration)
        self.setStartStopPosition(start, stop)
        self._aec = CTAECFilter()

    def startPosition(self):
        return list(self._d.start)

    def stopPosition(self):
        return list(self._d.stop)

    def setStartPosition(self, p):
        self._d.start[:] = [float(v) for v in p]

    def setStopPosition(self, p):
        self._d.stop[:] = [float(v) for v in p]

    def setStartStopPosition(self, a, b):
        self.setStartPosition(a)
        self.setStopPosition(b)

    def pitch(self):
        return self._d.pitch

    def setPitch(self, p):
        self._d.pitch = max(abs(float(p)), 0.01)

    def CTDIvol(self):
        return self._d.ctdi

    def setCTDIvol(self, v):
        self._d.ctdi = abs(float(v))

    def AECFilter(self):
        return self._aec

    def setAECFilter(self, start_or_filter, stop=None, weights=None):
        if isinstance(start_or_filter, CTAECFilter):
            self._aec = start_or_filter
        else:
            self._aec = CTAECFilter(start_or_filter, stop, weights)

    def _sync(self):
        super()._sync()
        self._d.aec = self._aec._desc()
        self._keep["aec"] = self._aec


class CTSpiralDualEnergyBeam(CTSpiralBeam):
    """dxmc::CTSpiralDualEnergyBeam<false> — R:src/libopendxmc/beamsettingsmodel.cpp:1413-1832.
    exposure(2i) = tube A, exposure(2i+1) = tube B (R:src/libopendxmc/beamactorcontainer.cpp:134-146)."""
    TYPE = K.BEAM_CT_SPIRAL_DUAL

    def __init__(self, start=(0, 0, 0), stop=(0, 0, 1), filtration=None):
        super().__init__(start, stop, filtration)
        self._tubeB = Tube()
        for z, mm in (filtration if filtration is not None else {13: 9.0}).items():
            self._tubeB.addFiltrationMaterial(z, mm)
        self._bowtieB = None
        self._d.tube_b_offset_angle = 90.0 * DEG_TO_RAD()

    def tubeA(self):
        return self._tube

    def tubeB(self):
        return self._tubeB

    def setTubeAVoltage(self, kv):
        self._tube.setVoltage(kv)

    def setTubeBVoltage(self, kv):
        self._tubeB.setVoltage(kv)

    def setTubesAnodeAngleDeg(self, a):
        self._tube.setAnodeAngleDeg(a)
        self._tubeB.setAnodeAngleDeg(a)

    def addTubeAFiltrationMaterial(self, Z, mm):
        self._tube.addFiltrationMaterial(Z, mm)

    def addTubeBFiltrationMaterial(self, Z, mm):
        self._tubeB.addFiltrationMaterial(Z, mm)

    def tubeAFiltration(self, Z):
        return self._tube.filtration(Z)

    def tubeBFiltration(self, Z):
        return self._tubeB.filtration(Z)

    def scanFieldOfViewA(self):
        return self._d.fov

    def scanFieldOfViewB(self):
        return self._d.fov_b

    def setScanFieldOfViewA(self, v):
        self._d.fov = max(abs(float(v)), 1.0)

    def setScanFieldOfViewB(self, v):
        self._d.fov_b = max(abs(float(v)), 1.0)

    def tubeBoffsetAngle(self):
        return self._d.tube_b_offset_angle

    def setTubeBoffsetAngle(self, a):
        self._d.tube_b_offset_angle = float(a)

    def tubeBoffsetAngleDeg(self):
        return self._d.tube_b_offset_angle * RAD_TO_DEG()

    def setTubeBoffsetAngleDeg(self, a):
        self._d.tube_b_offset_angle = float(a) * DEG_TO_RAD()

    def relativeMasTubeA(self):
        return self._d.relative_mas_a

    def relativeMasTubeB(self):
        return self._d.relative_mas_b

    def setRelativeMasTubeA(self, v):
        self._d.relative_mas_a = abs(float(v))

    def setRelativeMasTubeB(self, v):
        self._d.relative_mas_b = abs(float(v))

    def setBowtieFilterA(self, b):
        self._bowtie = b

    def setBowtieFilterB(self, b):
        self._bowtieB = b

    def tubeAAlHalfValueLayer(self):
        return self._tube.alHalfValueLayer()

    def tubeBAlHalfValueLayer(self):
        return self._tubeB.alHalfValueLayer()

    def tubeAMeanSpecterEnergy(self):
        return self._tube.meanSpecterEnergy()

    def tubeBMeanSpecterEnergy(self):
        return self._tubeB.meanSpecterEnergy()

    def tubeRelativeWeightA(self):
        return self._relw()[0]

    def tubeRelativeWeightB(self):
        return self._relw()[1]

    def _relw(self):
        # weights of the two tubes as the exposures carry them (mean 1)
        d = self.desc()
        sa = float(np.sum(self._keep[("spec", 0)][1])) * d.relative_mas_a
        sb = float(np.sum(self._keep[("spec", 1)][1])) * d.relative_mas_b
        return 2 * sa / (sa + sb), 2 * sb / (sa + sb)

    def _sync(self):
        super()._sync()
        self._set_spectrum(1, self._tubeB)
        self._set_bowtie(1, self._bowtieB)


# --------------------------------------------------------------------------- progress
class TransportProgress:
    """dxmc::TransportProgress — R:src/libopendxmc/simulationpipeline.cpp:37,114-119,139,169,234,261."""

    def __init__(self):
        self._h = _lib().dxb_progress_create()

    def __del__(self):
        try:
            _lib().dxb_progress_destroy(self._h)
        except Exception:
            pass

    def progress(self):
        d, t = C.c_uint64(), C.c_uint64()
        _lib().dxb_progress_read(self._h, C.byref(d), C.byref(t))
        # never 0: the reference's timer divides by the total (R:src/libopendxmc/simulationpipeline.cpp:114-115)
        return d.value, max(t.value, 1)

    def message(self):
        buf = C.create_string_buffer(128)
        _lib().dxb_progress_message(self._h, buf, 128)
        return buf.value.decode()

    def continueSimulation(self):
        return bool(_lib().dxb_progress_continue(self._h))

    def setStopSimulation(self):
        _lib().dxb_progress_stop(self._h)

    def reset(self):
        _lib().dxb_progress_reset(self._h)


# --------------------------------------------------------------------------- world / grid / transport
class DoseScore:
    def __init__(self, dose, variance, events):
        self._v = (dose, variance, events)

    def dose(self):
        return self._v[0]

    def variance(self):
        return self._v[1]

    def standardDeviation(self):
        return math.sqrt(self._v[1])

    def numberOfEvents(self):
        return self._v[2]


class AAVoxelGrid:
    """dxmc::AAVoxelGrid<5, CORRECTION, 255> — R:src/libopendxmc/simulationpipeline.cpp:127,145-150,174-219."""

    def __init__(self, lowEnergyCorrection=1):
        self.lowEnergyCorrection = int(lowEnergyCorrection)
        self._dim = None
        self._spacing = [1.0, 1.0, 1.0]
        self._density = None
        self._material = None
        self._materials = None
        self._dose = None

    def setData(self, dim, density, materialIdx, materials):
        n = int(dim[0]) * int(dim[1]) * int(dim[2])
        density = np.ascontiguousarray(density, dtype=np.float64).reshape(-1)
        materialIdx = np.ascontiguousarray(materialIdx, dtype=np.uint8).reshape(-1)
        if n == 0 or density.size != n or materialIdx.size != n or not materials or len(materials) > 255:
            return False
        if int(materialIdx.max()) >= len(materials):
            return False
        self._dim = [int(v) for v in dim]
        self._density, self._material, self._materials = density, materialIdx, list(materials)
        self._dose = None
        return True

    def setSpacing(self, spacing):
        self._spacing = [abs(float(v)) for v in spacing]

    def spacing(self):
        return list(self._spacing)

    def dimensions(self):
        return list(self._dim)

    def size(self):
        return 0 if self._dim is None else self._dim[0] * self._dim[1] * self._dim[2]

    def doseScored(self, i):
        d, v, n = self._dose
        return DoseScore(float(d[i]), float(v[i]), int(n[i]))

    def doseArrays(self):
        """(dose[mGy], variance, events) for all voxels: one D2H copy instead of size() doseScored calls."""
        return self._dose


class World:
    """dxmc::World<AAVoxelGrid<...>> with a single item — R:src/libopendxmc/simulationpipeline.cpp:128-131,153."""

    def __init__(self, devices=None):
        self._item = None
        self._devices = devices
        self._ctx = None

    def addItem(self, item=None):
        self._item = item if item is not None else AAVoxelGrid()
        return self._item

    def build(self):
        g = self._item
        if g is None or g._dim is None:
            raise K.DxbError(K.DXB_ESTATE, "World.build", "no voxel grid data")
        lib = _lib()
        if self._ctx is None:
            h = K.VP()
            if self._devices:
                arr = (C.c_int * len(self._devices))(*self._devices)
                rc = lib.dxb_create(C.byref(h), arr, len(self._devices))
            else:
                rc = lib.dxb_create(C.byref(h), None, 0)
            _check(rc, "dxb_create")
            self._ctx = h
        mats = (K.VP * len(g._materials))(*[m._h for m in g._materials])
        _check(lib.dxb_set_materials(self._ctx, len(g._materials), mats), "dxb_set_materials", self._ctx)
        dim = (C.c_uint64 * 3)(*g._dim)
        sp = (C.c_double * 3)(*g._spacing)
        _check(lib.dxb_set_grid(self._ctx, dim, sp, _dp(g._density), g._material.ctypes.data_as(K.c_u8_p)), "dxb_set_grid", self._ctx)

    def close(self):
        if self._ctx is not None:
            _lib().dxb_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- extras over the reference API (C ABI pass-throughs)
    def ctx(self):
        return self._ctx

    def set_option(self, key, value):
        _check(_lib().dxb_set_option(self._ctx, key.encode(), float(value)), "dxb_set_option", self._ctx)

    def set_seed(self, seed):
        """base Philox key; restarts the beam counter (beam k after this call runs on key seed + k * stride, dxb.h)."""
        _check(_lib().dxb_set_seed(self._ctx, int(seed)), "dxb_set_seed", self._ctx)

    def last_beam_key(self):
        """Philox key of the last beam (what a test hands to the CPU oracle as its seed)."""
        return int(_lib().dxb_last_beam_key(self._ctx))

    def flush(self):
        _check(_lib().dxb_flush(self._ctx), "dxb_flush", self._ctx)

    def set_history_range(self, rank, world):
        _check(_lib().dxb_set_history_range(self._ctx, int(rank), int(world)), "dxb_set_history_range", self._ctx)

    def set_calibration_histories(self, n):
        _check(_lib().dxb_set_calibration_histories(self._ctx, int(n)), "dxb_set_calibration_histories", self._ctx)

    def run_stats(self):
        s = K.dxb_run_stats()
        _check(_lib().dxb_get_run_stats(self._ctx, C.byref(s)), "dxb_get_run_stats")
        return {f: getattr(s, f) for f, _ in K.dxb_run_stats._fields_}

    def energy_scored(self):
        n = self._item.size()
        e, e2, cnt = np.zeros(n), np.zeros(n), np.zeros(n, dtype=np.uint64)
        _check(_lib().dxb_get_energy_scored(self._ctx, _dp(e), _dp(e2), cnt.ctypes.data_as(K.c_u64_p)), "dxb_get_energy_scored", self._ctx)
        return e, e2, cnt

    def fetch_dose(self):
        n = self._item.size()
        d, v, cnt = np.zeros(n), np.zeros(n), np.zeros(n, dtype=np.uint64)
        _check(_lib().dxb_get_dose(self._ctx, _dp(d), _dp(v), cnt.ctypes.data_as(K.c_u64_p)), "dxb_get_dose", self._ctx)
        self._item._dose = (d, v, cnt)
        return self._item._dose

    def fetch_dose_range(self, begin, end, out=None):
        """sharded read-out (dxb_get_dose_range): voxels [begin, end) of dose, variance, events into full-size arrays."""
        n = self._item.size()
        d, v, cnt = out if out is not None else (np.zeros(n), np.zeros(n), np.zeros(n, dtype=np.uint64))
        _check(_lib().dxb_get_dose_range(self._ctx, int(begin), int(end), _dp(d), _dp(v), cnt.ctypes.data_as(K.c_u64_p)),
               "dxb_get_dose_range", self._ctx)
        return d, v, cnt

    def clear_dose(self):
        _check(_lib().dxb_clear_dose(self._ctx), "dxb_clear_dose", self._ctx)

    def dose_postprocessed(self, delete_air):
        n = self._item.size()
        d, v, cnt = np.zeros(n), np.zeros(n), np.zeros(n)
        units = C.create_string_buffer(4)
        _check(_lib().dxb_get_dose_postprocessed(self._ctx, 1 if delete_air else 0, _dp(d), _dp(v), _dp(cnt), units),
               "dxb_get_dose_postprocessed", self._ctx)
        return d, v, cnt, units.value.decode()

    def organ_dose(self, organ, n_organs):
        organ = np.ascontiguousarray(organ, dtype=np.uint8).reshape(-1)
        d, m, var = np.zeros(n_organs), np.zeros(n_organs), np.zeros(n_organs)
        cnt = np.zeros(n_organs, dtype=np.uint64)
        _check(_lib().dxb_organ_dose(self._ctx, organ.ctypes.data_as(K.c_u8_p), n_organs, _dp(d), _dp(m),
                                     cnt.ctypes.data_as(K.c_u64_p), _dp(var)), "dxb_organ_dose", self._ctx)
        return d, m, cnt, var

    def device_attenuation(self, material_index, energies, physics_mode=1):
        e = np.ascontiguousarray(energies, dtype=np.float64)
        out = np.zeros((len(e), 4), dtype=np.float32)
        _check(_lib().dxb_device_attenuation(self._ctx, material_index, physics_mode, _dp(e), len(e),
                                             out.ctypes.data_as(K.c_float_p)), "dxb_device_attenuation", self._ctx)
        return out

    def dense_box(self):
        """the dense box built with the grid: dict(built, useful, box[6] voxel indices, faces[6] cm, ratio[16] outside / global majorant)"""
        built, useful = C.c_int(), C.c_int()
        box = (C.c_int * 6)()
        faces = np.zeros(6, dtype=np.float32)
        ratio = np.ones(16, dtype=np.float32)
        _check(_lib().dxb_get_dense_box(self._ctx, C.byref(built), C.byref(useful), box, faces.ctypes.data_as(K.c_float_p),
                                        ratio.ctypes.data_as(K.c_float_p)), "dxb_get_dense_box", self._ctx)
        return {"built": bool(built.value), "useful": bool(useful.value), "box": [int(v) for v in box], "faces": faces, "ratio": ratio}

    def local_majorant(self):
        """the slab-local majorant table built with the grid: (n_slabs, shift, useful, ratio[n_slabs, 16], local / global majorant in (0, 1])"""
        n, sh, us = C.c_int(), C.c_int(), C.c_int()
        _check(_lib().dxb_get_local_majorant(self._ctx, C.byref(n), C.byref(sh), C.byref(us), None), "dxb_get_local_majorant", self._ctx)
        t = np.ones((max(n.value, 1), 16), dtype=np.float32)
        if n.value > 0:
            _check(_lib().dxb_get_local_majorant(self._ctx, None, None, None, t.ctypes.data_as(K.c_float_p)), "dxb_get_local_majorant", self._ctx)
        return n.value, sh.value, bool(us.value), t

    def device_majorant(self, energies):
        e = np.ascontiguousarray(energies, dtype=np.float64)
        out = np.zeros(len(e), dtype=np.float32)
        _check(_lib().dxb_device_majorant(self._ctx, _dp(e), len(e), out.ctypes.data_as(K.c_float_p)), "dxb_device_majorant", self._ctx)
        return out


class Transport:
    """dxmc::Transport — R:src/libopendxmc/simulationpipeline.cpp:155-165.
    operator()(world, beam, progress, useBeamCalibration) is __call__."""

    def __init__(self):
        self._threads = 0

    def setNumberOfThreads(self, n):
        # CPU worker threads have no meaning on the GPU path; kept for API compatibility (R:...simulationpipeline.cpp:157)
        self._threads = int(n)

    def numberOfThreads(self):
        return self._threads

    def __call__(self, world, beam, progress=None, useBeamCalibration=True):
        g = world._item
        if progress is not None:
            # DXMClib's Transport starts the progress object itself; the reference's driver leaves its stop flag raised
            # after every run (R:src/libopendxmc/simulationpipeline.cpp:234) and reuses the object
            progress.reset()
        rc = _lib().dxb_run(world._ctx, C.byref(beam.desc()), g.lowEnergyCorrection, 1 if useBeamCalibration else 0,
                            progress._h if progress is not None else None)
        if rc == K.DXB_ECANCELLED:
            return False
        _check(rc, "dxb_run", world._ctx)
        world.fetch_dose()
        return True

    def run_transport(self, world, beam, progress=None):
        """tallies only (multi-process sharding: reduce the tally buffer across ranks, then finish_beam)."""
        g = world._item
        rc = _lib().dxb_run_transport(world._ctx, C.byref(beam.desc()), g.lowEnergyCorrection,
                                      progress._h if progress is not None else None)
        _check(rc, "dxb_run_transport", world._ctx)

    def finish_beam(self, world, beam, useBeamCalibration=True):
        g = world._item
        f = C.c_double()
        rc = _lib().dxb_finish_beam(world._ctx, C.byref(beam.desc()), g.lowEnergyCorrection, 1 if useBeamCalibration else 0, C.byref(f))
        _check(rc, "dxb_finish_beam", world._ctx)
        return f.value

"""Pins the CPU oracle (test infrastructure) against known answers.

The reference's own tests hold NO vectors for this path ("parity unpinned"), so the oracle is
pinned against (i) the Random123 Philox4x32-10 known-answer vectors, (ii) analytic physics: Beer-Lambert
transmission through homogeneous and layered voxel columns (Woodcock tracking must reproduce exp(-sum mu t)
whatever the majorant), the Klein-Nishina and Thomson angular laws (chi-square), energy conservation in a
thick absorber, and (iii) the input fixtures mined from the reference's data files.
"""
import math

import numpy as np
import pytest


def test_philox_random123_known_answers(orc):
    # Random123 kat_vectors: philox4x32 10
    kat = [
        ([0, 0], [0, 0, 0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
        ([0xffffffff, 0xffffffff], [0xffffffff] * 4, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
        ([0xa4093822, 0x299f31d0], [0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
    ]
    for key, ctr, out in kat:
        assert [int(v) for v in orc.philox(key, ctr)] == out


def _column_world(dx, orc, mats, densities, layer_of_slice, n_slices, thickness, width=1e-4):
    """a 1 x 1 x n column of voxels so thin that any scattered photon leaves it at once: the number of real
    interactions per history is then the probability of a first collision, 1 - exp(-sum mu t)."""
    dim = [1, 1, n_slices]
    spacing = [width, width, thickness / n_slices]
    material = np.array([layer_of_slice(k) for k in range(n_slices)], dtype=np.uint8)
    density = np.array([densities[m] for m in material], dtype=np.float64)
    return orc.OracleWorld(dim, spacing, density, material, mats), density, material, spacing


@pytest.mark.parametrize("energy", [20.0, 60.0, 120.0])
def test_beer_lambert_homogeneous(dx, orc, energy):
    water = dx.Material.byNistName("Water, Liquid")
    t = 5.0
    ow, density, material, spacing = _column_world(dx, orc, [water], [1.0], lambda k: 0, 50, t)
    beam = dx.PencilBeam([0, 0, -10.0], [0, 0, 1], energy)
    beam.setNumberOfExposures(4)
    beam.setNumberOfParticlesPerExposure(250_000)
    _, _, _, st = ow.run(beam, 1, seed=1234)
    n = st["histories"]
    mu = water.attenuationValues(energy).sum() * 1.0
    p = 1.0 - math.exp(-mu * t)
    got = st["interactions"] / n
    sigma = math.sqrt(p * (1 - p) / n)
    # scattered photons can interact again before they leave the 1 um column: bounded by width * mu ~ 1e-4
    assert abs(got - p) < 4 * sigma + 2e-4, (got, p, sigma)


def test_beer_lambert_layered_majorant(dx, orc):
    """water / bone / air layers: the bone majorant makes most collisions in water and air virtual; the first-collision
    probability must still be 1 - exp(-sum mu_i t_i)."""
    names = ["Water, Liquid", "Bone, Cortical (ICRP)", "Air, Dry (near sea level)"]
    mats = [dx.Material.byNistName(n) for n in names]
    rho = [1.0, 1.85, 1.2e-3]
    t = 6.0
    n_slices = 60
    ow, density, material, spacing = _column_world(dx, orc, mats, rho, lambda k: (k // 10) % 3, n_slices, t)
    energy = 70.0
    beam = dx.PencilBeam([0, 0, -10.0], [0, 0, 1], energy)
    beam.setNumberOfExposures(4)
    beam.setNumberOfParticlesPerExposure(250_000)
    _, _, _, st = ow.run(beam, 1, seed=99)
    n = st["histories"]
    # densities are stored with 2^-17 relative rounding; irrelevant at this tolerance
    tau = sum(mats[m].attenuationValues(energy).sum() * rho[m] * spacing[2] for m in material)
    p = 1.0 - math.exp(-tau)
    got = st["interactions"] / n
    sigma = math.sqrt(p * (1 - p) / n)
    assert abs(got - p) < 4 * sigma + 2e-4, (got, p, sigma)
    # and the tentative-step count follows the majorant: mean virtual+real collisions = mu_max * path (for those that traverse)
    mu_max = ow.majorant(energy)
    assert mu_max == pytest.approx(max(mats[m].attenuationValues(energy).sum() * rho[m] for m in range(3)), rel=2e-5)


def test_energy_conservation_thick_absorber(dx, orc):
    """a pencil source in the middle of a 60 cm water cube (11 mean free paths to the nearest face at 30 keV):
    essentially every photon is absorbed, so deposited == emitted (cut-off deposit and Russian roulette are unbiased)."""
    water = dx.Material.byNistName("Water, Liquid")
    n = 16
    dim = [n, n, n]
    spacing = [60.0 / n] * 3
    ow = orc.OracleWorld(dim, spacing, np.ones(n ** 3), np.zeros(n ** 3, dtype=np.uint8), [water])
    beam = dx.PencilBeam([0, 0, 0.0], [0, 0, 1], 30.0)
    beam.setNumberOfExposures(2)
    beam.setNumberOfParticlesPerExposure(100_000)
    e, e2, cnt, st = ow.run(beam, 1, seed=5)
    assert st["energy_emitted_kev"] == pytest.approx(30.0 * st["histories"], rel=1e-9)
    assert e.sum() == pytest.approx(st["energy_emitted_kev"], rel=2e-3)
    assert st["energy_deposited_kev"] == pytest.approx(e.sum(), rel=1e-12)


def _chi2(samples, pdf, lo, hi, bins=40):
    hist, edges = np.histogram(samples, bins=bins, range=(lo, hi))
    xs = np.linspace(lo, hi, bins * 50 + 1)
    cdf = np.concatenate([[0], np.cumsum(0.5 * (pdf(xs[1:]) + pdf(xs[:-1])) * np.diff(xs))])
    cdf /= cdf[-1]
    expect = np.diff(np.interp(edges, xs, cdf)) * len(samples)
    mask = expect > 20
    chi2 = (((hist - expect) ** 2) / np.where(mask, expect, 1))[mask].sum()
    return chi2, int(mask.sum()) - 1


@pytest.mark.parametrize("energy", [30.0, 100.0])
def test_klein_nishina_angular_law(dx, orc, energy):
    water = dx.Material.byNistName("Water, Liquid")
    cos_t, ratio = orc.sample_compton(water, 0, energy, 400_000, seed=7)
    k = energy / 510.99895
    # kinematics
    assert np.allclose(ratio, 1.0 / (1.0 + k * (1.0 - cos_t)), rtol=1e-9)

    def pdf(c):
        e = 1.0 / (1.0 + k * (1.0 - c))
        return e * e * (e + 1.0 / e - (1.0 - c * c))
    chi2, dof = _chi2(cos_t, pdf, -1.0, 1.0)
    assert chi2 < dof + 5 * math.sqrt(2 * dof), (chi2, dof)


def test_livermore_compton_suppresses_forward_scatter(dx, orc):
    water = dx.Material.byNistName("Water, Liquid")
    c0, _ = orc.sample_compton(water, 0, 30.0, 200_000, seed=3)
    c1, _ = orc.sample_compton(water, 1, 30.0, 200_000, seed=3)
    k = 30.0 / 510.99895

    def pdf(c):
        e = 1.0 / (1.0 + k * (1.0 - c))
        x = 30.0 / 12.398419843 * np.sqrt(0.5 * (1.0 - c))
        s = np.array([water.scatterFactor(v) for v in x])
        return e * e * (e + 1.0 / e - (1.0 - c * c)) * s
    chi2, dof = _chi2(c1, pdf, -1.0, 1.0, bins=30)
    assert chi2 < dof + 5 * math.sqrt(2 * dof), (chi2, dof)
    assert (c1 > 0.97).mean() < 0.7 * (c0 > 0.97).mean()  # binding suppresses small momentum transfers


def test_thomson_and_form_factor_rayleigh(dx, orc):
    water = dx.Material.byNistName("Water, Liquid")
    c = orc.sample_rayleigh(water, 0, 60.0, 300_000, seed=11)
    chi2, dof = _chi2(c, lambda x: 1.0 + x * x, -1.0, 1.0)
    assert chi2 < dof + 5 * math.sqrt(2 * dof), (chi2, dof)
    # form-factor law: pdf(cos) ~ (1 + cos^2) F(x)^2, x = E/hc sqrt((1-cos)/2); forward peaked
    e = 40.0
    c1 = orc.sample_rayleigh(water, 1, e, 300_000, seed=12)

    def pdf(cc):
        x = e / 12.398419843 * np.sqrt(0.5 * (1.0 - cc))
        f = np.array([water.formFactor(v) for v in x])
        return (1.0 + cc * cc) * f * f
    chi2, dof = _chi2(c1, pdf, -1.0, 1.0, bins=30)
    # the CDF table is piecewise linear in x^2: allow a small model error on top of the statistics
    assert chi2 < dof + 12 * math.sqrt(2 * dof), (chi2, dof)
    assert c1.mean() > 0.7


def test_source_sampling_geometry_and_spectrum(dx, orc):
    wl = dx.workloads.ctdi_body_phantom(n=16, histories=360 * 2000, step_deg=1.0)
    beam = wl.beam
    n = 200_000
    pos, dirs, energy, weight = orc.sample_source(beam, 0, n)
    ppe = beam.numberOfParticlesPerExposure()
    ex = beam.exposure(0)
    assert np.allclose(pos[:ppe], ex.position(), atol=1e-5)
    assert np.allclose(np.linalg.norm(dirs, axis=1), 1.0, atol=1e-12)
    # fan / cone angles inside the collimation
    c0, c1 = [np.array(v) for v in ex.directionCosines()]
    hx, hy = ex.collimationHalfAngles()
    assert hx == pytest.approx(math.atan(50.0 / 119.0)) and hy == pytest.approx(math.atan(3.84 / 119.0))
    d0 = dirs[:ppe]
    assert np.all(np.abs(np.arcsin(d0 @ c0)) <= hx * (1 + 1e-6))
    assert np.all(np.abs(np.arcsin(d0 @ c1)) <= hy * (1 + 1e-6))
    # energies follow the tube spectrum
    e_nodes = beam.tube().getEnergy()
    spec = beam.tube().getSpecter(e_nodes, True)
    hist, _ = np.histogram(energy, bins=np.append(e_nodes, e_nodes[-1] + 1.0))
    expect = spec * n
    m = expect > 50
    chi2 = (((hist - expect) ** 2) / np.where(m, expect, 1))[m].sum()
    dof = int(m.sum()) - 1
    assert chi2 < dof + 6 * math.sqrt(2 * dof), (chi2, dof)
    # bowtie weights are normalised: the mean over the fan is 1
    assert weight.mean() == pytest.approx(1.0, abs=0.01)


def test_sharded_oracle_runs_sum_to_the_whole(dx, orc):
    wl = dx.workloads.ctdi_body_phantom(n=16, histories=150_000, step_deg=10.0)
    ow = orc.OracleWorld.from_workload(wl)
    e, e2, cnt, st = ow.run(wl.beam, 1, seed=21, threads=1)
    acc = np.zeros_like(e)
    acc_c = np.zeros_like(cnt)
    hist = 0
    for r in range(2):
        a, b, c, s = ow.run(wl.beam, 1, seed=21, threads=1, rank=r, world=2)
        acc += a
        acc_c += c
        hist += s["histories"]
    assert hist == st["histories"]
    assert np.array_equal(acc_c, cnt)
    assert np.allclose(acc, e, rtol=1e-9, atol=1e-9)


# ------------------------------------------------------------------ physics mode 2 (impulse approximation, fluorescence)
def test_impulse_approximation_doppler_broadening(dx, orc):
    """mode 2: the scattered energy at a fixed angle is no longer the single Compton line: it is broadened around it
    (moving electrons); transfers below a shell's binding energy are rejected, which shifts the tightly bound
    shells to lower scattered energies."""
    bone = dx.Material.byNistName("Bone, Cortical (ICRP)")
    energy = 60.0
    k = energy / 510.99895
    c1, r1 = orc.sample_compton(bone, 1, energy, 300_000, seed=11)
    c2, r2 = orc.sample_compton(bone, 2, energy, 300_000, seed=11)
    line = 1.0 / (1.0 + k * (1.0 - c2))
    assert np.allclose(r1, 1.0 / (1.0 + k * (1.0 - c1)), rtol=1e-9)  # mode 1 keeps free-electron kinematics
    dev = r2 / line - 1.0
    t = bone.tables()
    assert t.n_shells >= 1
    assert t.rest_compton_j0 > 0 and 0.5 < t.rest_electrons_fraction < 1.0
    assert (np.abs(dev) < 1e-12).mean() < 1e-3         # nobody sits exactly on the line any more
    assert 2e-3 < dev.std() < 0.1                      # visible Doppler width (a few per cent of E')
    assert abs(np.median(dev)) < 5e-3                  # the bulk (outer electrons) is centred on the Compton line
    back = c2 < -0.5                                   # back-scatter: broadest line (largest momentum transfer)
    fwd = c2 > 0.5
    assert dev[back].std() > dev[fwd].std()
    assert (r2 <= 1.0).all() and (r2 > 0).all()
    # the angular law is still Klein-Nishina x S(x) to first order: mean cosine within 2 %
    assert c2.mean() == pytest.approx(c1.mean(), abs=0.02)


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_energy_conservation_all_modes_with_fluorescence(dx, orc, mode):
    """thick calcium-rich absorber, 20 keV (above the Ca K edge, fluorescence 3.7 keV > cut-off): whatever the
    physics mode does with binding, Doppler shifts and fluorescence photons, deposited == emitted."""
    bone = dx.Material.byNistName("Bone, Cortical (ICRP)")
    n = 12
    ow = orc.OracleWorld([n, n, n], [40.0 / n] * 3, np.full(n ** 3, 1.85), np.zeros(n ** 3, dtype=np.uint8), [bone])
    beam = dx.PencilBeam([0, 0, 0.0], [0, 0, 1], 20.0)
    beam.setNumberOfExposures(2)
    beam.setNumberOfParticlesPerExposure(50_000)
    e, e2, cnt, st = ow.run(beam, mode, seed=9)
    assert e.sum() == pytest.approx(st["energy_emitted_kev"], rel=1e-3)
    if mode == 2:
        # fluorescence photons add scoring events: more deposits per history than in mode 1
        _, _, _, st1 = ow.run(beam, 1, seed=9)
        assert st["deposits"] > st1["deposits"]


def _roi_sigma(e_a, e2_a, e_b, e2_b, masks):
    worst = 0.0
    for m in masks.values():
        s = math.sqrt(e2_a[m].sum() + e2_b[m].sum())
        if s > 0:
            worst = max(worst, abs(e_a[m].sum() - e_b[m].sum()) / s)
    return worst


@pytest.mark.parametrize("config", ["C1", "C2-small"])
def test_device_mirroring_quantisations_are_harmless(dx, orc, config):
    """The oracle normally mirrors the device's storage formats (24-bit voxel densities, f32 majorant, bowtie knots,
    alias acceptance values, shell constants, exposure geometry) so that GPU and oracle differ by rounding only.  With
    the mirroring switched OFF the oracle runs on the f64 inputs as handed over; both modes draw the same Philox
    streams, so their dose maps must agree far inside the statistics (a quantisation that biased the transport would
    show up here): totals within 2e-4, every ROI within a fraction of a standard error, and almost all voxels equal."""
    if config == "C1":
        wl = dx.workloads.ctdi_body_phantom(n=32, histories=400_000, step_deg=5.0)
        masks = {"pmma": wl.material == 1, "air": wl.material == 0}
    else:
        wl = dx.workloads.ct_spiral_patient(scale=8, histories=400_000, step_deg=10.0)
        masks = {nm: wl.organ == i for i, nm in enumerate(wl.organ_names)}
    assert orc.load().orc_get_device_mirroring() == 1
    e, e2, cnt, st = orc.OracleWorld.from_workload(wl).run(wl.beam, 1)
    with orc.unmirrored():
        assert orc.load().orc_get_device_mirroring() == 0
        ow = orc.OracleWorld.from_workload(wl)
        f, f2, fcnt, ft = ow.run(wl.beam, 1)
        # the un-mirrored majorant is the exact f64 maximum of density x total attenuation
        tot = max(m.attenuationValues(60.0).sum() * float(wl.density[wl.material == i].max()) for i, m in enumerate(wl.materials)
                  if np.any(wl.material == i))
        assert ow.majorant(60.0) == pytest.approx(tot, rel=1e-9)
    assert orc.load().orc_get_device_mirroring() == 1
    assert st["histories"] == ft["histories"]
    assert abs(e.sum() - f.sum()) / e.sum() <= 2e-4
    for k in ("steps", "interactions", "deposits"):
        assert abs(st[k] - ft[k]) / st[k] <= 2e-4, k
    assert _roi_sigma(e, e2, f, f2, masks) <= 0.5
    assert np.count_nonzero(cnt == fcnt) / cnt.size >= 0.995
    assert not np.array_equal(e, f)  # the switch does change the inputs


@pytest.mark.parametrize("energy", [30.0, 80.0])
def test_beer_lambert_layered_with_slab_local_majorants(dx, orc, energy):
    """Woodcock tracking with SLAB-LOCAL majorants (steps stop on slab faces, the walk restarts with a fresh draw) must
    still give the first-collision probability 1 - exp(-sum mu t) of a layered water / bone / air column, with fewer
    tentative steps than under the global (bone) majorant."""
    water = dx.Material.byNistName("Water, Liquid")
    bone = dx.Material.byNistName("Bone, Cortical (ICRP)")
    air = dx.Material.byNistName("Air, Dry (near sea level)")
    mats, dens = [water, bone, air], [1.0, 1.92, 1.2e-3]
    n, t = 64, 8.0
    layer = lambda k: 0 if k < 24 else (1 if k < 36 else (2 if k < 48 else 0))
    beam = dx.PencilBeam([0.0, 0.0, -10.0], [0, 0, 1], energy)
    beam.setNumberOfExposures(4)
    beam.setNumberOfParticlesPerExposure(100_000)
    res = {}
    for local in (False, True):
        ow, density, material, spacing = _column_world(dx, orc, mats, dens, layer, n, t)
        if local:
            assert ow.build_local_majorant(2) == 16     # slabs of 4 layers: none straddles two media
        e, e2, cnt, st = ow.run(beam, 1)
        res[local] = st
    thick = t / n
    tau = sum(dens[layer(k)] * mats[layer(k)].attenuationValues(energy).sum() * thick for k in range(n))
    expect = 1.0 - math.exp(-tau)
    nh = res[True]["histories"]
    for local in (False, True):
        p = res[local]["interactions"] / nh
        sigma = math.sqrt(expect * (1 - expect) / nh)
        assert abs(p - expect) < 4 * sigma, (local, p, expect, sigma)
    assert res[True]["hops"] > 0 and res[False]["hops"] == 0
    assert res[True]["steps"] < 0.75 * res[False]["steps"]


def test_slab_local_majorants_leave_the_dose_unchanged(dx, orc):
    """ICRP-shaped phantom (54 media, teeth at 2.75 g/cm3 set the global majorant) under a chest spiral: with slab-local
    majorants the oracle takes far fewer tentative steps and scores a statistically identical dose."""
    wl = dx.workloads.icrp_phantom("AM", scale=4, histories=300_000)
    a = orc.OracleWorld.from_workload(wl)
    e, e2, cnt, st = a.run(wl.beam, 1)
    b = orc.OracleWorld.from_workload(wl)
    assert b.build_local_majorant(0) >= 2
    f, f2, fcnt, ft = b.run(wl.beam, 1)
    assert ft["hops"] > 0 and ft["steps"] < 0.8 * st["steps"]
    s = math.sqrt(e2.sum() + f2.sum())
    assert abs(e.sum() - f.sum()) / s < 4.0
    assert abs(st["interactions"] - ft["interactions"]) / st["interactions"] < 0.01
    masks = {"upper": np.repeat(np.arange(wl.dim[2]), wl.dim[0] * wl.dim[1]) >= wl.dim[2] // 2}
    masks["lower"] = ~masks["upper"]
    masks["dense"] = wl.density > 1.2
    assert _roi_sigma(e, e2, f, f2, masks) < 4.0


@pytest.mark.parametrize("energy", [30.0, 80.0])
def test_beer_lambert_through_the_dense_box(dx, orc, energy):
    """Dense-box tracking (flights through the thin voxels around a box that holds everything else, the global majorant
    inside): the first-collision probability of a pencil beam through air / water / bone / water / air must still be
    1 - exp(-sum mu t) - the air in front of and behind the box is crossed in flights, not in tentative steps."""
    water = dx.Material.byNistName("Water, Liquid")
    bone = dx.Material.byNistName("Bone, Cortical (ICRP)")
    air = dx.Material.byNistName("Air, Dry (near sea level)")
    mats, dens = [water, bone, air], [1.0, 1.92, 1.2e-3]
    n, t = 64, 16.0
    layer = lambda k: 2 if k < 20 else (0 if k < 32 else (1 if k < 38 else (0 if k < 44 else 2)))
    beam = dx.PencilBeam([0.0, 0.0, -10.0], [0, 0, 1], energy)
    beam.setNumberOfExposures(4)
    beam.setNumberOfParticlesPerExposure(100_000)
    res = {}
    for box in (False, True):
        ow, density, material, spacing = _column_world(dx, orc, mats, dens, layer, n, t)
        if box:
            assert ow.build_dense_box(0.02)
        e, e2, cnt, st = ow.run(beam, 1)
        res[box] = st
    thick = t / n
    tau = sum(dens[layer(k)] * mats[layer(k)].attenuationValues(energy).sum() * thick for k in range(n))
    expect = 1.0 - math.exp(-tau)
    nh = res[True]["histories"]
    for box in (False, True):
        p = res[box]["interactions"] / nh
        sigma = math.sqrt(expect * (1 - expect) / nh)
        assert abs(p - expect) < 4 * sigma + 2e-4, (box, p, expect, sigma)
    assert res[True]["hops"] > 0 and res[False]["hops"] == 0
    assert res[True]["steps"] < 0.8 * res[False]["steps"]


def test_dense_box_leaves_the_dose_unchanged(dx, orc):
    """C2-shaped patient (air around an elliptical body) under a spiral beam that is wider than the body: with the dense box
    the oracle takes less than half the tentative steps and scores a statistically identical dose in every tissue class."""
    wl = dx.workloads.ct_spiral_patient(scale=4, histories=400_000, step_deg=5.0)
    a = orc.OracleWorld.from_workload(wl)
    e, e2, cnt, st = a.run(wl.beam, 1)
    b = orc.OracleWorld.from_workload(wl)
    assert b.build_dense_box(0.02)
    f, f2, fcnt, ft = b.run(wl.beam, 1)
    assert ft["hops"] > 0 and ft["steps"] < 0.5 * st["steps"]
    s = math.sqrt(e2.sum() + f2.sum())
    assert abs(e.sum() - f.sum()) / s < 4.0
    assert abs(st["interactions"] - ft["interactions"]) / st["interactions"] < 0.01
    masks = {nm: wl.material.reshape(-1) == i for i, nm in enumerate(["air", "lung", "soft", "bone"])}
    assert _roi_sigma(e, e2, f, f2, masks) < 4.0


@pytest.mark.parametrize("case", ["dx_cone_child", "dual_source_thorax_mode2"])
def test_dense_box_other_beams_and_modes(dx, orc, case):
    """The dense box under beams that enter and leave through other faces than a fan beam does (a radiography cone on a
    child phantom: photons enter through the front face and leave through the back and the sides; a dual-source spiral in
    physics mode 2): fewer tentative steps, statistically the same deposited energy in every density class."""
    if case == "dx_cone_child":
        wl, mode = dx.workloads.icrp_phantom("10M", scale=6, histories=300_000, beam_kind="dx"), 1
    else:
        wl, mode = dx.workloads.ct_dual_source_thorax(scale=8, histories=300_000, step_deg=10.0), 2
    a = orc.OracleWorld.from_workload(wl)
    e, e2, cnt, st = a.run(wl.beam, mode)
    b = orc.OracleWorld.from_workload(wl)
    assert b.build_dense_box(0.02)
    f, f2, fcnt, ft = b.run(wl.beam, mode)
    assert ft["hops"] > 0 and ft["steps"] < 0.9 * st["steps"]
    assert abs(st["interactions"] - ft["interactions"]) / st["interactions"] < 0.015
    s = math.sqrt(e2.sum() + f2.sum())
    assert abs(e.sum() - f.sum()) / s < 4.0
    rho = np.asarray(wl.density).reshape(-1)
    masks = {"thin": rho <= 0.05, "lung-like": (rho > 0.05) & (rho <= 0.6), "soft": (rho > 0.6) & (rho <= 1.2), "dense": rho > 1.2}
    masks = {k: m for k, m in masks.items() if m.any() and e[m].sum() > 0}
    assert _roi_sigma(e, e2, f, f2, masks) < 4.0

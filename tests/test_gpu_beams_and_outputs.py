"""GPU parity, part 2: every beam type, the scaled C3/C4/C5 configurations, calibration, post-processing, per-organ
dose, CT segmentation, progress / cancel, and in-process multi-GPU invariance — CUDA path through the C ABI vs the oracle."""
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np
import pytest

from parity import assert_same_stream_parity, mirror_local_majorant

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu
SEED = 0x0DDC0FFEE


def _compare_run(dx, orc, wl, mode=1, rois=None, tol=5e-3):
    world = wl.build_world(mode, [0])
    tr = dx.Transport()
    tr.run_transport(world, wl.beam)
    e, e2, cnt = world.energy_scored()
    st = world.run_stats()
    ow = orc.OracleWorld.from_workload(wl)
    assert mirror_local_majorant(world, ow) == bool(st["local_majorant"])
    oe, oe2, ocnt, ost = ow.run(wl.beam, mode, SEED)
    assert st["histories"] == ost["histories"] == wl.beam.numberOfParticles()
    assert st["hops"] == pytest.approx(ost["hops"], rel=1e-4, abs=30)
    assert abs(st["energy_emitted_kev"] - ost["energy_emitted_kev"]) / ost["energy_emitted_kev"] < 1e-5
    assert abs(e.sum() - oe.sum()) / oe.sum() <= tol, (e.sum(), oe.sum())
    for k in ("steps", "interactions", "deposits"):
        assert abs(st[k] - ost[k]) / max(ost[k], 1) < tol, (k, st[k], ost[k])
    assert_same_stream_parity(e, cnt, st, oe, ocnt, ost, f"{wl.name} mode {mode}", voxel_cm=min(wl.spacing), mode=mode)
    for name, m in (rois or {}).items():
        a, b = e[m].sum(), oe[m].sum()
        s = np.sqrt(e2[m].sum() + oe2[m].sum())
        assert s == 0 or abs(a - b) / s <= 3.0, (name, a, b, s)
    world.close()
    return e, oe


def _small_block(dx):
    """24^3 water block with a bone insert and an air gap, 0.5 cm voxels."""
    n = 24
    names = ["Air, Dry (near sea level)", "Water, Liquid", "Bone, Cortical (ICRP)"]
    mats = [dx.Material.byNistName(nm) for nm in names]
    material = np.ones((n, n, n), dtype=np.uint8)
    material[:, :, :3] = 0
    material[8:16, 8:16, 10:14] = 2
    density = np.choose(material, [1.2e-3, 1.0, 1.85]).astype(np.float64)
    return [n, n, n], [0.5, 0.5, 0.5], density.reshape(-1), material.reshape(-1), mats, names


@pytest.mark.parametrize("kind", ["dx", "pencil", "cbct", "sequential", "spiral_aec_organ", "dual"])
def test_every_beam_type_matches_the_oracle(dx, orc, kind):
    dim, sp, density, material, mats, names = _small_block(dx)
    bt = dx.workloads.read_bowtie_filters()[dx.workloads.DEFAULT_BOWTIE]
    if kind == "dx":
        beam = dx.DXBeam()
        beam.setTubeVoltage(80)
        beam.setRotationCenter([0, 0, 0])
        beam.setSourcePatientDistance(60)
        beam.setPrimaryAngleDeg(20)
        beam.setSecondaryAngleDeg(10)
        beam.setCollimation([12, 9])
        beam.setNumberOfExposures(8)
        beam.setNumberOfParticlesPerExposure(60_000)
    elif kind == "pencil":
        beam = dx.PencilBeam([0.3, -0.2, -20], [0.05, 0.02, 1], 45.0)
        beam.setNumberOfExposures(4)
        beam.setNumberOfParticlesPerExposure(100_000)
    elif kind == "cbct":
        beam = dx.CBCTBeam([0, 0, 0], [0, 0.1, 1])
        beam.setTubeVoltage(100)
        beam.setSourceDetectorDistance(90)
        beam.setStartAngleDeg(0)
        beam.setStopAngleDeg(200)
        beam.setStepAngleDeg(10)
        beam.setCollimationHalfAnglesDeg(6, 5)
        beam.setNumberOfParticlesPerExposure(25_000)
    elif kind == "sequential":
        beam = dx.CTSequentialBeam([0, 0, -3], [0, 0, 1], {13: 9.0})
        beam.setNumberOfSlices(3)
        beam.setSliceSpacing(3.0)
        beam.setCollimation(2.0)
        beam.setScanFieldOfView(20)
        beam.setStepAngleDeg(10)
        beam.setBowtieFilter(bt)
        beam.setNumberOfParticlesPerExposure(5_000)
    elif kind == "spiral_aec_organ":
        beam = dx.CTSpiralBeam([0, 0, -5], [0, 0, 5], {13: 9.0})
        beam.setCollimation(2.0)
        beam.setPitch(0.9)
        beam.setScanFieldOfView(20)
        beam.setStepAngleDeg(10)
        beam.setBowtieFilter(bt)
        beam.setAECFilter([0, 0, -5], [0, 0, 5], [1.0, 2.5, 0.7, 1.3])
        o = beam.organAECFilter()
        o.setUseFilter(True)
        o.setStartAngleDeg(-40)
        o.setStopAngleDeg(40)
        o.setLowWeightFactor(0.4)
        o.setCompensateOutside(True)
        beam.setNumberOfParticlesPerExposure(3_000)
    else:
        beam = dx.CTSpiralDualEnergyBeam([0, 0, -5], [0, 0, 5], {13: 9.0})
        beam.setTubeAVoltage(140)
        beam.setTubeBVoltage(80)
        beam.addTubeAFiltrationMaterial(50, 0.4)
        beam.setTubeBoffsetAngleDeg(95)
        beam.setScanFieldOfViewA(20)
        beam.setScanFieldOfViewB(13)
        beam.setCollimation(2.0)
        beam.setPitch(1.5)
        beam.setStepAngleDeg(10)
        beam.setRelativeMasTubeB(2.0)
        beam.setBowtieFilterA(bt)
        beam.setBowtieFilterB(bt)
        beam.setNumberOfParticlesPerExposure(2_500)
    wl = dx.workloads.Workload(kind, dim, sp, density, material, mats, names, beam)
    rois = {"bone": material == 2, "water": material == 1, "air": material == 0}
    for mode in (0, 1):
        _compare_run(dx, orc, wl, mode, rois)


def test_c3_icrp_shape_per_organ_dose(dx, orc):
    """C3 scaled: ICRP AM shape, 53 media from the real tables, chest spiral CT; per-organ dose within 3 combined sigma."""
    wl = dx.workloads.icrp_phantom("AM", scale=3, histories=3_000_000)
    assert len(wl.materials) > 40  # more tables than fit in shared memory at once -> global-memory table path too
    world = wl.build_world(1, [0])
    tr = dx.Transport()
    assert tr(world, wl.beam, None, False)
    d, v, n = world._item.doseArrays()
    ow = orc.OracleWorld.from_workload(wl)
    # 54 media, teeth (2.75 g/cm3) set the global majorant: the kernel tracks with slab-local majorants here
    assert mirror_local_majorant(world, ow) and world.run_stats()["local_majorant"] == 1 and world.run_stats()["hops"] > 0
    od, ov, on, ost = ow.transport(wl.beam, 1, False, SEED)
    vol = wl.spacing[0] * wl.spacing[1] * wl.spacing[2]
    n_org = len(wl.organ_names)
    gd, gm, gc, gv = world.organ_dose(wl.organ, n_org)
    cd, cm, cc = orc.organ_dose(od, wl.density, wl.organ, vol, n_org)
    # the device reduction reproduces the reference formula on its own dose array
    rd, rm, rc = orc.organ_dose(d, wl.density, wl.organ, vol, n_org)
    ok = rm > 0
    assert np.array_equal(gc, rc)
    assert np.allclose(gd[ok], rd[ok], rtol=2e-5, atol=0) and np.allclose(gm[ok], rm[ok], rtol=2e-5)
    # GPU vs oracle per organ: 3 combined standard errors (variance of the organ mean from the voxel variances)
    var_c = np.zeros(n_org)
    var_g = np.zeros(n_org)
    mass = wl.density * vol
    np.add.at(var_c, wl.organ, ov * mass ** 2)
    np.add.at(var_g, wl.organ, v * mass ** 2)
    checked = 0
    for o in range(n_org):
        if cm[o] <= 0 or cd[o] <= 0:
            continue
        s = np.sqrt(var_c[o] + var_g[o]) / cm[o]
        assert abs(gd[o] - cd[o]) <= 3.0 * s + 1e-5 * cd[o], (wl.organ_names[o], gd[o], cd[o], s)
        checked += 1
    assert checked > 50
    world.close()


def test_c4_dual_source_aec_small(dx, orc):
    wl = dx.workloads.ct_dual_source_thorax(scale=8, histories=2_000_000, step_deg=5.0)
    rois = {nm: wl.organ == i for i, nm in enumerate(wl.organ_names)}
    _compare_run(dx, orc, wl, 1, rois)


def test_c5_child_dx_small(dx, orc):
    wl = dx.workloads.icrp_phantom("10M", scale=6, histories=2_000_000, beam_kind="dx")
    _compare_run(dx, orc, wl, 1, {"all": np.ones(wl.n_voxels, dtype=bool)})


def test_calibrated_dose_dx_and_ct(dx, orc):
    dim, sp, density, material, mats, names = _small_block(dx)
    # DX: DAP calibration (analytic)
    beam = dx.DXBeam()
    beam.setTubeVoltage(70)
    beam.setSourcePatientDistance(60)
    beam.setCollimation([10, 10])
    beam.setDAPvalue(2.5)
    beam.setNumberOfExposures(4)
    beam.setNumberOfParticlesPerExposure(100_000)
    wl = dx.workloads.Workload("dxcal", dim, sp, density, material, mats, names, beam)
    world = wl.build_world(1, [0])
    tr = dx.Transport()
    assert tr(world, beam, None, True)
    d, v, n = world._item.doseArrays()
    ow = orc.OracleWorld.from_workload(wl)
    mirror_local_majorant(world, ow)
    od, ov, on, ost = ow.transport(beam, 1, True, SEED)
    assert world.run_stats()["calibration_factor"] == pytest.approx(ost["calibration_factor"], rel=1e-9)
    assert d.sum() == pytest.approx(od.sum(), rel=5e-3)
    sel = n > 20
    assert np.all(v[sel] > 0)
    # a second beam accumulates into the same dose score (repeated transport() on one world); replayed with the same key
    world.set_seed(SEED)
    assert tr(world, beam, None, True)
    d2, v2, n2 = world._item.doseArrays()
    assert d2.sum() == pytest.approx(2 * d.sum(), rel=1e-9) and int(n2.sum()) == 2 * int(n.sum())
    world.clear_dose()
    assert world.fetch_dose()[0].sum() == 0
    # CT: nested CTDI run; the factor of GPU and oracle agree statistically
    ct = dx.CTSpiralBeam([0, 0, -4], [0, 0, 4], {13: 9.0})
    ct.setStepAngleDeg(10)
    ct.setCTDIvol(12.0)
    ct.setNumberOfParticlesPerExposure(10_000)
    world.set_calibration_histories(3_600_000)
    assert tr(world, ct, None, True)
    f_gpu = world.run_stats()["calibration_factor"]
    f_cpu = ow.ct_calibration(ct, 1, world.last_beam_key(), 3_600_000)
    assert f_gpu == pytest.approx(f_cpu, rel=0.02)
    world.close()


def test_postprocess_matches_reference_rules(dx, orc):
    wl = dx.workloads.ctdi_body_phantom(n=32, histories=400_000, step_deg=5.0)
    world = wl.build_world(1, [0])
    tr = dx.Transport()
    assert tr(world, wl.beam, None, True)
    d, v, n = world._item.doseArrays()
    for delete_air in (False, True):
        gd, gv, gn, units = world.dose_postprocessed(delete_air)
        od, ov, on, ounits = orc.postprocess(d, v, n.astype(np.float64), wl.material, delete_air)
        assert units == ounits
        assert np.array_equal(gd, od) and np.array_equal(gn, on) and np.allclose(gv, ov, rtol=1e-15)
    world.close()


def test_ct_segmentation_matches_oracle(dx, orc):
    wl = dx.workloads.ctdi_body_phantom(n=16, histories=1000)
    world = wl.build_world(1, [0])
    from opendxmc_b200 import _capi as K
    lib = K.load()
    rng = np.random.default_rng(3)
    hu = np.concatenate([rng.uniform(-1100, 2500, 200_000), [-1000.0, 0.0, 55.0, 3000.0]])
    tube = dx.Tube(120.0)
    tube.setAlFiltration(9.0)
    mat = np.zeros(hu.size, dtype=np.uint8)
    dens = np.zeros(hu.size)
    handles = (K.VP * 5)()
    rc = lib.dxb_segment_ct(world.ctx(), hu.ctypes.data_as(K.c_double_p), hu.size, tube._sync(), mat.ctypes.data_as(K.c_u8_p),
                            dens.ctypes.data_as(K.c_double_p), handles)
    assert rc == 0
    # restate R:src/libopendxmc/ctsegmentationpipeline.cpp:61-156 with the library's own materials / tube
    names = ["Air, Dry (near sea level)", "Adipose Tissue (ICRP)", "Tissue, Soft (ICRP)", "Muscle, Skeletal", "Bone, Cortical (ICRP)"]
    mats = [dx.Material.byNistName(nm) for nm in names]
    dn = [dx.NISTMaterials.density(nm) for nm in names]
    dn[-1] = 1.09
    e = tube.getEnergy()
    w = tube.getSpecter(e, True)
    water, air = dx.Material.byNistName("Water, Liquid"), mats[0]
    wd, ad = 1.0, dx.NISTMaterials.density(names[0])
    uw = np.array([water.attenuationValues(x).sum() for x in e])
    ua = np.array([air.attenuationValues(x).sum() for x in e])
    um = [np.array([m.attenuationValues(x).sum() for x in e]) for m in mats]
    HU = [1000 * np.sum(w * (um[i] * dn[i] - uw * wd) / (uw * wd - ua * ad)) for i in range(5)]
    sep = np.array([(HU[i] + HU[i + 1]) / 2 for i in range(4)])
    att = np.array([np.sum(w * um[i]) for i in range(5)])
    omat = np.zeros(hu.size, dtype=np.uint8)
    odens = np.zeros(hu.size)
    orc.load().orc_segment(hu.ctypes.data_as(K.c_double_p), hu.size, sep.ctypes.data_as(K.c_double_p), 4, att.ctypes.data_as(K.c_double_p),
                           float(np.sum(w * uw) * wd), float(np.sum(w * ua) * ad), omat.ctypes.data_as(K.c_u8_p), odens.ctypes.data_as(K.c_double_p))
    assert np.array_equal(mat, omat)
    assert np.allclose(dens, odens, rtol=1e-12, atol=1e-15)
    assert set(np.unique(mat)) == {0, 1, 2, 3, 4}
    for h in handles:
        lib.dxb_material_destroy(h)
    world.close()


def test_sharded_dose_read_out(dx):
    """dxb_get_dose_range writes exactly the voxels of the range and agrees with the full read-out."""
    wl = dx.workloads.ctdi_body_phantom(n=32, histories=300_000, step_deg=10.0)
    world = wl.build_world(1, [0])
    assert dx.Transport()(world, wl.beam, None, False)
    d, v, c = world.fetch_dose()
    n = d.size
    out = (np.full(n, -1.0), np.full(n, -1.0), np.full(n, 7, dtype=np.uint64))
    cuts = [0, n // 3, n // 3, (2 * n) // 3 + 5]
    for b, e in zip(cuts[:-1], cuts[1:]):
        world.fetch_dose_range(b, e, out)
    last = cuts[-1]
    assert np.array_equal(out[0][:last], d[:last]) and np.array_equal(out[1][:last], v[:last]) and np.array_equal(out[2][:last], c[:last])
    assert np.all(out[0][last:] == -1.0) and np.all(out[2][last:] == 7)
    with pytest.raises(Exception):
        world.fetch_dose_range(5, n + 1, out)
    world.close()


def test_progress_and_cancel(dx):
    wl = dx.workloads.ct_spiral_patient(scale=4, histories=400_000_000, step_deg=5.0)
    world = wl.build_world(1, [0])
    world.set_option("batch_histories", 1 << 22)  # stop must be observed within one batch
    prog = dx.TransportProgress()
    tr = dx.Transport()
    result = {}

    def run():
        result["ok"] = tr(world, wl.beam, prog, False)
    t = threading.Thread(target=run)
    t0 = time.time()
    t.start()
    seen = 0
    while time.time() - t0 < 20:
        done, total = prog.progress()
        if done > 0:
            seen = done
            assert total == wl.beam.numberOfParticles()
            assert "Remaining" in prog.message()
            break
        time.sleep(0.001)
    prog.setStopSimulation()
    t.join(timeout=60)
    assert not t.is_alive()
    assert result["ok"] is False and seen > 0  # cancelled is distinguishable from finished
    assert not prog.continueSimulation()
    world.close()


def test_in_process_multi_gpu_is_bit_identical(dx):
    """dxb_create with several devices (what dxmc::Transport reaches with DXMC_B200_DEVICES=0,1,...): sharded upload,
    sharded nested calibration run, double-buffered tallies with the pipelined copy-engine exchange, distributed dose
    score.  Tallies, calibration factor, DOSE, variance, events, the reference's post-processing and the per-organ
    dose must equal the single-GPU results bit for bit."""
    from opendxmc_b200 import _capi as K
    lib = K.load()
    n_dev = lib.dxb_device_count()
    if n_dev < 2:
        pytest.skip("needs 2 GPUs")
    devs = list(range(min(n_dev, 4)))
    wl = dx.workloads.ct_spiral_patient(scale=4, histories=2_000_000, step_deg=5.0)
    w1 = wl.build_world(1, [0])
    wn = wl.build_world(1, devs)
    tr = dx.Transport()
    # (1) the tallies of one beam, read before any exchange
    tr.run_transport(w1, wl.beam)
    tr.run_transport(wn, wl.beam)
    for x, y in zip(w1.energy_scored(), wn.energy_scored()):
        assert np.array_equal(x, y)
    s1, sn = w1.run_stats(), wn.run_stats()
    for k in ("histories", "steps", "interactions", "deposits"):
        assert s1[k] == sn[k], k
    # (2) four beams back to back without reading anything in between: the exchange of beam i runs under the transport
    # of beam i + 1 and the tally buffers alternate; the first and the last beam are CT-calibrated (nested run)
    desc = wl.beam.desc()
    factors = {}
    for w in (w1, wn):
        w.set_seed(SEED)
        w.set_calibration_histories(720_000)
        factors[id(w)] = []
        for k in range(4):
            assert lib.dxb_run_transport(w.ctx(), C.byref(desc), 1, None) == 0, lib.dxb_last_error(w.ctx())
            f = C.c_double()
            assert lib.dxb_finish_beam(w.ctx(), C.byref(desc), 1, 1 if k in (0, 3) else 0, C.byref(f)) == 0, lib.dxb_last_error(w.ctx())
            # a beam is finished once
            assert lib.dxb_finish_beam(w.ctx(), C.byref(desc), 1, 0, None) == K.DXB_ESTATE
            factors[id(w)].append(f.value)
    assert factors[id(w1)] == factors[id(wn)]
    d1, dn = [a.copy() for a in w1.fetch_dose()], [a.copy() for a in wn.fetch_dose()]
    for a, b, name in zip(d1, dn, ("dose", "variance", "events")):
        assert np.array_equal(a, b), f"{name}: {np.count_nonzero(a != b)} voxels differ"
    assert d1[2].sum() > 0
    # (3) ranged read-out, the reference's post-processing and the per-organ dose run on the gathered score
    nvox = d1[0].size
    part = wn.fetch_dose_range(nvox // 5, nvox // 2 + 3)
    assert np.array_equal(part[0][nvox // 5:nvox // 2 + 3], d1[0][nvox // 5:nvox // 2 + 3]) and part[0][:nvox // 5].sum() == 0
    for delete_air in (False, True):
        for a, b in zip(w1.dose_postprocessed(delete_air), wn.dose_postprocessed(delete_air)):
            assert np.array_equal(a, b)
    # (the per-organ reduction adds doubles with atomics: last bits depend on the arrival order, on one GPU too)
    for a, b in zip(w1.organ_dose(wl.organ, len(wl.organ_names)), wn.organ_dose(wl.organ, len(wl.organ_names))):
        assert np.allclose(a, b, rtol=1e-12, atol=0)
    # (4) a new grid on the same context (setData again), then one more beam
    for w in (w1, wn):
        w.build()
        w.set_seed(SEED + 5)
        assert tr(w, wl.beam, None, False)
    for a, b in zip(w1.fetch_dose(), wn.fetch_dose()):
        assert np.array_equal(a, b)
    w1.close()
    wn.close()


def test_ipc_pipelined_exchange_matches_single_gpu_bit_for_bit():
    """one process per GPU with the library-managed exchange (CUDA IPC peers, copy-engine pulls under the next beam's
    transport) against the single-GPU dose score - tests/mp_ipc_exchange.py under torchrun."""
    import subprocess
    from opendxmc_b200 import _capi as K
    n = K.load().dxb_device_count()
    if n < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={min(n, 4)}", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tests", "mp_ipc_exchange.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0 and "IPC_OK" in out.stdout and "NCCL_REDUCE_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


def test_fused_exchange_matches_single_gpu_bit_for_bit():
    """one process per GPU: symmetric-memory tallies + dxb_finish_beam_sharded (NVSwitch multicast sum and P2P pull)
    against the single-GPU dose score — tests/mp_fused_exchange.py under torchrun."""
    import os
    import subprocess
    import sys
    from opendxmc_b200 import _capi as K
    n = K.load().dxb_device_count()
    if n < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={min(n, 4)}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(root, "tests", "mp_fused_exchange.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=root)
    assert out.returncode == 0 and "FUSED_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "opendxmc_ref")),
                    reason="needs oracle/_ref/opendxmc_ref, built where the reference tree is mounted (the binary travels to the GPU box)")
def test_reference_pipeline_binary_matches_python_mirror():
    """OpenDXMC's own SimulationPipeline (compiled unmodified, oracle/ref_driver.cpp) against the Python mirror."""
    sys.path.insert(0, os.path.join(ROOT, "profiles"))
    import run_reference_pipeline as rp
    same, units, ref_units, mine, ref = rp.run(1, 1, 20000)
    assert units == ref_units
    assert np.array_equal(mine[2], ref[2])
    assert np.allclose(mine[0], ref[0], rtol=1e-12, atol=0) and np.allclose(mine[1], ref[1], rtol=1e-12, atol=0)


def test_icrp_import_on_the_device(dx):
    """dxb_icrp_import (presence scan + one look-up-table gather per voxel on the GPU) against (a) the golden
    outputs of the reference's own importPhantom (tests/golden/icrp_import_golden.json, made by tests/golden/make_icrp_golden.py)
    and (b) the host-side plan on a full-size organ array of the ICRP AM shape (254 x 127 x 222, odd length: vector + tail path)."""
    import json
    world = dx.workloads.ctdi_body_phantom(n=16, histories=1000).build_world(1, [0])   # any context with a device
    golden = json.load(open(os.path.join(ROOT, "tests", "golden", "icrp_import_golden.json")))
    for case in golden["cases"]:
        organs_text, media_text = dx.workloads.icrp_dat_text(case["phantom"])
        raw = np.array(case["input"], dtype=np.uint8)
        organ, names, material, density, media_names, comps = dx.workloads.icrp_import(raw, organs_text, media_text, case["remove_arms"], world=world)
        assert np.array_equal(organ, case["organ"]) and names == case["organ_names"], case["name"]
        assert np.array_equal(material, case["material"]) and np.array_equal(density, np.array(case["density"])), case["name"]
        assert media_names == case["media_names"]
    wl = dx.workloads.icrp_phantom("AM", scale=1, histories=1000)
    raw = wl.organ_raw if hasattr(wl, "organ_raw") else None
    if raw is None:
        rng = np.random.default_rng(3)
        ids = np.array([o["id"] for o in dx.workloads.icrp_tables()["AM"]["organs"]] + [0], dtype=np.uint8)
        raw = rng.choice(ids, 254 * 127 * 222 - 3).astype(np.uint8)
    organs_text, media_text = dx.workloads.icrp_dat_text("AM")
    for remove in (False, True):
        dev = dx.workloads.icrp_import(raw, organs_text, media_text, remove, world=world)
        host = dx.workloads.icrp_import(raw, organs_text, media_text, remove)
        for a, b in zip(dev, host):
            assert np.array_equal(a, b) if isinstance(a, np.ndarray) else a == b
    world.close()


def test_scene_file_in_dose_file_out(dx, tmp_path):
    """a save file in OpenDXMC's HDF5 layout (written here with the library's own writer, the way
    R:src/libopendxmc/hdf5wrapper.cpp:384-459 lays it out) drives the engine head-less - dxb_load_scene builds materials
    and grid from it - and dxb_save_dose hands dose / variance / event count back in the same file format, in z-y-x
    order, after the reference's post-processing.  Where the reference's own loader is available (oracle/_ref), it reads
    the result file."""
    import subprocess
    from opendxmc_b200 import _capi as K
    lib = K.load()
    wl = dx.workloads.ctdi_body_phantom(n=24, histories=200_000, step_deg=10.0)
    # the scene file
    h = lib.dxb_h5_create()

    def put(path, a, deflate):
        a = np.ascontiguousarray(a)
        code = {np.dtype(np.float64): 1, np.dtype(np.uint64): 2, np.dtype(np.uint8): 3}[a.dtype]
        dims = (C.c_uint64 * a.ndim)(*a.shape)
        assert lib.dxb_h5_put_dataset(h, path, code, a.ndim, dims, a.ctypes.data_as(K.VP), deflate) == 0
    nx, ny, nz = wl.dim
    put(b"/dimensions", np.array(wl.dim, dtype=np.uint64), 0)
    put(b"/spacing", np.array(wl.spacing), 0)
    put(b"/densityarray", wl.density.reshape(nz, ny, nx), 1)
    put(b"/materialarray", wl.material.reshape(nz, ny, nx), 1)
    names = wl.material_names
    comps = []
    for nm in names:
        comp = dx.NISTMaterials.Composition(nm)
        comps.append("".join("%s%s" % (dx.AtomHandler.toSymbol(z), "%.6f" % w) for z, w in sorted(comp.items())))
    for path, strings in ((b"/materialnames", names), (b"/materialcomposition", comps)):
        arr = (C.c_char_p * len(strings))(*[s.encode() for s in strings])
        assert lib.dxb_h5_put_strings(h, path, len(strings), arr) == 0
    scene = tmp_path / "scene.h5"
    assert lib.dxb_h5_save(h, str(scene).encode()) == 0
    lib.dxb_h5_close(h)
    # head-less run
    ctx = K.VP()
    assert lib.dxb_create(C.byref(ctx), None, 0) == 0
    dim, sp, nm = (C.c_uint64 * 3)(), (C.c_double * 3)(), C.c_uint32()
    assert lib.dxb_load_scene(ctx, str(scene).encode(), dim, sp, C.byref(nm)) == 0, lib.dxb_last_error(ctx)
    assert list(dim) == wl.dim and list(sp) == pytest.approx(wl.spacing) and nm.value == len(names)
    lib.dxb_set_calibration_histories(ctx, 360_000)
    assert lib.dxb_run(ctx, C.byref(wl.beam.desc()), 1, 1, None) == 0, lib.dxb_last_error(ctx)
    n = wl.n_voxels
    d, v, cnt = np.zeros(n), np.zeros(n), np.zeros(n)
    units = C.create_string_buffer(4)
    assert lib.dxb_get_dose_postprocessed(ctx, 1, d.ctypes.data_as(K.c_double_p), v.ctypes.data_as(K.c_double_p), cnt.ctypes.data_as(K.c_double_p), units) == 0
    result = tmp_path / "result.h5"
    u2 = C.create_string_buffer(4)
    assert lib.dxb_save_dose(ctx, str(result).encode(), 1, u2) == 0, lib.dxb_last_error(ctx)
    assert u2.value == units.value
    lib.dxb_destroy(ctx)
    # the result file: scene + three result arrays, z-y-x, deflated
    r = K.VP()
    assert lib.dxb_h5_open(C.byref(r), str(result).encode()) == 0
    t, rk, dd, z = C.c_int(), C.c_int(), (C.c_uint64 * 8)(), C.c_int()
    for name, ref in (("dosearray", d), ("dosevariancearray", v), ("doseeventcountarray", cnt), ("densityarray", wl.density)):
        assert lib.dxb_h5_dataset_info(r, ("/" + name).encode(), C.byref(t), C.byref(rk), dd, C.byref(z)) == 0
        assert (t.value, list(dd)[:3], z.value) == (1, [nz, ny, nx], 1)
        got = np.zeros(n)
        assert lib.dxb_h5_dataset_read(r, ("/" + name).encode(), got.ctypes.data_as(K.VP), got.nbytes) == 0
        assert np.array_equal(got, ref), name
    assert d.sum() > 0 and lib.dxb_h5_dataset_string(r, b"/materialnames", 1) == names[1].encode()
    lib.dxb_h5_close(r)
    # the reference's own HDF5Wrapper::load() (compiled unmodified over the H5Cpp shim) reads the file
    exe = os.path.join(ROOT, "oracle", "_ref", "opendxmc_ref")
    if os.path.exists(exe):
        out = subprocess.run([exe, "h5load", str(result)], capture_output=True, text=True, timeout=120)
        assert out.returncode == 0, out.stdout + out.stderr
        info = json.loads(out.stdout)
        assert info["dimensions"] == wl.dim and info["materials"] == len(names)
        assert info["dose_sum"] == pytest.approx(d.sum(), rel=1e-12) and info["count_sum"] == pytest.approx(cnt.sum(), rel=1e-12)


def test_slab_local_majorants_on_the_device(dx, orc):
    """SURVEY §7 step 7 (region-local majorants), as z slabs: on the ICRP-shaped phantom the densest medium (teeth) lives in
    a few slabs, so the pool kernel's LM build needs a fraction of the tentative steps.  It must (a) follow the oracle
    on the same table draw for draw, (b) give a dose statistically equal to global tracking, (c) stay GPU-count
    invariant, (d) switch itself off where it cannot pay (C2: bone in every slab)."""
    wl = dx.workloads.icrp_phantom("AM", scale=3, histories=2_000_000)
    res = {}
    for lm in (0, -1):
        world = wl.build_world(1, [0])
        world.set_option("local_majorant", lm)
        world.set_option("dense_box", 0)   # (this test is about the slab table; the dense box has its own below)
        dx.Transport().run_transport(world, wl.beam)
        e, e2, cnt = [a.copy() for a in world.energy_scored()]
        st = world.run_stats()
        assert st["local_majorant"] == (1 if lm else 0)
        ow = orc.OracleWorld.from_workload(wl)
        if lm:
            assert mirror_local_majorant(world, ow)
            # shards of the histories sum to the whole, bit for bit, with hops in the walk too
            acc = [np.zeros_like(e), np.zeros_like(e2), np.zeros_like(cnt)]
            for rank in range(3):
                world.set_history_range(rank, 3)
                world.set_seed(SEED)
                dx.Transport().run_transport(world, wl.beam)
                for a, b in zip(acc, world.energy_scored()):
                    a += b
            assert all(np.array_equal(a, b) for a, b in zip(acc, (e, e2, cnt)))
        oe, oe2, ocnt, ost = ow.run(wl.beam, 1, SEED)
        assert_same_stream_parity(e, cnt, st, oe, ocnt, ost, f"AM scale 3 local_majorant={lm}", voxel_cm=min(wl.spacing))
        assert abs(st["hops"] - ost["hops"]) <= max(30, 1e-5 * ost["hops"])
        res[lm] = (e, e2, st)
        world.close()
    (e0, s0, st0), (e1, s1, st1) = res[0], res[-1]
    assert st1["hops"] > 0 and st0["hops"] == 0 and st1["steps"] < 0.5 * st0["steps"]
    sigma = np.sqrt(s0.sum() + s1.sum())
    assert abs(e0.sum() - e1.sum()) / sigma < 4.0
    for name, m in {"dense": wl.density > 1.2, "soft": (wl.density > 0.5) & (wl.density <= 1.2), "lung / air": wl.density <= 0.5}.items():
        s = np.sqrt(s0[m].sum() + s1[m].sum())
        assert s == 0 or abs(e0[m].sum() - e1[m].sum()) / s < 4.0, name
    c2 = dx.workloads.ct_spiral_patient(scale=8, histories=100_000, step_deg=10.0)
    world = c2.build_world(1, [0])
    n, shift, useful, table = world.local_majorant()
    assert n >= 2 and not useful and 0.0 < table.min() and table.max() <= 1.0
    dx.Transport().run_transport(world, c2.beam)
    assert world.run_stats()["local_majorant"] == 0
    world.close()


def test_dense_box_tracking_on_the_device(dx, orc):
    """SURVEY §7 step 7 (region-local majorants), as a dense box: most of a CT volume is air around the patient, and the beam
    is wider than the body.  The pool kernel's DB build crosses that air in flights and runs the quad step only inside the
    bounding box of the non-thin voxels.  It must (a) follow the oracle on the same box draw for draw - all three physics
    modes -, (b) take less than half the tentative steps of global tracking and give a statistically equal dose per tissue
    class, (c) stay GPU-count invariant (shards sum to the whole bit for bit), (d) switch itself off where there is nothing
    to skip (the CTDI phantom fills its grid) and when the option says so."""
    wl = dx.workloads.ct_spiral_patient(scale=4, histories=2_000_000, step_deg=5.0)
    res = {}
    for db in (0, -1):
        world = wl.build_world(1, [0])
        world.set_option("dense_box", db)
        dx.Transport().run_transport(world, wl.beam)
        e, e2, cnt = [a.copy() for a in world.energy_scored()]
        st = world.run_stats()
        assert st["dense_box"] == (1 if db else 0) and st["local_majorant"] == 0
        ow = orc.OracleWorld.from_workload(wl)
        mirror_local_majorant(world, ow)
        if db:
            box = world.dense_box()
            assert box["built"] and box["useful"] and 0 < box["ratio"].max() < 0.01
            n = [box["box"][a + 3] - box["box"][a] for a in range(3)]
            assert n[0] < wl.dim[0] and n[1] < wl.dim[1] and n[2] == wl.dim[2]
            acc = [np.zeros_like(e), np.zeros_like(e2), np.zeros_like(cnt)]
            for rank in range(3):
                world.set_history_range(rank, 3)
                world.set_seed(SEED)
                dx.Transport().run_transport(world, wl.beam)
                for a, b in zip(acc, world.energy_scored()):
                    a += b
            assert all(np.array_equal(a, b) for a, b in zip(acc, (e, e2, cnt)))
        oe, oe2, ocnt, ost = ow.run(wl.beam, 1, SEED)
        assert_same_stream_parity(e, cnt, st, oe, ocnt, ost, f"C2 scale 4 dense_box={db}", voxel_cm=min(wl.spacing))
        assert abs(st["hops"] - ost["hops"]) <= max(30, 1e-5 * ost["hops"])
        res[db] = (e, e2, st)
        world.close()
    (e0, s0, st0), (e1, s1, st1) = res[0], res[-1]
    assert st1["hops"] > 0 and st0["hops"] == 0 and st1["steps"] < 0.5 * st0["steps"]
    assert abs(st0["interactions"] - st1["interactions"]) / st0["interactions"] < 5e-3
    sigma = np.sqrt(s0.sum() + s1.sum())
    assert abs(e0.sum() - e1.sum()) / sigma < 4.0
    for i, name in enumerate(wl.organ_names):
        m = wl.organ == i
        s = np.sqrt(s0[m].sum() + s1[m].sum())
        assert s == 0 or abs(e0[m].sum() - e1[m].sum()) / s < 4.0, name
    # modes 0 and 2 run their own builds of the kernel
    small = dx.workloads.ct_spiral_patient(scale=8, histories=1_500_000, step_deg=10.0)
    for mode in (0, 2):
        world = small.build_world(mode, [0])
        dx.Transport().run_transport(world, small.beam)
        st = world.run_stats()
        assert st["dense_box"] == 1
        ow = orc.OracleWorld.from_workload(small)
        mirror_local_majorant(world, ow)
        oe, oe2, ocnt, ost = ow.run(small.beam, mode, SEED)
        e, e2, cnt = world.energy_scored()
        assert_same_stream_parity(e, cnt, st, oe, ocnt, ost, f"C2 scale 8 dense box mode {mode}", voxel_cm=min(small.spacing))
        world.close()
    # the CTDI phantom nearly fills its grid: the flights would cost more than the steps they save, auto mode leaves it alone
    c1 = dx.workloads.ctdi_body_phantom(n=32, histories=50_000, step_deg=10.0)
    world = c1.build_world(1, [0])
    dx.Transport().run_transport(world, c1.beam)
    box = world.dense_box()
    assert box["built"] and not box["useful"] and world.run_stats()["dense_box"] == 0
    world.close()
    # a grid without thin voxels around its content: nothing to skip, the plain build runs
    water = dx.Material.byNistName("Water, Liquid")
    world = dx.World([0])
    grid = world.addItem(dx.AAVoxelGrid(1))
    assert grid.setData([24, 24, 24], np.ones(24 ** 3), np.zeros(24 ** 3, dtype=np.uint8), [water])
    grid.setSpacing([1.0, 1.0, 1.0])
    world.build()
    beam = dx.PencilBeam([0.0, 0.0, -30.0], [0, 0, 1], 60.0)
    beam.setNumberOfExposures(4)
    beam.setNumberOfParticlesPerExposure(20_000)
    dx.Transport().run_transport(world, beam)
    assert world.run_stats()["dense_box"] == 0 and not world.dense_box()["useful"] and world.run_stats()["deposits"] > 0
    world.close()


def test_reference_driver_on_several_gpus_equals_one_gpu(tmp_path):
    """OpenDXMC's own SimulationPipeline / worker<CORRECTION>() (compiled unmodified, oracle/_ref/opendxmc_ref) with
    DXMC_B200_DEVICES=0,1,...: dxmc::World creates a multi-device context, so the reference's single worker thread drives
    all GPUs through the library-managed exchange.  Dose, variance and event count after the reference's post-processing
    must equal the one-GPU run bit for bit (spiral beam with bowtie, WED AEC, organ AEC; CT calibration)."""
    import subprocess
    from opendxmc_b200 import _capi as K
    n = K.load().dxb_device_count()
    exe = os.path.join(ROOT, "oracle", "_ref", "opendxmc_ref")
    if n < 2 or not os.path.exists(exe):
        pytest.skip("needs 2 GPUs and oracle/_ref/opendxmc_ref")
    out = {}
    for tag, devs in (("one", None), ("many", ",".join(str(i) for i in range(min(n, 4))))):
        prefix = str(tmp_path / tag)
        env = dict(os.environ)
        env.pop("DXMC_B200_DEVICES", None)
        if devs:
            env["DXMC_B200_DEVICES"] = devs
        r = subprocess.run([exe, "run", "1", "1", "20000", prefix, "5.0", "spiral"], capture_output=True, text=True, timeout=300, env=env)
        assert r.returncode == 0, r.stdout + r.stderr
        out[tag] = [np.fromfile(prefix + f".{k}.bin", dtype=np.float64) for k in ("dose", "variance", "count")]
        out[tag].append(json.load(open(prefix + ".json"))["dose_units"])
    assert out["one"][3] == out["many"][3]
    for a, b, name in zip(out["one"][:3], out["many"][:3], ("dose", "variance", "count")):
        assert np.array_equal(a, b), name
    assert out["one"][2].sum() > 0

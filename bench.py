#!/usr/bin/env python3
"""bench.py — histories/s of the dxmc::Transport hot path on the BASELINE.json workload (C2; --workload C4 for config 4).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--histories H] [--impl ours|reference] [--workload C2|C4]
                    [--inprocess] [--exchange pipelined|multicast|p2p|nccl]

A *step* is one full beam: H histories (default 1e9 = BASELINE.json configs[1]) of the 120 kV spiral CT beam with bowtie
through the synthetic 512x512x300 patient volume.  Every step uses a fresh Philox key, so nothing is cached.

  value  histories/s with the voxel grid, tables and beam already resident in HBM.  Timed region: K back-to-back beams,
         each = tally hand-over + transport kernels + [N > 1: the exchange of the fixed-point tallies] + energy -> dose.
         N > 1, default exchange "pipelined": the library's own exchange (copy-engine pulls of every rank's voxel slab,
         slab reduce -> dose) runs underneath the NEXT beam's transport kernels; the timed region ends when the last
         beam's exchange has finished on every rank.  Device time (dxb_timer_*: CUDA events on every stream of the
         context), barrier + synchronize on both sides, max over ranks.
         One process per GPU under torchrun (the contract), or --inprocess: ONE process drives the N GPUs through a
         multi-device dxb_ctx - what dxmc::Transport reaches from the reference's single worker thread.
  e2e    the same metric through the reference-facing C ABI with HOST buffers, doing what the reference's driver does
         per run (R:src/libopendxmc/simulationpipeline.cpp:145-219): dxb_set_grid (H2D of the f64 density and u8
         material arrays) + transport WITH use_beam_calibration = 1 (the nested CTDI Monte Carlo run, :165) +
         dxb_get_dose (D2H of dose, variance, event count).  `e2e.with_materials` adds dxb_set_materials (table build).
  roofline  the transport kernel: algorithmic bytes/history A = S*5 + D*48 (S, D counted by the
         kernel in the same run) x histories / CUDA-event kernel time, against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline  the CPU oracle (restated DXMClib algorithm, std::thread on all host cores; "port": the real
         DXMClib is absent from the reference tree) on a bounded sample of the same workload; at N = 1
         the deposited energy per history of a timed GPU step is checked against that sample (0.5 %).

Inputs (630 MB f64 density + 79 MB material -> 315 MB packed voxels + 2.5 GB tallies) are far larger than the
126 MB L2, so no explicit L2 flush is needed between steps.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED0 = 0x0DDC0FFEE
METRICS = {"C2": "histories/sec (512x512x300 CT, 120 kV spiral)",
           "C4": "histories/sec (512x512x400 thorax, dual-source spiral CT with AEC)"}
CALIBRATION_HISTORIES = 36_000_000  # the library's default size of the nested CTDI run


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="C2", choices=["C2", "C4"],
                    help="C2 = BASELINE.json configs[1] (the headline); C4 = configs[3]: dual-source thorax with AEC, 1e10 histories")
    ap.add_argument("--histories", type=float, default=None, help="histories per step (whole job); default 1e9 (C2) / 1e10 (C4)")
    ap.add_argument("--scale", type=int, default=1, help="grid coarsening (1 = the BASELINE volume)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--inprocess", action="store_true",
                    help="one process drives all --gpus devices through a multi-device dxb_ctx (no torchrun)")
    ap.add_argument("--exchange", default="pipelined", choices=["pipelined", "multicast", "p2p", "nccl"],
                    help="N > 1, one process per GPU: 'pipelined' = the library-managed exchange over CUDA IPC (copy-engine slab "
                         "pulls under the next beam's transport; falls back to the next choice if IPC is unavailable); 'multicast' / "
                         "'p2p' = fused reduce->dose kernel per slab reading the NVSwitch multicast sum / peer buffers (synchronous); "
                         "'nccl' = NCCL reduce to rank 0 + energy->dose there")
    args = ap.parse_args()
    if args.histories is None:
        args.histories = 1e10 if args.workload == "C4" else 1e9
    return args


class ClockSampler:
    """SM clock, power and throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe).

    NVML is used in-process (nvidia_ml_py) from a thread that is started BEFORE the warm-up steps: spawning
    `nvidia-smi -lms` at the start of the timed region attaches to every GPU of the box and stalled the 8-GPU run by
    several milliseconds per step.  Only samples taken between mark_begin() and mark_end() are reported.  Falls back
    to an `nvidia-smi` subprocess (also started early) when the NVML bindings are missing."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index, period_s=0.02):
        self.dev, self.period = device_index, period_s
        self.samples = []  # (t, sm_mhz, sm_max_mhz, power_w, reasons:set)
        self.t0 = self.t1 = None
        self._stop = threading.Event()
        self.thread = self.proc = self.nvml = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.nvml = (pynvml, h)
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self._physical_index())], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read_smi, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[self.dev])
            except (ValueError, IndexError):
                pass
        return self.dev

    def _poll_nvml(self):
        nv, h = self.nvml
        bits = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        try:
            mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        except Exception:
            mx = 0.0
        while not self._stop.is_set():
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                self.samples.append((time.perf_counter(), sm, mx, pw, {k for k, b in bits.items() if r & b}))
            except Exception:
                pass
            self._stop.wait(self.period)

    def _read_smi(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.strip().split(",")]
            if len(f) < 9:
                continue
            try:
                sm, mx, pw = float(f[1]), float(f[2]), float(f[3])
            except ValueError:
                continue
            reasons = {name for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9])
                       if v.lower().startswith("active")}
            self.samples.append((time.perf_counter(), sm, mx, pw, reasons))

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        self._stop.set()
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        if self.thread:
            self.thread.join(timeout=2)
        if self.nvml:
            try:
                self.nvml[0].nvmlShutdown()
            except Exception:
                pass
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML / nvidia-smi samples"]}
        t0, t1 = self.t0 or 0.0, self.t1 or float("inf")
        inside = [x for x in self.samples if t0 <= x[0] <= t1]
        used = inside or self.samples[-3:]
        sm = sorted(x[1] for x in used)
        reasons = set()
        for x in used:
            reasons |= x[4]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(x[2] for x in used), "reasons": sorted(reasons), "samples": len(inside),
                "power_w_max": max(x[3] for x in used), "source": "nvml" if self.nvml else "nvidia-smi"}


class SharedHostArrays:
    """dose f64, variance f64, events u64 (n voxels each) in /dev/shm, mapped and page-locked by every rank."""

    def __init__(self, rank, dist, n, torch, begin, end):
        import numpy as np
        self.ok, self.arrays, self.paths, self._torch = False, [], [], torch
        names = [None]
        if rank == 0:
            names = ["/dev/shm/dxb_bench_%d_%d" % (os.getpid(), k) for k in range(3)]
            try:
                for nm in names:
                    with open(nm, "wb") as f:
                        f.truncate(n * 8)
            except OSError:
                names = [None]
        box = [names]
        dist.broadcast_object_list(box, src=0)
        names = box[0]
        if not names or names[0] is None:
            return
        self.paths = names
        good = 1
        try:
            for nm, dt in zip(names, (np.float64, np.float64, np.uint64)):
                a = np.memmap(nm, dtype=dt, mode="r+", shape=(n,))
                a[begin:end] = 0  # first touch by the rank that will write the slab: its pages land on its socket
                self.arrays.append(a)
        except Exception:
            good = 0
        dist.barrier()  # every slab is placed before anybody pins the whole array
        try:
            for a in self.arrays:
                rc = torch.cuda.cudart().cudaHostRegister(a.ctypes.data, a.nbytes, 0)
                if int(rc) != 0:
                    good = 0
        except Exception:
            good = 0
        t = torch.tensor([good], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        self.ok = bool(int(t))

    def ptr(self, k, ctype):
        import ctypes as C
        return C.cast(self.arrays[k].ctypes.data, ctype)

    def close(self, rank):
        for a in self.arrays:
            try:
                self._torch.cuda.cudart().cudaHostUnregister(a.ctypes.data)
            except Exception:
                pass
        self.arrays = []
        if rank == 0:
            for nm in self.paths:
                try:
                    os.unlink(nm)
                except OSError:
                    pass


def make_workload(args):
    import opendxmc_b200 as dx
    if args.workload == "C4":
        return dx.workloads.ct_dual_source_thorax(scale=args.scale, histories=int(args.histories))
    return dx.workloads.ct_spiral_patient(scale=args.scale, histories=int(args.histories))


def config_dict(args, wl, extra=None):
    nvox = wl.n_voxels
    what = ("C2 synthetic CT patient %dx%dx%d, 120 kV spiral CT, bowtie, pitch 1, %d exposures" if args.workload == "C2" else
            "C4 synthetic thorax %dx%dx%d, dual-source 120/120 kV spiral CT, two bowties, WED AEC, pitch 3.2, %d exposures")
    c = {"workload": what % (wl.dim[0], wl.dim[1], wl.dim[2], wl.beam.numberOfExposures()),
         "histories_per_step": wl.beam.numberOfParticles(), "physics_mode": 1,
         "l2": "inputs larger than L2 (%.2f GB voxels + %.1f GB tallies vs 126 MB), no flush needed" % (nvox * 4 / 1e9, nvox * 32 / 1e9)}
    if extra:
        c.update(extra)
    return c


def library_build_id():
    """identity of the transport kernel compiled into the LOADED libdxmc_b200.so (hash of its sources + compiler flags,
    dxb_kernel_build_id): an ncu capture taken from another kernel build does not describe it"""
    from opendxmc_b200 import _capi as K
    return K.load().dxb_kernel_build_id().decode()


def measured_traffic(build_id):
    """DRAM bytes per history of the transport kernel from the ncu --set full capture of THIS build (profiles/*traffic*.json
    carry the build id they were captured from); None if the shipped kernel has changed since."""
    best = None
    pdir = os.path.join(ROOT, "profiles")
    try:
        names = sorted(os.listdir(pdir))
    except OSError:
        return None
    for nm in names:
        if "traffic" not in nm or not nm.endswith(".json"):
            continue
        try:
            tj = json.load(open(os.path.join(pdir, nm)))
        except Exception:
            continue
        if tj.get("library_build") == build_id and "dram_bytes_per_history" in tj:
            best = (float(tj["dram_bytes_per_history"]), nm)
    return best


# --------------------------------------------------------------------------------------- CPU oracle legs
def oracle_sample(wl, target_seconds, seed_base, with_calibration):
    """Times the CPU oracle (all host cores) on a bounded sample of the workload: the same beam with fewer histories per
    exposure through the full volume (+ the nested CTDI calibration run scaled by the same factor when asked).
    bench.py is one of the few places allowed to execute oracle/ (as the measured *baseline*, never as the product)."""
    from oracle import oracle_py as orc
    ow = orc.OracleWorld.from_workload(wl)
    nexp = wl.beam.numberOfExposures()
    full_ppe = wl.beam.numberOfParticlesPerExposure()
    ppe = max(1, 200_000 // nexp)
    st = e = None
    for attempt in range(4):
        wl.beam.setNumberOfParticlesPerExposure(ppe)
        e, _, _, st = ow.run(wl.beam, 1, seed_base + attempt, 0)
        if st["seconds"] >= 0.6 * target_seconds:
            break
        rate = st["histories"] / max(st["seconds"], 1e-9)
        ppe = max(ppe + 1, int(rate * target_seconds / nexp))
    cal_seconds, cal_hist = 0.0, 0
    if with_calibration:
        cal_hist = max(36_000, int(CALIBRATION_HISTORIES * (ppe / full_ppe)))
        t0 = time.perf_counter()
        ow.ct_calibration(wl.beam, 1, seed_base, cal_hist)
        cal_seconds = time.perf_counter() - t0
    wl.beam.setNumberOfParticlesPerExposure(full_ppe)
    return {"ow": ow, "ppe": ppe, "nexp": nexp, "stats": st, "deposited_per_history": float(e.sum()) / st["histories"],
            "calibration_seconds": cal_seconds, "calibration_histories": cal_hist}


def cpu_oracle_rate(args, wl, target_seconds):
    s = oracle_sample(wl, target_seconds, SEED0, False)
    st = s["stats"]
    return {"value": st["histories"] / st["seconds"], "unit": "histories/s", "cores": int(st["threads"]), "kind": "port",
            "sample": "%d histories (%d per exposure x %d exposures) of the same beam through the full volume, %.1f s" % (
                st["histories"], s["ppe"], s["nexp"], st["seconds"]),
            "steps_per_history": st["steps"] / st["histories"], "deposits_per_history": st["deposits"] / st["histories"],
            "deposited_kev_per_history": s["deposited_per_history"],
            "tracking": "Woodcock tracking with one global majorant, the reference's rule (the GPU arm may track the air around the "
                        "patient in flights, roofline.tracking: same expectation, fewer tentative steps)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    wl = make_workload(args)
    from oracle import oracle_py as orc
    ow = orc.OracleWorld.from_workload(wl)
    nexp = wl.beam.numberOfExposures()
    full = wl.beam.numberOfParticles()
    full_ppe = wl.beam.numberOfParticlesPerExposure()
    per_step_seconds = min(20.0, 180.0 / max(1, args.steps + args.warmup))
    ppe = max(1, 200_000 // nexp)
    for attempt in range(3):
        wl.beam.setNumberOfParticlesPerExposure(ppe)
        _, _, _, st = ow.run(wl.beam, 1, SEED0 + attempt, 0)
        if st["seconds"] >= 0.6 * per_step_seconds:
            break
        rate = st["histories"] / max(st["seconds"], 1e-9)
        ppe = max(ppe + 1, int(rate * per_step_seconds / nexp))
    wl.beam.setNumberOfParticlesPerExposure(ppe)
    for i in range(args.warmup):
        ow.run(wl.beam, 1, SEED0 + i, 0)
    hist, secs, threads = 0, 0.0, 0
    for i in range(args.steps):
        _, _, _, st = ow.run(wl.beam, 1, SEED0 + 100 + i, 0)
        hist += st["histories"]
        secs += st["seconds"]
        threads = st["threads"]
    value = hist / secs
    # e2e of the CPU arm: the same step + the nested CTDI calibration run the reference's transport() call includes
    # (R:src/libopendxmc/simulationpipeline.cpp:165, useBeamCalibration = true), scaled like the sample
    cal_hist = max(36_000, int(CALIBRATION_HISTORIES * (ppe / full_ppe)))
    t0 = time.perf_counter()
    ow.ct_calibration(wl.beam, 1, SEED0, cal_hist)
    cal_seconds = time.perf_counter() - t0
    e2e_value = hist / (secs + cal_seconds * args.steps)
    wl.beam.setNumberOfParticlesPerExposure(full_ppe)  # `config` names the workload, `cpu_baseline.sample` the bounded sample of it
    sample = "%d histories per step (%d per exposure x %d exposures) of the %d-history beam, full volume" % (ppe * nexp, ppe, nexp, full)
    out = {"impl": "reference", "metric": METRICS[args.workload], "value": value, "unit": "histories/s", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "strong",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": config_dict(args, wl, {"reference": "CPU oracle = in-repo restatement of the DXMClib algorithm (DXMClib is not in "
                                                           "/root/reference and cannot be built offline), std::thread on all host cores"}),
           "cpu_baseline": {"value": value, "unit": "histories/s", "cores": int(threads), "kind": "port", "sample": sample},
           "e2e": {"value": e2e_value, "unit": "histories/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                   "note": "transport + nested CTDI calibration run of %d histories (the library's 36e6 scaled like the sample), %.2f s per step" % (
                       cal_hist, cal_seconds)}}
    print(json.dumps(out), flush=True)
    return 0


# --------------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import ctypes as C
    import numpy as np
    import torch
    import opendxmc_b200 as dx
    from opendxmc_b200 import _capi as K
    from opendxmc_b200 import distributed as D

    rank = int(os.environ.get("RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; libdxmc_b200 has no CPU fallback")
    if args.inprocess and world_size > 1:
        raise SystemExit("bench.py: --inprocess is ONE process driving --gpus devices; do not launch it under torchrun")
    n_local = args.gpus if args.inprocess else 1   # devices behind this process's context
    n_gpus = args.gpus if args.inprocess else world_size
    if args.inprocess and torch.cuda.device_count() < n_local:
        raise SystemExit("bench.py: --inprocess --gpus %d but only %d devices visible" % (n_local, torch.cuda.device_count()))
    torch.cuda.set_device(local_rank)
    dist = None
    if world_size > 1:
        import torch.distributed as dist
        # stdout carries the ONE JSON line: NCCL's own banner ("NCCL version ...", printed when NCCL_DEBUG is set) goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        # one process per GPU: run on the CPUs next to this GPU, so that the pinned host buffers it allocates and the
        # slab of the shared result arrays it touches first live in the memory of that socket (PCIe DMA stays local)
        try:
            import pynvml
            pynvml.nvmlInit()
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(ClockSampler(local_rank)._physical_index()))
        except Exception:
            pass
    lib = K.load()
    build_id = library_build_id()

    wl = make_workload(args)
    n_hist = wl.beam.numberOfParticles()
    nvox = wl.n_voxels
    # pinned host copies of the caller-owned arrays (the reference hands in std::vector<double>/<uint8_t>)
    dens_pin = torch.empty(nvox, dtype=torch.float64).pin_memory()
    mat_pin = torch.empty(nvox, dtype=torch.uint8).pin_memory()
    dens_pin.numpy()[:] = wl.density
    mat_pin.numpy()[:] = wl.material
    wl.density, wl.material = dens_pin.numpy(), mat_pin.numpy()
    out_pin = [torch.empty(nvox, dtype=torch.float64).pin_memory(), torch.empty(nvox, dtype=torch.float64).pin_memory(),
               torch.empty(nvox, dtype=torch.int64).pin_memory()] if rank == 0 else None

    devices = list(range(n_local)) if args.inprocess else [local_rank]
    world = wl.build_world(1, devices)
    ctx = world.ctx()
    world.set_history_range(rank, world_size)
    stream = None
    if not args.inprocess:
        stream = torch.cuda.Stream(device=local_rank)
        lib.dxb_set_stream(ctx, C.c_void_p(stream.cuda_stream))
    desc = wl.beam.desc()

    # ---- N > 1, one process per GPU: the exchange step
    pipelined, fused, tally = None, None, None
    if dist is not None:
        def agreed(obj):
            ok = torch.tensor([1 if obj is not None else 0], device=f"cuda:{local_rank}")
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            return int(ok) == 1
        choice = args.exchange
        if choice == "pipelined":
            try:
                pipelined = D.PipelinedExchange(world, local_rank)
            except Exception as exc:
                print(f"bench.py: rank {rank}: library-managed exchange unavailable ({exc!r}); trying the fused multicast kernel", file=sys.stderr, flush=True)
                pipelined = None
            if not agreed(pipelined):
                if pipelined is not None:
                    lib.dxb_exchange_close(ctx)
                pipelined, choice = None, "multicast"
        if pipelined is None and choice in ("multicast", "p2p"):
            try:
                fused = D.FusedExchange(world, local_rank, multicast=(choice == "multicast"))
            except Exception as exc:  # symmetric memory unavailable on this box
                print(f"bench.py: rank {rank}: fused exchange unavailable ({exc!r}); using the NCCL reduce", file=sys.stderr, flush=True)
                fused = None
            if not agreed(fused):
                if fused is not None:
                    fused.close()
                fused = None
        if pipelined is None and fused is None:
            tally = D.tally_tensor(world, local_rank)
    sharded = pipelined if pipelined is not None else fused   # exchanges that leave the dose score distributed over the ranks
    exchange_kind = ("none (1 GPU)" if n_gpus == 1 else
                     "in-process multi-device context: copy-engine slab pulls + slab reduce->dose, pipelined under the next beam" if args.inprocess else
                     "library-managed over CUDA IPC: copy-engine slab pulls + slab reduce->dose, pipelined under the next beam" if pipelined is not None else
                     ("fused reduce->dose per slab, " + fused.kind) if fused is not None else "NCCL int64 reduce to rank 0 + energy->dose")

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        lib.dxb_flush(ctx)

    shared_out = None
    if sharded is not None and not args.no_e2e:
        shared_out = SharedHostArrays(rank, dist, nvox, torch, sharded.begin, sharded.end)
        if not shared_out.ok:
            shared_out = None

    def transport_and_finish(i, use_calibration, timed_stats=None):
        """one beam with everything resident: tallies -> exchange -> dose."""
        lib.dxb_set_seed(ctx, SEED0 + 7919 * (i + 1))
        rc = lib.dxb_run_transport(ctx, C.byref(desc), 1, None)
        assert rc == 0, lib.dxb_last_error(ctx)
        if timed_stats is not None:
            timed_stats.append(world.run_stats())
        if pipelined is not None:
            pipelined.finish_beam(wl.beam, 1, use_calibration)   # barrier over the ranks + enqueue (asynchronous)
            return
        if fused is not None:
            with torch.cuda.stream(stream):
                fused.finish_beam(wl.beam, 1, use_calibration)
            return
        if dist is not None:
            with torch.cuda.stream(stream):
                D.reduce_tallies(tally, 0)
            stream.synchronize()
        if rank == 0 or dist is None:
            rc = lib.dxb_finish_beam(ctx, C.byref(desc), 1, 1 if use_calibration else 0, None)
            assert rc == 0, lib.dxb_last_error(ctx)

    def step(i, timed_stats=None):
        transport_and_finish(i, False, timed_stats)

    e2e_cal_ms = []
    e2e_parts = {"set_materials": 0.0, "set_grid": 0.0, "transport": 0.0, "finish_beam": 0.0, "get_dose": 0.0}

    def timed(name, fn):
        t = time.perf_counter()
        r = fn()
        e2e_parts[name] += time.perf_counter() - t
        return r

    def e2e_step(i, with_materials=False):
        """the reference-facing call sequence with host buffers: [materials ->] setData/build -> transport(useBeamCalibration = true) -> read dose."""
        t_mark = time.perf_counter()
        if with_materials:
            g = world._item
            mats = (K.VP * len(g._materials))(*[m._h for m in g._materials])
            rc = lib.dxb_set_materials(ctx, len(g._materials), mats)
            assert rc == 0, lib.dxb_last_error(ctx)
            e2e_parts["set_materials"] += time.perf_counter() - t_mark
            t_mark = time.perf_counter()
        if sharded is not None:
            # one process per GPU: every rank uploads its slab, the packed slabs travel over NVLink
            D.set_grid_sharded(world, wl.dim, wl.spacing, wl.density, wl.material, local_rank, stream=stream)
        else:
            dim = (C.c_uint64 * 3)(*wl.dim)
            sp = (C.c_double * 3)(*wl.spacing)
            rc = lib.dxb_set_grid(ctx, dim, sp, wl.density.ctypes.data_as(K.c_double_p), wl.material.ctypes.data_as(K.c_u8_p))
            assert rc == 0, lib.dxb_last_error(ctx)
        e2e_parts["set_grid"] += time.perf_counter() - t_mark
        t_mark = time.perf_counter()
        transport_and_finish(1000 + i, True)
        e2e_parts["transport"] += time.perf_counter() - t_mark   # transport + (barrier) + calibration + finish / exchange enqueue
        t_mark = time.perf_counter()
        e2e_cal_ms.append(world.run_stats()["calibration_ms"])
        if shared_out is not None:
            rc = lib.dxb_get_dose_range(ctx, sharded.begin, sharded.end, shared_out.ptr(0, K.c_double_p), shared_out.ptr(1, K.c_double_p),
                                        shared_out.ptr(2, K.c_u64_p))
            assert rc == 0, lib.dxb_last_error(ctx)
            e2e_parts["get_dose"] += time.perf_counter() - t_mark
            return
        if fused is not None:
            with torch.cuda.stream(stream):
                fused.gather_dose(0)
        if rank == 0:
            rc = lib.dxb_get_dose(ctx, C.cast(out_pin[0].data_ptr(), K.c_double_p), C.cast(out_pin[1].data_ptr(), K.c_double_p),
                                  C.cast(out_pin[2].data_ptr(), K.c_u64_p))
            assert rc == 0, lib.dxb_last_error(ctx)
        e2e_parts["get_dose"] += time.perf_counter() - t_mark

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()  # before the warm-up: attaching to the GPUs is not free
    for i in range(args.warmup):
        step(i)
    stats = []
    barrier()
    sampler.mark_begin()
    t0 = time.perf_counter()
    rc = lib.dxb_timer_begin(ctx)
    assert rc == 0, lib.dxb_last_error(ctx)
    for i in range(args.steps):
        step(args.warmup + i, stats)
    tm = C.c_double()
    rc = lib.dxb_timer_end(ctx, C.byref(tm))   # after every stream of the context, the pending exchange included
    assert rc == 0, lib.dxb_last_error(ctx)
    barrier()
    wall = time.perf_counter() - t0
    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None
    xt = (C.c_double * 3)()
    lib.dxb_exchange_times(ctx, xt)
    tms = torch.tensor([tm.value, wall * 1e3, xt[0], xt[1], xt[2]], dtype=torch.float64, device=f"cuda:{local_rank}")
    if dist is not None:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_total = float(tms[0])
    # per-rank kernel statistics -> global S, D and the slowest rank's kernel time
    k = torch.tensor([sum(s["histories"] for s in stats), sum(s["steps"] for s in stats), sum(s["deposits"] for s in stats),
                      sum(s["kernel_launches"] for s in stats)], dtype=torch.float64, device=f"cuda:{local_rank}")
    kms = torch.tensor([sum(s["transport_ms"] for s in stats)], dtype=torch.float64, device=f"cuda:{local_rank}")
    if dist is not None:
        dist.all_reduce(k, op=dist.ReduceOp.SUM)
        dist.all_reduce(kms, op=dist.ReduceOp.MAX)
    hist_done, steps_done, deps_done, launches = [float(v) for v in k]
    kernel_ms = float(kms[0])

    # ---- N = 1: the deposited energy per history of the last timed beam (read straight from its tallies)
    dep_per_hist = None
    if n_gpus == 1:
        e = np.zeros(nvox)
        rc = lib.dxb_get_energy_scored(ctx, e.ctypes.data_as(K.c_double_p), None, None)
        assert rc == 0, lib.dxb_last_error(ctx)
        dep_per_hist = float(e.sum()) / stats[-1]["histories"]
        del e

    # ---- e2e
    e2e = None
    if not args.no_e2e:
        e2e_step(-1)  # warm (builds the calibration phantom once per context, like the first beam of a session)
        barrier()
        del e2e_cal_ms[:]
        for k in e2e_parts:
            e2e_parts[k] = 0.0
        t0 = time.perf_counter()
        n_e2e = max(1, min(args.steps, 2))
        for i in range(n_e2e):
            e2e_step(i)
        barrier()
        te = torch.tensor([time.perf_counter() - t0, sum(e2e_cal_ms) / max(len(e2e_cal_ms), 1)], dtype=torch.float64, device=f"cuda:{local_rank}")
        parts_ms = {k: 1e3 * v / n_e2e for k, v in e2e_parts.items() if k != "set_materials"}
        e2e_step(100, with_materials=True)  # warm
        barrier()
        t0 = time.perf_counter()
        e2e_step(101, with_materials=True)
        barrier()
        tmat = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=f"cuda:{local_rank}")
        if dist is not None:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
            dist.all_reduce(tmat, op=dist.ReduceOp.MAX)
        per_rank_upload = sharded is not None or args.inprocess
        e2e = {"value": n_hist * n_e2e / float(te[0]), "unit": "histories/s",
               "h2d_bytes_per_step": int(nvox * 9 * (1 if per_rank_upload else n_gpus)), "d2h_bytes_per_step": int(nvox * 24),
               "steps": n_e2e, "ms_per_step": 1e3 * float(te[0]) / n_e2e,
               "calibration_ms": float(te[1]),
               "parts_ms": dict(parts_ms, note="host clock of rank 0 per step: set_grid = upload + pack + (all-gather) + majorant; transport = "
                                               "dxb_run_transport + (barrier) + nested calibration run + dxb_finish_beam (with the pipelined exchange: "
                                               "enqueue only); get_dose = wait for the exchange + D2H of this rank's slab"),
               "calibration": "use_beam_calibration = 1: nested CTDI run of %d histories inside every timed step (device time reported as calibration_ms)" % CALIBRATION_HISTORIES,
               "with_materials": {"value": n_hist / float(tmat[0]), "ms_per_step": 1e3 * float(tmat[0]),
                                  "note": "the same step preceded by dxb_set_materials (host table build + upload), 1 timed step"},
               "note": ("host clock around: dxb_set_grid%s + dxb_run_transport + %sdxb_finish_beam(use_beam_calibration = 1) + %s" % (
                   "_sharded (every rank uploads its slab, packed slabs all-gathered over NVLink)" if sharded is not None else
                   " (every device uploads its slab, packed slabs all-gathered by peer copies)" if args.inprocess and n_gpus > 1 else "",
                   "barrier + " if dist is not None else "",
                   "dxb_get_dose_range of every rank's slab into pinned host arrays shared by the processes" if shared_out is not None else
                   "dxb_get_dose (every device copies its slab over its own PCIe link)" if args.inprocess and n_gpus > 1 else
                   "(N>1: slab gather to rank 0) + dxb_get_dose, pinned host buffers"))}
        if shared_out is not None and rank == 0:
            # the caller's view: every voxel of the three arrays was written by exactly one rank
            e2e["events_read_back"] = int(shared_out.arrays[2].sum())
        elif rank == 0:
            e2e["events_read_back"] = int(out_pin[2].sum())

    if shared_out is not None:
        shared_out.close(rank)
    if pipelined is not None:
        pipelined.close()
    if fused is not None:
        fused.close()
    if rank != 0:
        world.close()
        if dist is not None:
            dist.destroy_process_group()
        return 0

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    S = steps_done / hist_done
    Dd = deps_done / hist_done
    a_hist = S * 5.0 + Dd * 48.0
    # per-GPU roofline of the transport kernel: this GPU's share of the histories over the slowest GPU's kernel time
    achieved = (hist_done / n_gpus) * a_hist / (kernel_ms * 1e-3) / 1e9
    # measured DRAM traffic of the kernel: one `ncu --set full` capture of THIS library build (profiles/*traffic*.json), per launch
    tr = measured_traffic(build_id)
    traffic_per_hist = tr[0] if tr else None
    traffic = traffic_per_hist * hist_done / max(launches, 1.0) if tr else None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_unit": "DRAM bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum per history x histories per launch)",
                "traffic_bytes_per_history": traffic_per_hist, "traffic_build": build_id if tr else None, "traffic_source": tr[1] if tr else
                "none: no profiles/*traffic*.json was captured from the loaded library build " + build_id,
                "library_build": build_id,
                "algorithmic_bytes_per_launch": a_hist * hist_done / max(launches, 1.0),
                "kernel": "transportKernelPool<1,false,true,16,0%s>" % (",false,false,true" if stats[-1].get("dense_box") else
                                                                          ",true" if stats[-1].get("local_majorant") else ""),
                "tracking": ("dense box: flights through the air around the patient, quad step at the global majorant inside"
                             if stats[-1].get("dense_box") else "slab-local majorants" if stats[-1].get("local_majorant") else "global majorant"),
                "flights_per_history": sum(s.get("hops", 0) for s in stats) / max(sum(s["histories"] for s in stats), 1),
                "peak_source": "measured" if peaks else "fallback",
                "algorithmic_bytes_per_history": a_hist, "steps_per_history": S, "deposits_per_history": Dd,
                "sector_bytes_per_history": (S + 3 * Dd) * 32.0, "sector_frac": (hist_done / n_gpus) * (S + 3 * Dd) * 32.0 / (kernel_ms * 1e-3) / 1e9 / peak,
                "kernel_ms_per_step": kernel_ms / args.steps, "kernel_share_of_step": kernel_ms / ms_total}
    value = n_hist * args.steps / (ms_total * 1e-3)
    finish_kernels = args.steps * (n_gpus if (sharded is not None or args.inprocess) else 1)
    out = {"metric": METRICS[args.workload], "value": value, "unit": "histories/s", "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic",
           "config": config_dict(args, wl, {"parallelism": "histories sharded over %d GPU(s), %s" % (
               n_gpus, "one process drives all of them (multi-device dxb_ctx)" if args.inprocess else "one process per GPU"),
               "exchange": exchange_kind, "calibration_in_value": "off (device-resident beams); e2e runs with use_beam_calibration = 1"}),
           "clocks": clocks, "gpu_launches": int(launches + finish_kernels),  # transport kernels (all GPUs) + energy->dose / slab reduce kernels
           "roofline": roofline}
    if n_gpus > 1:
        out["exchange_ms"] = {"pulls": float(tms[2]), "slab_reduce_to_dose": float(tms[3]), "clear": float(tms[4]),
                              "non_kernel_ms_per_step": (ms_total - kernel_ms) / args.steps,
                              "note": "device time of the LAST beam's exchange (max over ranks); with the pipelined exchange the earlier ones ran "
                                      "underneath the next beam's transport kernels - non_kernel_ms_per_step is what stayed exposed per beam "
                                      "(barrier, tail of the slab reduce, the last exchange / K)"}
    if e2e:
        out["e2e"] = e2e
    world.close()
    if not args.no_cpu_baseline and n_gpus == 1 and world_size == 1:  # the CPU baseline is reported at N = 1 only
        cb = cpu_oracle_rate(args, wl, args.cpu_seconds)
        out["cpu_baseline"] = cb
        # the timed GPU step is a correct one: same physics, same tables, independent histories
        rel = abs(dep_per_hist - cb["deposited_kev_per_history"]) / cb["deposited_kev_per_history"]
        out["dose_check"] = {"gpu_deposited_kev_per_history": dep_per_hist, "oracle_deposited_kev_per_history": cb["deposited_kev_per_history"],
                             "rel": rel, "tolerance": 5e-3, "ok": bool(rel <= 5e-3)}
        assert rel <= 5e-3, "the timed GPU step deposits %.6g keV/history, the oracle sample %.6g (> 0.5 %%)" % (
            dep_per_hist, cb["deposited_kev_per_history"])
    print(json.dumps(out), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())

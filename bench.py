#!/usr/bin/env python3
"""bench.py — histories/s of the dxmc::Transport hot path on the BASELINE.json workload (C2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--histories H] [--impl ours|reference]

A *step* is one full beam: H histories (default 1e9 = BASELINE.json configs[1]) of the 120 kV spiral CT beam with
bowtie through the synthetic 512x512x300 patient volume.  Every step uses a fresh Philox key, so nothing is cached.

  value  histories/s with the voxel grid, tables and beam already resident in HBM (timed region: tally clear +
         transport kernels + [N>1: NCCL reduce of the fixed-point tallies to rank 0] + energy->dose), CUDA events,
         barrier + synchronize on both sides, max over ranks.
  e2e    the same metric through the reference-facing C ABI with HOST buffers: dxb_set_grid (H2D of the f64 density
         and u8 material arrays from pinned memory) + dxb_run_transport/dxb_finish_beam + dxb_get_dose (D2H of dose,
         variance, event count), i.e. what R:src/libopendxmc/simulationpipeline.cpp:145-219 does around transport().
  roofline  the transport kernel: algorithmic bytes/history A = S*5 + D*48 (S, D counted by the
         kernel in the same run) x histories / CUDA-event kernel time, against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline  the CPU oracle (restated DXMClib algorithm, std::thread on all host cores; "port": the real
         DXMClib is absent from the reference tree) on a bounded sample of the same workload.

Inputs (630 MB f64 density + 79 MB material -> 629 MB packed voxels + 2.5 GB tallies) are far larger than the
126 MB L2, so no explicit L2 flush is needed between steps.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED0 = 0x0DDC0FFEE
METRIC = "histories/sec (512x512x300 CT, 120 kV spiral)"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--histories", type=float, default=1e9, help="histories per step (whole job)")
    ap.add_argument("--scale", type=int, default=1, help="grid coarsening (1 = the BASELINE 512x512x300 volume)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--exchange", default="multicast", choices=["multicast", "p2p", "nccl"],
                    help="N > 1: fused reduce->dose per voxel slab, reading the sum over ranks through the NVSwitch multicast "
                         "address (default; falls back to P2P pulls without multicast support) or by NVLink P2P pulls, or an "
                         "NCCL reduce to rank 0 followed by energy->dose there")
    return ap.parse_args()


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
    return dx.workloads.ct_spiral_patient(scale=args.scale, histories=int(args.histories))


def config_dict(args, wl, extra=None):
    c = {"workload": "C2 synthetic CT patient %dx%dx%d, 120 kV spiral CT, bowtie, pitch 1, %d exposures" % (
        wl.dim[0], wl.dim[1], wl.dim[2], wl.beam.numberOfExposures()),
        "histories_per_step": wl.beam.numberOfParticles(), "physics_mode": 1,
        "l2": "inputs larger than L2 (0.63 GB voxels + 2.5 GB tallies vs 126 MB), no flush needed"}
    if extra:
        c.update(extra)
    return c


# --------------------------------------------------------------------------------------- CPU oracle legs
def cpu_oracle_rate(args, wl, target_seconds):
    """Times the CPU oracle (all host cores) on a bounded sample of the workload.  bench.py is one of the few places
    allowed to execute oracle/ (as the measured *baseline*, never as the product)."""
    from oracle import oracle_py as orc
    import opendxmc_b200 as dx
    ow = orc.OracleWorld.from_workload(wl)
    nexp = wl.beam.numberOfExposures()
    full_ppe = wl.beam.numberOfParticlesPerExposure()
    # size the sample so that it takes about target_seconds: probe, then rescale until the run is long enough
    ppe = max(1, 200_000 // nexp)
    st = None
    for attempt in range(4):
        wl.beam.setNumberOfParticlesPerExposure(ppe)
        _, _, _, st = ow.run(wl.beam, 1, SEED0 + attempt, 0)
        if st["seconds"] >= 0.6 * target_seconds:
            break
        rate = st["histories"] / max(st["seconds"], 1e-9)
        ppe = max(ppe + 1, int(rate * target_seconds / nexp))
    wl.beam.setNumberOfParticlesPerExposure(full_ppe)
    return {"value": st["histories"] / st["seconds"], "unit": "histories/s", "cores": int(st["threads"]), "kind": "port",
            "sample": "%d histories (%d per exposure x %d exposures) of the same beam through the full volume, %.1f s" % (
                st["histories"], ppe, nexp, st["seconds"]),
            "steps_per_history": st["steps"] / st["histories"], "deposits_per_history": st["deposits"] / st["histories"]}, ow


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    wl = make_workload(args)
    from oracle import oracle_py as orc
    ow = orc.OracleWorld.from_workload(wl)
    nexp = wl.beam.numberOfExposures()
    full = wl.beam.numberOfParticles()
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
    sample = "%d histories per step (%d per exposure x %d exposures) of the %d-history beam, full volume" % (ppe * nexp, ppe, nexp, full)
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": "histories/s", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "strong",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": config_dict(args, wl, {"reference": "CPU oracle = in-repo restatement of the DXMClib algorithm (DXMClib is not in "
                                                           "/root/reference and cannot be built offline), std::thread on all host cores"}),
           "cpu_baseline": {"value": value, "unit": "histories/s", "cores": int(threads), "kind": "port", "sample": sample},
           "e2e": {"value": value, "unit": "histories/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
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

    world = wl.build_world(1, [local_rank])
    ctx = world.ctx()
    world.set_history_range(rank, world_size)
    stream = torch.cuda.Stream(device=local_rank)
    K.load().dxb_set_stream(ctx, C.c_void_p(stream.cuda_stream))
    tr = dx.Transport()
    desc = wl.beam.desc()
    # N > 1: the exchange step.  Preferred: tallies in symmetric memory + the fused reduce->dose kernel (NVSwitch
    # multicast sum, else P2P pull); fallback: one NCCL reduce of the tally buffer to rank 0 + energy->dose there.
    exchange, tally = None, None
    if dist is not None and args.exchange != "nccl":
        try:
            exchange = D.FusedExchange(world, local_rank, multicast=(args.exchange == "multicast"))
        except Exception as exc:  # symmetric memory unavailable on this box
            if rank == 0:
                print(f"bench.py: fused exchange unavailable ({exc!r}); using the NCCL reduce", file=sys.stderr, flush=True)
            exchange = None
        ok = torch.tensor([1 if exchange is not None else 0], device=f"cuda:{local_rank}")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok) == 0 and exchange is not None:
            exchange.close()
            exchange = None
    if exchange is None:
        tally = D.tally_tensor(world, local_rank)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # N > 1 with the fused exchange: the dose score stays distributed (rank r holds its slab), so the e2e read-out is
    # sharded too - every rank copies ITS slab over its own PCIe link into result arrays that rank 0 (the caller) placed
    # in host memory shared by the processes (dxb_get_dose_range), instead of gathering everything through rank 0.
    shared_out = None
    if exchange is not None and not args.no_e2e:
        shared_out = SharedHostArrays(rank, dist, nvox, torch, exchange.begin, exchange.end)
        if not shared_out.ok:
            shared_out = None

    def step(i, timed_stats=None):
        """one beam with everything resident: tallies -> (reduce) -> dose."""
        lib.dxb_set_seed(ctx, SEED0 + 7919 * (i + 1))
        rc = lib.dxb_run_transport(ctx, C.byref(desc), 1, None)
        assert rc == 0, lib.dxb_last_error(ctx)
        if timed_stats is not None:
            timed_stats.append(world.run_stats())
        if exchange is not None:
            with torch.cuda.stream(stream):
                exchange.finish_beam(wl.beam, 1, False)
            return
        if dist is not None:
            with torch.cuda.stream(stream):
                D.reduce_tallies(tally, 0)
            stream.synchronize()
        if rank == 0:
            rc = lib.dxb_finish_beam(ctx, C.byref(desc), 1, 0, None)
            assert rc == 0, lib.dxb_last_error(ctx)

    def e2e_step(i):
        """the reference-facing call sequence with host buffers: setData/build -> transport -> read dose."""
        if exchange is not None:
            # one process per GPU: every rank uploads its slab, the packed slabs travel over NVLink
            D.set_grid_sharded(world, wl.dim, wl.spacing, wl.density, wl.material, local_rank, stream=stream)
        else:
            dim = (C.c_uint64 * 3)(*wl.dim)
            sp = (C.c_double * 3)(*wl.spacing)
            rc = lib.dxb_set_grid(ctx, dim, sp, wl.density.ctypes.data_as(K.c_double_p), wl.material.ctypes.data_as(K.c_u8_p))
            assert rc == 0, lib.dxb_last_error(ctx)
        step(1000 + i)
        if shared_out is not None:
            rc = lib.dxb_get_dose_range(ctx, exchange.begin, exchange.end, shared_out.ptr(0, K.c_double_p), shared_out.ptr(1, K.c_double_p),
                                        shared_out.ptr(2, K.c_u64_p))
            assert rc == 0, lib.dxb_last_error(ctx)
            return
        if exchange is not None:
            with torch.cuda.stream(stream):
                exchange.gather_dose(0)
        if rank == 0:
            rc = lib.dxb_get_dose(ctx, C.cast(out_pin[0].data_ptr(), K.c_double_p), C.cast(out_pin[1].data_ptr(), K.c_double_p),
                                  C.cast(out_pin[2].data_ptr(), K.c_u64_p))
            assert rc == 0, lib.dxb_last_error(ctx)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()  # before the warm-up: attaching to the GPUs is not free
    for i in range(args.warmup):
        step(i)
    stats = []
    barrier()
    sampler.mark_begin()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record(stream)
    for i in range(args.steps):
        step(args.warmup + i, stats)
    ev1.record(stream)
    barrier()
    wall = time.perf_counter() - t0
    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None
    ms = ev0.elapsed_time(ev1)
    tms = torch.tensor([ms, wall * 1e3], dtype=torch.float64, device=f"cuda:{local_rank}")
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

    # ---- e2e
    e2e = None
    if not args.no_e2e:
        e2e_step(-1)  # warm
        barrier()
        t0 = time.perf_counter()
        n_e2e = max(1, min(args.steps, 2))
        for i in range(n_e2e):
            e2e_step(i)
        barrier()
        te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=f"cuda:{local_rank}")
        if dist is not None:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": n_hist * n_e2e / float(te[0]), "unit": "histories/s",
               "h2d_bytes_per_step": int(nvox * 9 * (1 if exchange is not None else world_size)), "d2h_bytes_per_step": int(nvox * 24),
               "steps": n_e2e, "ms_per_step": 1e3 * float(te[0]) / n_e2e,
               "note": ("host-clock around dxb_set_grid_sharded (every rank uploads its slab, packed slabs all-gathered over NVLink) + "
                        "dxb_run_transport + fused exchange/finish + dxb_get_dose_range of every rank's slab into "
                        "pinned host arrays shared by the processes" if shared_out is not None else
                        "host-clock around dxb_set_grid + dxb_run_transport + exchange/finish + (N>1: slab gather to rank 0) + dxb_get_dose, pinned host buffers")}
        if shared_out is not None and rank == 0:
            # the caller's view: every voxel of the three arrays was written by exactly one rank
            e2e["events_read_back"] = int(shared_out.arrays[2].sum())

    if shared_out is not None:
        shared_out.close(rank)
    if rank != 0:
        if exchange is not None:
            exchange.close()
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
    D = deps_done / hist_done
    a_hist = S * 5.0 + D * 48.0
    # per-GPU roofline of the transport kernel: this rank's share of the histories over the slowest rank's kernel time
    achieved = (hist_done / world_size) * a_hist / (kernel_ms * 1e-3) / 1e9
    # measured DRAM traffic of the kernel (one `ncu --set full` capture, profiles/r01_pool_traffic.json), per launch
    traffic, traffic_per_hist = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01_pool_traffic.json")))
        traffic_per_hist = float(tj["dram_bytes_per_history"])
        traffic = traffic_per_hist * hist_done / max(launches, 1.0)
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_unit": "DRAM bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum per history x histories per launch)",
                "traffic_bytes_per_history": traffic_per_hist, "algorithmic_bytes_per_launch": a_hist * hist_done / max(launches, 1.0),
                "kernel": "transportKernelPool<1,false,true,16>", "peak_source": "measured" if peaks else "fallback",
                "algorithmic_bytes_per_history": a_hist, "steps_per_history": S, "deposits_per_history": D,
                "sector_bytes_per_history": (S + 3 * D) * 32.0,
                "kernel_ms_per_step": kernel_ms / args.steps, "kernel_share_of_step": kernel_ms / ms_total}
    value = n_hist * args.steps / (ms_total * 1e-3)
    out = {"metric": METRIC, "value": value, "unit": "histories/s", "n_gpus": world_size, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic", "config": config_dict(args, wl, {"parallelism": "histories sharded over %d GPU(s)" % world_size,
                                                                 "exchange": ("none (1 GPU)" if dist is None else
                                                                              ("fused reduce->dose per slab, " + exchange.kind) if exchange is not None
                                                                              else "NCCL int64 reduce to rank 0 + energy->dose")}),
           "clocks": clocks, "gpu_launches": int(launches + args.steps * (world_size if exchange is not None else 1)),  # transport kernels (all ranks) + energy->dose (rank 0, or one fused slab kernel per rank)
           "roofline": roofline}
    if e2e:
        out["e2e"] = e2e
    if exchange is not None:
        exchange.close()
    world.close()
    if not args.no_cpu_baseline and world_size == 1:  # the CPU baseline is reported at N = 1 only
        cb, _ = cpu_oracle_rate(args, wl, args.cpu_seconds)
        out["cpu_baseline"] = cb
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

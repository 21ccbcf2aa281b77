#!/usr/bin/env python3
"""Instruction / stall shares of code regions of the transport kernels from an ncu capture (source page, cuda+sass).

Region boundaries are found by searching the source files for marker strings, so they follow the source.
Usage: ncu_buckets.py file.ncu-rep [histories]   (histories -> warp instructions per history)
"""
import csv, io, os, subprocess, sys

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "opendxmc_b200", "csrc")
MARKERS = {
    "transport_common.cuh": [
        ("struct PhiloxBlock", "philox"), ("struct TabPos", "tabpos/lerp"), ("float exitDistance(", "exitDistance"),
        ("void deflect(", "deflect"), ("void scoreEnergy(", "score"), ("bool comptonTry(", "comptonTry"),
        ("bool rayleighTry(", "rayleighTry"), ("unsigned int voxelIndex(", "voxelIndex"),
        ("bool dopplerBroaden(", "mode-2 samplers"), ("void isotropic(", "isotropic"), ("bool sampleSource(", "source sampling")],
    "transport.cu": [
        ("// ------------------------------------------------------------------ the history kernel", "kernel prologue"),
        ("auto finishScatter", "finishScatter"), ("    for (;;) {", "vote/policy"), ("        if (phase == 0) {", "step phase"),
        ("        } else if (phase == 1) {", "interact phase"), ("        } else if (phase == 3) {", "rayleigh phase"),
        ("            // ------------------------------------------------------------ refill", "refill/source"),
        ("    // ---------------- statistics", "stats"),
        ("// ------------------------------------------------------------------ grid preparation kernels", "other")],
    "transport_mux.cu": [
        ("__global__ void __launch_bounds__", "kernel prologue"), ("    for (;;) {", "vote/policy"),
        ("        if (phase == kPhStep) {", "step: load state"), ("            bool stepping = active;", "step: pairs"),
        ("            if (active) { // step: store", "step: store state"),
        ("        } else if (phase == kPhInt || phase == kPhRay) {", "interact: load state"),
        ("// interact: channel", "interact: channel"), ("// interact: sample", "interact: sample"),
        ("// interact: scatter", "interact: scatter"), ("// interact: store", "interact: store/score"),
        ("            // ------------------------------------------------------------ refill", "refill: source sampling"),
        ("            // lanes with a dead slot pop", "refill: pop"),
        ("    // ---------------- statistics", "stats")],
    "transport_pool.cu": [
        ("__device__ __forceinline__ int phaseWord", "claim/publish"), ("__global__ void __launch_bounds__", "kernel prologue"),
        ("    for (;;) {", "vote/policy"),
        ("        if (phase == kPhStep) {", "step: claim+load"), ("            bool stepping = active;", "step: pairs"),
        ("            if (active) {\n                if (newPhase != kPhDead) {\n                    sp[kWPx", "step: store"),
        ("        } else if (phase == kPhInt || phase == kPhRay) {", "interact"),
        ("            // ------------------------------------------------------------ refill", "refill/source"),
        ("    // ---------------- statistics", "stats")],
    "device_types.cuh": [("__host__ __device__ inline unsigned int quantizeDensityBits", "voxel unpack")],
}


def load_bounds():
    out = {}
    for fname, markers in MARKERS.items():
        path = os.path.join(CSRC, fname)
        if not os.path.exists(path):
            continue
        src = open(path).read().split("\n")
        b = []
        for marker, name in markers:
            ln = next((i + 1 for i, l in enumerate(src) if marker in l), None)
            if ln is not None:
                b.append((ln, name))
        b.sort()
        out[fname] = b
    return out


def region_of(bounds, fname, ln):
    b = bounds.get(fname)
    if not b:
        return fname
    name = fname + " (head)"
    for start, n in b:
        if ln >= start:
            name = n
    return name


def main():
    rep = sys.argv[1]
    nhist = float(sys.argv[2]) if len(sys.argv) > 2 else None
    bounds = load_bounds()
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    agg = {}
    fname = None
    iI = iT = iS = None
    num = lambda v: int(v) if v.strip().isdigit() else 0
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = os.path.basename(r[1])
            continue
        if r[0] == "Line No":
            iI, iT, iS = r.index("Instructions Executed"), r.index("Thread Instructions Executed"), r.index("# Samples")
            continue
        if iI is None or len(r) <= iT or not r[0].strip().isdigit():
            continue
        reg = region_of(bounds, fname, int(r[0]))
        a = agg.setdefault(reg, [0, 0, 0])
        a[0] += num(r[iI])
        a[1] += num(r[iT])
        a[2] += num(r[iS])
    ti = sum(a[0] for a in agg.values())
    tt = sum(a[1] for a in agg.values())
    ts = sum(a[2] for a in agg.values())
    print(f"total warp instr {ti:.4e}  thread instr {tt:.4e}  avg lanes {tt / ti:.2f}" + (f"  warp-inst/history {ti / nhist:.1f}" if nhist else ""))
    for reg, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        line = f"{reg:26s} inst {100 * a[0] / ti:5.1f}%  stall {100 * a[2] / max(ts, 1):5.1f}%  lanes {a[1] / max(a[0], 1):5.1f}"
        if nhist:
            line += f"  warp-inst/history {a[0] / nhist:6.1f}"
        print(line)


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Instruction / stall shares of code regions of transport.cu from an ncu capture (source page, cuda+sass).
Region boundaries are found by searching transport.cu for marker strings, so they follow the source."""
import csv, io, subprocess, sys, os, re

SRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "opendxmc_b200", "csrc", "transport.cu")
MARKERS = [("struct PhiloxBlock", "philox"), ("struct TabPos", "tabpos/lerp"), ("float exitDistance(", "exitDistance"),
           ("void deflect(", "deflect"), ("void scoreEnergy(", "score"), ("bool comptonTry(", "comptonTry"),
           ("bool rayleighTry(", "rayleighTry"), ("// ------------------------------------------------------------------ the history kernel", "kernel prologue"),
           ("auto finishScatter", "finishScatter"), ("    for (;;) {", "vote/policy"), ("        if (phase == 0) {", "step phase"),
           ("        } else if (phase == 1) {", "interact phase"), ("        } else if (phase == 3) {", "rayleigh phase"),
           ("            // ------------------------------------------------------------ refill", "refill/source"),
           ("    // ---------------- statistics", "stats"), ("// ------------------------------------------------------------------ grid preparation kernels", "other")]


def main():
    rep = sys.argv[1]
    nhist = float(sys.argv[2]) if len(sys.argv) > 2 else None
    src = open(SRC).read().split("\n")
    bounds = []
    for marker, name in MARKERS:
        ln = next((i + 1 for i, l in enumerate(src) if l.startswith(marker) or marker in l and marker.startswith("//")), None)
        if ln is None:
            ln = next(i + 1 for i, l in enumerate(src) if marker in l)
        bounds.append((ln, name))
    bounds.sort()
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
    hdr = rows[hi]
    iI, iT, iS = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
    num = lambda v: int(v) if v.strip().isdigit() else 0
    agg = {}
    for r in rows[hi + 1:]:
        if len(r) <= iT or not r[0].strip().isdigit():
            continue
        ln = int(r[0])
        name = "other"
        for b, n in bounds:
            if ln >= b:
                name = n
        a = agg.setdefault(name, [0, 0, 0])
        a[0] += num(r[iI]); a[1] += num(r[iT]); a[2] += num(r[iS])
    ti = sum(a[0] for a in agg.values()); ts = sum(a[2] for a in agg.values()); tt = sum(a[1] for a in agg.values())
    print(f"total warp instr {ti:.4e}  thread instr {tt:.4e}  avg lanes {tt/ti:.2f}")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        extra = f"  warp-inst/history {a[0]/nhist:6.1f}" if nhist else ""
        print(f"{k:18s} inst {100*a[0]/ti:5.1f}%  stall {100*a[2]/max(ts,1):5.1f}%  lanes {a[1]/max(a[0],1):5.1f}{extra}")


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Per-CUDA-source-line totals from `ncu --page source --print-source cuda,sass --csv`: share of warp instructions,
share of stall samples, and average active lanes per line.  Usage: ncu_lines.py file.ncu-rep [topN]"""
import csv, io, subprocess, sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
    hdr = rows[hi]
    iI, iT, iS = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
    lines = []
    for r in rows[hi + 1:]:
        if len(r) <= iT or not r[0].strip().isdigit():
            continue
        num = lambda v: int(v) if v.strip().isdigit() else 0
        lines.append((int(r[0]), r[1].strip(), num(r[iI]), num(r[iT]), num(r[iS])))
    tot_i = sum(l[2] for l in lines)
    tot_t = sum(l[3] for l in lines)
    tot_s = sum(l[4] for l in lines)
    print(f"total warp instr {tot_i:.4e}, thread instr {tot_t:.4e} (avg lanes {tot_t/tot_i:.2f}), samples {tot_s}")
    for ln, src, i, t, s in sorted(lines, key=lambda l: -l[2])[:top]:
        print(f"{100*i/tot_i:5.1f}% inst {100*s/max(tot_s,1):5.1f}% stall  lanes {t/max(i,1):5.1f} | {ln:4d} {src[:120]}")


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Summarise one kernel of an .ncu-rep (read with `ncu -i ... --page raw --csv`): the metrics the roofline and
the design discussion in DESIGN.md quote.  Usage: python profiles/ncu_summary.py gpurun_out/x.ncu-rep [row]"""
import csv, subprocess, sys, io

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum",
    "lts__t_sectors_srcunit_tex_op_read.sum",
]


def main():
    rep = sys.argv[1]
    row = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, r = rows[0], rows[1], rows[2 + row]
    idx = {h: i for i, h in enumerate(hdr)}
    print("kernel:", r[idx.get("Kernel Name", 4)])
    for k in KEYS:
        if k in idx:
            print(f"{k:75s} {r[idx[k]]:>18s} {units[idx[k]]}")
    for h, i in idx.items():
        if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
            print(f"{h:75s} {r[i]:>18s}")


if __name__ == "__main__":
    main()

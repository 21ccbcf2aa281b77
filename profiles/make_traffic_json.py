#!/usr/bin/env python3
"""profiles/*traffic*.json from an .ncu-rep and the target's own output line (profiles/ncu_target.py): DRAM bytes and warp
instructions per history, keyed by the kernel build id of the library that was profiled.  bench.py prints roofline.traffic
only when that id equals dxb_kernel_build_id() of the library it loaded.
Usage: python profiles/make_traffic_json.py gpurun_out/r02_pool.ncu-rep gpurun_out/r02_pool_target.json out.json"""
import csv
import io
import json
import subprocess
import sys

rep, target, out = sys.argv[1:4]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units, r = rows[0], rows[1], rows[2]
col = {h: i for i, h in enumerate(hdr)}


def val(name):
    v, u = float(r[col[name]].replace(",", "")), units[col[name]]
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "Tbyte": 1e12}.get(u, 1.0)


t = [json.loads(l) for l in open(target) if l.startswith("{")][-1]
h = t["histories"]
rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
doc = {
    "kernel": r[col["Kernel Name"]] if "Kernel Name" in col else "transportKernelPool",
    "library_build": t["kernel_build"],
    "config": t["workload"],
    "source": "ncu --set full --clock-control none --import-source on, one launch (profiles/ncu_capture_r02.sh, profiles/ncu_target.py, "
              "profiles/make_traffic_json.py)",
    "histories_in_capture": h,
    "steps_per_history": t["steps"] / h,
    "hops_per_history": t["hops"] / h,
    "deposits_per_history": t["deposits"] / h,
    "local_majorant": t["local_majorant"],
    "dram_bytes_read": rd,
    "dram_bytes_write": wr,
    "dram_bytes_per_history": (rd + wr) / h,
    "warp_instructions_per_history": val("smsp__inst_executed.sum") / h,
    "active_lanes_per_instruction": val("smsp__thread_inst_executed_per_inst_executed.ratio"),
    "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "duration_ms_under_ncu": val("gpu__time_duration.sum") * {"ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}.get(units[col["gpu__time_duration.sum"]], 1.0),
}
json.dump(doc, open(out, "w"), indent=1)
print(json.dumps(doc))

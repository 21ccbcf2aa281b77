"""Where does the host read-out bandwidth go?  For every visible GPU and every NUMA node of the box: D2H copy rate into
pinned memory allocated while the thread runs on that node, alone and all GPUs at once with local / remote placement.
Prints JSON lines.  torch is used for the copies only."""
import glob
import json
import os
import time

import torch

n = torch.cuda.device_count()
nodes = {}
for p in sorted(glob.glob("/sys/devices/system/node/node[0-9]*")):
    k = int(p.rsplit("node", 1)[1])
    cpus = []
    for part in open(p + "/cpulist").read().strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus += list(range(int(a), int(b or a) + 1))
    nodes[k] = cpus
allowed = sorted(os.sched_getaffinity(0))
info = {"gpus": n, "nodes": {k: "%d cpus (%s..%s)" % (len(v), v[0] if v else None, v[-1] if v else None) for k, v in nodes.items()},
        "allowed_cpus": len(allowed), "allowed_first_last": [allowed[0], allowed[-1]]}
gpu_node = {}
for i in range(n):
    bus = torch.cuda.get_device_properties(i)
    try:
        pci = "%04x:%02x:%02x.0" % (bus.pci_domain_id, bus.pci_bus_id, bus.pci_device_id)
        gpu_node[i] = int(open("/sys/bus/pci/devices/%s/numa_node" % pci).read())
    except Exception as e:  # noqa: BLE001
        gpu_node[i] = "? (%s)" % e
info["gpu_numa_node"] = gpu_node
try:
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    pynvml.nvmlDeviceSetCpuAffinity(h)
    info["nvml_affinity_gpu0"] = "ok: %d cpus" % len(os.sched_getaffinity(0))
    os.sched_setaffinity(0, allowed)
except Exception as e:  # noqa: BLE001
    info["nvml_affinity_gpu0"] = "failed: %r" % (e,)
print(json.dumps(info), flush=True)

MB = 240
src = [torch.empty(MB << 20, dtype=torch.uint8, device=f"cuda:{i}") for i in range(n)]
streams = [torch.cuda.Stream(device=i) for i in range(n)]
bufs = {}
for k, cpus in nodes.items():
    use = [c for c in cpus if c in allowed]
    if not use:
        continue
    os.sched_setaffinity(0, use)
    bufs[k] = [torch.empty(MB << 20, dtype=torch.uint8).pin_memory() for _ in range(n)]
    for b in bufs[k]:
        b.fill_(1)
os.sched_setaffinity(0, allowed)


def run(pairs, reps=5):
    for g, _ in pairs:
        torch.cuda.synchronize(g)
    t0 = time.perf_counter()
    for _ in range(reps):
        for g, k in pairs:
            with torch.cuda.stream(streams[g]):
                bufs[k][g].copy_(src[g], non_blocking=True)
    for g, _ in pairs:
        streams[g].synchronize()
    return round(reps * len(pairs) * MB * (1 << 20) / (time.perf_counter() - t0) / 1e9, 1)


ks = sorted(bufs)
run([(g, ks[0]) for g in range(n)], 1)
single = {"gpu%d->node%d" % (g, k): run([(g, k)]) for g in range(min(n, 8)) for k in ks}
print(json.dumps({"single_GBps": single}), flush=True)
best = [(g, max(ks, key=lambda k: single["gpu%d->node%d" % (g, k)])) for g in range(n)]
out = {"all_best_node_GBps": run(best), "best": best}
for k in ks:
    out["all_to_node%d_GBps" % k] = run([(g, k) for g in range(n)])
print(json.dumps(out), flush=True)

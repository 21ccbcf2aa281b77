"""Host read-out ceiling of the box: D2H copy rate of one GPU and of all GPUs at once into pinned host memory (what bounds
dxb_get_dose, 24 B per voxel).  Prints one JSON line.  torch is used for the copies only."""
import json
import sys
import time

import torch

n = torch.cuda.device_count()
MB = 240
src = [torch.empty(MB << 20, dtype=torch.uint8, device=f"cuda:{i}") for i in range(n)]
dst = [torch.empty(MB << 20, dtype=torch.uint8).pin_memory() for _ in range(n)]
streams = [torch.cuda.Stream(device=i) for i in range(n)]


def run(devs, reps=5):
    for i in devs:
        torch.cuda.synchronize(i)
    t0 = time.perf_counter()
    for _ in range(reps):
        for i in devs:
            with torch.cuda.stream(streams[i]):
                dst[i].copy_(src[i], non_blocking=True)
    for i in devs:
        streams[i].synchronize()
    dt = time.perf_counter() - t0
    return reps * len(devs) * MB * (1 << 20) / dt / 1e9


run(range(n), 1)
out = {"gpus": n, "MB_per_copy": MB, "one_gpu_GBps": round(run([0]), 1), "all_gpus_GBps": round(run(range(n)), 1)}
if n >= 4:
    out["half_GBps"] = round(run(range(n // 2)), 1)
    out["other_half_GBps"] = round(run(range(n // 2, n)), 1)
# H2D for the grid upload
def run_h2d(devs, reps=5):
    for i in devs:
        torch.cuda.synchronize(i)
    t0 = time.perf_counter()
    for _ in range(reps):
        for i in devs:
            with torch.cuda.stream(streams[i]):
                src[i].copy_(dst[i], non_blocking=True)
    for i in devs:
        streams[i].synchronize()
    return reps * len(devs) * MB * (1 << 20) / (time.perf_counter() - t0) / 1e9
out["h2d_one_gpu_GBps"] = round(run_h2d([0]), 1)
out["h2d_all_gpus_GBps"] = round(run_h2d(range(n)), 1)
print(json.dumps(out))

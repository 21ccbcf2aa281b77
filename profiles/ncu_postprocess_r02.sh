#!/bin/bash
# After profiles/ncu_capture_r02.sh came back (gpurun_out/*.ncu-rep): summaries, per-region split, traffic files, launch list.
set -e
python profiles/ncu_summary.py gpurun_out/r02_pool.ncu-rep > profiles/r02_transport_pool_summary.txt 2>&1
python profiles/ncu_summary.py gpurun_out/r02_pool_lm.ncu-rep > profiles/r02_transport_pool_lm_summary.txt 2>&1
python profiles/ncu_buckets.py gpurun_out/r02_pool.ncu-rep > profiles/r02_transport_pool_regions.txt 2>&1
python profiles/ncu_buckets.py gpurun_out/r02_pool_lm.ncu-rep > profiles/r02_transport_pool_lm_regions.txt 2>&1
python profiles/make_traffic_json.py gpurun_out/r02_pool.ncu-rep gpurun_out/r02_pool_target.json profiles/r02_pool_traffic.json
python profiles/make_traffic_json.py gpurun_out/r02_pool_lm.ncu-rep gpurun_out/r02_pool_lm_target.json profiles/r02_pool_lm_c3_capture.json > /dev/null
cp gpurun_out/r02_launches.csv profiles/r02_launches.csv
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(l for l in open('gpurun_out/r02_launches.csv') if not l.startswith('=='))]
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
tot, cnt = collections.OrderedDict(), collections.Counter()
for r in rows[1:]:
    if len(r) < len(hdr) or r[idx['Metric Name']] != 'gpu__time_duration.sum':
        continue
    k = r[idx['Kernel Name']]
    v = float(r[idx['Metric Value']].replace(',', '')) * {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}.get(r[idx['Metric Unit']], 1e-6)
    tot[k] = tot.get(k, 0) + v
    cnt[k] += 1
s = sum(tot.values())
out = ["ncu launch list of `python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e` (1 GPU, C2, 1e9 histories per step; "
       "per-launch times are cold-cache and serialised: shares, not absolutes)", ""]
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    out.append("%-90s launches %4d  total %10.3f ms  share %6.2f %%" % (k[:90], cnt[k], v, 100 * v / s))
open('profiles/r02_launches_summary.txt', 'w').write("\n".join(out) + "\n")
print(out[2])
PY

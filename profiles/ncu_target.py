"""Target of the round-2 ncu captures (profiles/ncu_capture_r02.sh): two transport launches of one configuration (the first
is skipped as warm-up), 3e7 histories each.  Prints what the traffic file needs: kernel build id, histories, counters."""
import json
import sys
sys.path.insert(0, ".")
import opendxmc_b200 as dx
from opendxmc_b200 import _capi as K

which = sys.argv[1] if len(sys.argv) > 1 else "c2"
nh = int(float(sys.argv[2])) if len(sys.argv) > 2 else 30_000_000
wl = {"c2": lambda: dx.workloads.ct_spiral_patient(scale=1, histories=nh),
      "c3": lambda: dx.workloads.icrp_phantom("AM", histories=nh),
      "c5": lambda: dx.workloads.icrp_phantom("10M", histories=nh, beam_kind="dx")}[which]()
world = wl.build_world(1, [0])
tr = dx.Transport()
for _ in range(2):
    tr.run_transport(world, wl.beam)
st = world.run_stats()
print(json.dumps({"config": which, "workload": wl.name, "kernel_build": K.load().dxb_kernel_build_id().decode(), "histories": st["histories"],
                  "steps": st["steps"], "hops": st["hops"], "deposits": st["deposits"], "local_majorant": st["local_majorant"],
                  "transport_ms_under_ncu": st["transport_ms"]}))
world.close()

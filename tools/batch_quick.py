"""Batch timing without torch: `python tools/batch_quick.py <n_graphs> [reps]` solves n config-2 graphs (BASELINE config 4's unit)
in one persistent launch and prints the kernel time.  Used to compare team modes (PUS_CLUSTER=0/8/16 in the environment)."""
import sys
sys.path.insert(0, '.')
from pop_up_slam_b200 import capi, graphgen as gg
from pop_up_slam_b200.capi import GpuGraphAPI

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
apis = []
for s in range(n):
    g = gg.make_config(2, seed=s)
    a = GpuGraphAPI()
    gg.build_bulk(a, g)
    gg.configure(a, g)
    apis.append(a)
capi.upload_many(apis)
capi.solve_resident_many(apis)
ms = []
for _ in range(reps):
    its = capi.solve_resident_many(apis)
    ms.append(apis[0].stats()["kernel_ms"])
st = apis[0].stats()
print("batch of %d: kernel_ms min %.3f median %.3f  grid %d  iters %d  chi2[0] %.9g" %
      (n, min(ms), sorted(ms)[len(ms) // 2], st["grid_ctas"], int(its.sum()), st["chi2_final"]))

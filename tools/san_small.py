"""Small solves for compute-sanitizer runs (memcheck / racecheck / synccheck): a few-CTA graph through the default path,
the forced large-graph path (bulk-copy staging) and the spanning emulation.
usage: compute-sanitizer --tool memcheck python tools/san_small.py [n_poses]"""
import ctypes, sys
sys.path.insert(0, '.')
from pop_up_slam_b200 import graphgen as gg, capi
from pop_up_slam_b200.capi import GpuGraphAPI
n = int(sys.argv[1]) if len(sys.argv) > 1 else 48
g = gg.make_config(2, seed=0, n_poses=n, n_planes=12, max_iterations=3)
for flag in (0, 2):
    a = GpuGraphAPI(); gg.build_bulk(a, g); gg.configure(a, g)
    o = a.get_solver_options(); o.reserved[2] = flag
    a._chk(a.lib.pus_set_solver_options(a.h, ctypes.byref(o)))
    print("flag", flag, "iters", a.batch_optimize(), "chi2 %.6g" % a.chi2(), "grid", a.stats()["grid_ctas"])
apis = []
for _ in range(2):
    a = GpuGraphAPI(); gg.build_bulk(a, g); gg.configure(a, g); apis.append(a)
print("spanning emulation iters", capi.span_emulate_optimize(apis), "chi2 %.6g" % apis[0].chi2())

"""lazy preconditioner refresh threshold sweep on a config (solver option reserved[3] = percent of the post-build PCG count)"""
import sys, ctypes
sys.path.insert(0, '.')
from pop_up_slam_b200 import graphgen as gg
from pop_up_slam_b200.capi import GpuGraphAPI
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 3
g = gg.make_config(cfg, seed=0)
for pct in [int(x) for x in sys.argv[2:]] or [200]:
    a = GpuGraphAPI(); gg.build_bulk(a, g); gg.configure(a, g)
    o = a.get_solver_options(); o.reserved[3] = pct
    if pct == 1: o.reserved[0] = 1
    if pct < 0: o.reserved[0] = 2; o.reserved[3] = -pct
    a._chk(a.lib.pus_set_solver_options(a.h, ctypes.byref(o)))
    a.upload()
    for _ in range(3): it = a.solve_resident()
    st = a.stats(); a.download()
    print("pct", pct, "iters", it, "pcg", st["pcg_iterations"], "builds", round(st["phase_ms"][5]), "kernel_ms %.2f" % st["kernel_ms"], "pcg/LM", a.trace()["pcg"].tolist())

import sys, ctypes
sys.path.insert(0,'.')
from pop_up_slam_b200 import graphgen as gg
from pop_up_slam_b200.capi import GpuGraphAPI
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
g = gg.make_config(cfg, seed=0)
for team in [int(x) for x in sys.argv[2:]] or [0]:
    a = GpuGraphAPI(); gg.build_bulk(a, g); gg.configure(a, g)
    o = a.get_solver_options(); o.team_ctas = team
    a._chk(a.lib.pus_set_solver_options(a.h, ctypes.byref(o)))
    a.upload()
    for _ in range(3): it = a.solve_resident()
    st = a.stats()
    print("cfg", cfg, "team", team, "-> grid", st["grid_ctas"], "iters", it, "pcg", st["pcg_iterations"], "kernel_ms %.3f" % st["kernel_ms"])

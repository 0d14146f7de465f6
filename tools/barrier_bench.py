import sys, ctypes
sys.path.insert(0,'.')
from pop_up_slam_b200 import graphgen as gg
from pop_up_slam_b200.capi import GpuGraphAPI
g = gg.make_config(3, seed=0)
for team in (2, 8, 32, 64, 120, 148):
    a = GpuGraphAPI(); gg.build_bulk(a, g); gg.configure(a, g)
    o = a.get_solver_options(); o.team_ctas = team
    a._chk(a.lib.pus_set_solver_options(a.h, ctypes.byref(o)))
    try:
        a.upload(); a.debug_run_stage(9, 0.0); a.debug_run_stage(9, 0.0)
    except Exception as e:
        print(team, "failed:", e); continue
    ph = a.stats()["phase_ms"]
    print("team %3d: barrier %.2f us, reduce %.2f us" % (team, ph[0], ph[1]))

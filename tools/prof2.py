import sys
sys.path.insert(0,'.')
import numpy as np
from pop_up_slam_b200 import graphgen as gg
from pop_up_slam_b200.capi import GpuGraphAPI
g = gg.make_config(3, seed=0)
for team, always in [(0,0),(0,1),(64,1),(148,1),(48,1)]:
    a = GpuGraphAPI(); gg.build_bulk(a, g); gg.configure(a, g)
    o = a.get_solver_options(); o.team_ctas = team; o.reserved[0] = always
    a._chk(a.lib.pus_set_solver_options(a.h, __import__('ctypes').byref(o)))
    a.upload()
    for _ in range(2): it = a.solve_resident()
    st = a.stats(); ph = st['phase_ms']
    print("team", st['grid_ctas'], "always_rebuild", always, "iters", it, "pcg", st['pcg_iterations'], "kernel_ms %.1f"%st['kernel_ms'],
          "setup %.1f pcg %.1f  | sweepPl %.2f pose %.2f prec %.2f us/it"%(ph[1], ph[2], ph[16]/st['pcg_iterations']*1e3, ph[19]/st['pcg_iterations']*1e3, ph[20]/st['pcg_iterations']*1e3), "chi2 %.6f"%st['chi2_final'])

import sys, ctypes
sys.path.insert(0,'.')
from pop_up_slam_b200 import graphgen as gg
from pop_up_slam_b200.capi import GpuGraphAPI
g = gg.make_config(3, seed=0)
for pct in (150, 200, 300, 500, 1000):
    a = GpuGraphAPI(); gg.build_bulk(a, g); gg.configure(a, g)
    o = a.get_solver_options(); o.reserved[3] = pct
    a._chk(a.lib.pus_set_solver_options(a.h, ctypes.byref(o)))
    a.upload()
    for _ in range(2): it = a.solve_resident()
    a.download()
    st = a.stats(); ph = st['phase_ms']
    print("pct", pct, "builds", round(ph[5]), "pcg", st['pcg_iterations'], "kernel_ms %.1f"%st['kernel_ms'], "AcInv %.1f blocks %.1f"%(ph[12], ph[9]), "pcg per LM:", a.trace()['pcg'].tolist())

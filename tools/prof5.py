import sys, ctypes
sys.path.insert(0,'.')
from pop_up_slam_b200 import graphgen as gg
from pop_up_slam_b200.capi import GpuGraphAPI
g = gg.make_config(3, seed=0)
a = GpuGraphAPI(); gg.build_bulk(a, g); gg.configure(a, g)
o = a.get_solver_options(); o.reserved[2] = 1
a._chk(a.lib.pus_set_solver_options(a.h, ctypes.byref(o)))
a.upload(); a.debug_run_stage(1, 1e-6)
st = a.stats(); ph = st['phase_ms']
print("one setup: Hinv %.3f blocks %.3f Wc %.3f Ac %.3f AcInv %.3f ms"%(ph[8],ph[9],ph[10],ph[11],ph[12]))
print(" AcInv parts (40 steps): load %.1f  pivot-inv %.1f  T %.1f  chunks %.1f  barrier %.1f us/step"%tuple(ph[k]*1e3/40 for k in (6,7,21,22,23)))

import sys
sys.path.insert(0,'.')
import numpy as np
from pop_up_slam_b200 import graphgen as gg
from pop_up_slam_b200.capi import GpuGraphAPI
cfg = int(sys.argv[1]) if len(sys.argv)>1 else 3
reps = int(sys.argv[2]) if len(sys.argv)>2 else 2
kw = {}
if len(sys.argv)>3: kw['max_iterations']=int(sys.argv[3])
g = gg.make_config(cfg, seed=0, **kw)
a = GpuGraphAPI(); gg.build_bulk(a, g); gg.configure(a, g)
import ctypes
o = a.get_solver_options(); o.reserved[2] = int(sys.argv[4]) if len(sys.argv)>4 else 0
if len(sys.argv)>5: o.pcg_rel_tol = float(sys.argv[5])
if len(sys.argv)>6: o.reserved[3] = int(sys.argv[6])
a._chk(a.lib.pus_set_solver_options(a.h, ctypes.byref(o)))
a.upload()
for _ in range(reps):
    it = a.solve_resident()
st = a.stats()
ph = st['phase_ms']
print("prec builds", round(st["phase_ms"][5]), end=" | ")
print("cfg", cfg, g.dims(), "iters", it, "pcg", st['pcg_iterations'], "kernel_ms %.2f"%st['kernel_ms'], "ctas", st['grid_ctas'])
names = {0:'linearize',1:'setup',2:'pcg',3:'update',4:'chi2',8:'s.Hinv',9:'s.blocks',10:'s.Wc',11:'s.Ac',12:'s.AcInv',16:'p.sweepPl',17:'r.tiles',18:'r.side',19:'p.poseRed',20:'p.precRed',21:'p.pose.vg',22:'probe.22',23:'probe.23',6:'probe.6',13:'p.prec.rc',14:'p.prec.coarse',15:'p.prec.blocks'}
nset = st['lm_iterations']+1
for k,n in names.items():
    per = ph[k]/nset*1e3 if (k<13) else ph[k]/max(1,st['pcg_iterations'])*1e3
    print("  %-10s %9.3f ms  (%8.2f us per %s)"%(n, ph[k], per, 'LM solve' if k<13 else 'PCG iter'))

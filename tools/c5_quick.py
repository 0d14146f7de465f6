"""BASELINE config 5 (or a reduced version), a few LM iterations, resident: kernel ms, PCG iterations, phase split.
usage: python tools/c5_quick.py [max_iterations=3] [n_poses=50000] [reserved2 flags=0]"""
import sys, ctypes
sys.path.insert(0, '.')
from pop_up_slam_b200 import graphgen as gg
from pop_up_slam_b200.capi import GpuGraphAPI
it = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n = int(sys.argv[2]) if len(sys.argv) > 2 else 50000
fl = int(sys.argv[3]) if len(sys.argv) > 3 else 0
g = gg.make_config(5, seed=0, max_iterations=it, n_poses=n, n_planes=n // 10)
a = GpuGraphAPI(); gg.build_bulk(a, g); gg.configure(a, g)
o = a.get_solver_options(); o.reserved[2] = fl | 1
a._chk(a.lib.pus_set_solver_options(a.h, ctypes.byref(o)))
a.upload()
for _ in range(2): its = a.solve_resident()
st = a.stats()
ph = st["phase_ms"]
print("N", n, "flags", fl, "iters", its, "pcg", st["pcg_iterations"], "kernel_ms %.2f" % st["kernel_ms"], "chi2 %.6f" % st["chi2_final"], "grid", st["grid_ctas"])
print("phases: lin %.2f setup %.2f pcg %.2f upd %.2f chi2 %.2f | Hll^-1 %.2f blocks %.2f Wc %.2f A_c+groups %.2f A_c^-1 %.2f | sweep %.2f pose %.2f prec %.2f" %
      (ph[0], ph[1], ph[2], ph[3], ph[4], ph[8], ph[9], ph[10], ph[11], ph[12], ph[16], ph[19], ph[20]))

import sys, ctypes
sys.path.insert(0,'.')
from pop_up_slam_b200 import graphgen as gg
from pop_up_slam_b200 import capi
g = gg.make_config(3, seed=0)
for name in sys.argv[1:]:
    lib = ctypes.CDLL(name)
    a = capi.GpuGraphAPI(lib=lib); gg.build_bulk(a, g); gg.configure(a, g)
    a.upload()
    for _ in range(2): it = a.solve_resident()
    st = a.stats(); ph = st['phase_ms']
    print(name, "iters", it, "pcg", st['pcg_iterations'], "kernel_ms %.1f"%st['kernel_ms'], "builds", "setup %.1f pcg %.1f | sweepPl %.2f pose %.2f prec %.2f us/it"%(ph[1], ph[2], ph[16]/st['pcg_iterations']*1e3, ph[19]/st['pcg_iterations']*1e3, ph[20]/st['pcg_iterations']*1e3), "s.AcInv %.1f"%ph[12])

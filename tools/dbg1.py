import sys
sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import numpy as np
from oracle_api import OracleAPI
from pop_up_slam_b200 import graphgen as gg
from pop_up_slam_b200.capi import GpuGraphAPI
np.set_printoptions(linewidth=220, precision=9)
for cfg, seed, jac in [(2,0,0),(2,3,1)]:
    g = gg.make_config(cfg, seed=seed)
    gpu, orc = GpuGraphAPI(), OracleAPI(); orc.set_jacobian_mode(jac)
    ig, io = gg.build_interleaved(gpu, g), gg.build_interleaved(orc, g)
    gg.configure(gpu, g); gg.configure(orc, g)
    it_g, it_o = gpu.batch_optimize(), orc.batch_optimize()
    tg, to = gpu.trace(), orc.trace()
    print(cfg, seed, jac, "iters", it_g, it_o)
    print(" gpu chi2", tg['chi2_new'], tg['accepted'], "pcg", tg['pcg'], "dn", tg['delta_norm'], "lam", tg['lam'])
    print(" orc chi2", to['chi2_new'], to['accepted'], "dn", to['delta_norm'], "lam", to['lam'])
    print(" stats", {k:v for k,v in gpu.stats().items() if k in ('pcg_iterations','kernel_ms','phase_ms','grid_ctas')})
    Pg, Po = gpu.get_poses(ig['pose_ids']), orc.get_poses(io['pose_ids'])
    print(" max pos diff", np.abs(Pg[:,:3]-Po[:,:3]).max(), "chi2", gpu.chi2(), orc.chi2())

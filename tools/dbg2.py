import sys
sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import numpy as np
from oracle_api import OracleAPI
from pop_up_slam_b200 import graphgen as gg
from pop_up_slam_b200.capi import GpuGraphAPI
np.set_printoptions(linewidth=220, precision=6)
g = gg.make_config(2, seed=0)
for build in ("bulk", "inter"):
    for lam in (0.0, 1e-6):
        gpu, orc = GpuGraphAPI(), OracleAPI(); orc.set_jacobian_mode(1)
        bf = gg.build_bulk if build == "bulk" else gg.build_interleaved
        ig, io = bf(gpu, g), bf(orc, g)
        gg.configure(gpu, g); gg.configure(orc, g)
        ref = orc.solve_step(lam)
        gpu.upload(); gpu.debug_run_stage(2, lam)
        N, M = g.n_poses, g.n_planes
        x = gpu.debug_fetch("x", 6*N); dl = gpu.debug_fetch("dl", 3*M)
        # reference in compiled order: poses then planes
        starts_p = np.array([orc.node_start(i) for i in io['pose_ids']]); starts_l = np.array([orc.node_start(i) for i in io['plane_ids']])
        refp = np.concatenate([ref[s:s+6] for s in starts_p]); refl = np.concatenate([ref[s:s+3] for s in starts_l])
        st = gpu.stats()
        print(build, lam, "|x|", np.linalg.norm(x), "|ref|", np.linalg.norm(refp), "relerr poses", np.linalg.norm(x-refp)/np.linalg.norm(refp), "planes", np.linalg.norm(dl-refl)/np.linalg.norm(refl), "pcg", st['pcg_iterations'], "ctas", st['grid_ctas'])
        if np.linalg.norm(x-refp)/np.linalg.norm(refp) > 1e-6:
            # which stage is off? compare operator & rhs
            b = gpu.debug_fetch("b", 6*N)
            A, bb = orc.normal_equations(lam); A = A.toarray()
            idx = np.concatenate([np.arange(s, s+6) for s in starts_p] + [np.arange(s, s+3) for s in starts_l])
            A = A[np.ix_(idx, idx)]; bb = bb[idx]
            App, Apl, All = A[:6*N,:6*N], A[:6*N,6*N:], A[6*N:,6*N:]
            S = App - Apl @ np.linalg.solve(All, Apl.T)
            rhs = bb[:6*N] - Apl @ np.linalg.solve(All, bb[6*N:])
            print("   rhs relerr", np.linalg.norm(b-rhs)/np.linalg.norm(rhs))
            rng = np.random.default_rng(0); xx = rng.normal(size=6*N)
            gpu.debug_store("pv0", xx); gpu.debug_run_stage(3, lam); q = gpu.debug_fetch("q", 6*N)
            print("   S*x relerr", np.linalg.norm(q - S@xx)/np.linalg.norm(S@xx))
            xs = np.linalg.solve(S, rhs); print("   dense solve vs ref", np.linalg.norm(xs-refp)/np.linalg.norm(refp), " gpu x vs dense", np.linalg.norm(x-xs)/np.linalg.norm(xs))
            print("   residual of gpu x: ", np.linalg.norm(S@x-rhs)/np.linalg.norm(rhs), "cond(S) %.3g"%np.linalg.cond(S))

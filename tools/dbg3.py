import sys
sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import numpy as np
from oracle_api import OracleAPI
from pop_up_slam_b200 import graphgen as gg
from pop_up_slam_b200.capi import GpuGraphAPI
np.set_printoptions(linewidth=220, precision=6)
g = gg.make_config(2, seed=0)
gpu, orc = GpuGraphAPI(), OracleAPI(); orc.set_jacobian_mode(1)
ig, io = gg.build_interleaved(gpu, g), gg.build_interleaved(orc, g)
gg.configure(gpu, g); gg.configure(orc, g)
N, M = g.n_poses, g.n_planes
A, bb = orc.normal_equations(0.0); A = A.toarray()
sp_ = np.array([orc.node_start(i) for i in io['pose_ids']]); sl_ = np.array([orc.node_start(i) for i in io['plane_ids']])
gpu.upload(); gpu.debug_run_stage(0)
Hpp = gpu.debug_fetch("Hpp", N*36).reshape(N,6,6); gp = gpu.debug_fetch("gp", N*6).reshape(N,6)
Hll = gpu.debug_fetch("Hll", M*9).reshape(M,3,3); gl = gpu.debug_fetch("gl", M*3).reshape(M,3)
pose_node = gpu.debug_fetch("pose_node", N).astype(int); plane_node = gpu.debug_fetch("plane_node", M).astype(int)
print("pose_node ok", np.array_equal(pose_node, io['pose_ids']), "plane_node ok", np.array_equal(plane_node, io['plane_ids']))
eh = [np.abs(Hpp[p]-A[sp_[p]:sp_[p]+6, sp_[p]:sp_[p]+6]).max() for p in range(N)]
eg = [np.abs(gp[p]+bb[sp_[p]:sp_[p]+6]).max() for p in range(N)]
el = [np.abs(Hll[l]-A[sl_[l]:sl_[l]+3, sl_[l]:sl_[l]+3]).max() for l in range(M)]
egl = [np.abs(gl[l]+bb[sl_[l]:sl_[l]+3]).max() for l in range(M)]
print("Hpp err max", max(eh), "at", int(np.argmax(eh)), " gp err", max(eg), int(np.argmax(eg)), " Hll", max(el), int(np.argmax(el)), " gl", max(egl), int(np.argmax(egl)))
print("bad gp poses:", [p for p in range(N) if eg[p] > 1e-6][:20])
print("bad gl planes:", [l for l in range(M) if egl[l] > 1e-6][:20])
p = int(np.argmax(eg)); print(gp[p], -bb[sp_[p]:sp_[p]+6])

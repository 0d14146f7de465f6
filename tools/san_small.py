"""Small solves for compute-sanitizer runs (memcheck / racecheck / synccheck): a few-CTA graph through the default path,
the forced large-graph path (bulk-copy staging), the forced three-level preconditioner, the reference-Jacobian mode, the
resident measurement refresh and the spanning emulation.
usage: compute-sanitizer --tool memcheck python tools/san_small.py [n_poses]"""
import ctypes, sys
sys.path.insert(0, '.')
from pop_up_slam_b200 import graphgen as gg, capi
from pop_up_slam_b200.capi import GpuGraphAPI
n = int(sys.argv[1]) if len(sys.argv) > 1 else 48
g = gg.make_config(2, seed=0, n_poses=n, n_planes=12, max_iterations=3)
for flag in (0, 2, 8, 8 | 2):   # default; bulk-copy data path; three-level preconditioner; both
    a = GpuGraphAPI(); gg.build_bulk(a, g); gg.configure(a, g)
    o = a.get_solver_options(); o.reserved[2] = flag
    a._chk(a.lib.pus_set_solver_options(a.h, ctypes.byref(o)))
    print("flag", flag, "iters", a.batch_optimize(), "chi2 %.6g" % a.chi2(), "grid", a.stats()["grid_ctas"])
apis = []
for _ in range(2):
    a = GpuGraphAPI(); gg.build_bulk(a, g); gg.configure(a, g); apis.append(a)
print("spanning emulation iters", capi.span_emulate_optimize(apis), "chi2 %.6g" % apis[0].chi2())

import numpy as np
a = GpuGraphAPI(); a.set_jacobian_mode(0); ids = gg.build_bulk(a, g); gg.configure(a, g)
print("reference-Jacobian mode iters", a.batch_optimize(), "chi2 %.6g" % a.chi2())
rng = np.random.default_rng(0)
nf = g.n_poses
seg_ptr = (np.arange(nf + 1) * 3).astype(np.int32)
segs = np.stack([rng.uniform(0, 640, 3 * nf), rng.uniform(300, 480, 3 * nf), rng.uniform(0, 640, 3 * nf), rng.uniform(300, 480, 3 * nf)], axis=1).astype(np.float32)
invK = np.linalg.inv(np.array([[535.4, 0, 320.1], [0, 539.2, 247.6], [0, 0, 1.0]])).astype(np.float32)
mf = g.pp_pose.astype(np.int32); mr = (np.arange(len(mf)) % 4).astype(np.int32)
a.refresh_bind(ids["pose_ids"], seg_ptr, segs, invK, ids["pp_fids"], mf, mr)
a.refresh_run(); a.refresh_run(want_output=True)
print("resident refresh ok, next solve iters", a.batch_optimize())

"""Frame-by-frame use as Mapper_mono::processFrame drives the back end (Mapping.cpp:464-554): per key-frame one new
pose + odometry + new planes + pose-plane factors, then Slam::update() (batch_optimization() every 5th frame).
Reports the per-frame latency of the back-end calls for the GPU library and for the CPU port (oracle).
usage: python tools/incremental_bench.py [n_poses] [n_planes]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from pop_up_slam_b200 import graphgen as gg, geometry as geo
from pop_up_slam_b200.capi import GpuGraphAPI
from oracle_api import OracleAPI

n_poses = int(sys.argv[1]) if len(sys.argv) > 1 else 300
n_planes = int(sys.argv[2]) if len(sys.argv) > 2 else 60
g = gg.make_config(2, seed=0, n_poses=n_poses, n_planes=n_planes)
order = np.argsort(g.pp_pose, kind="stable")
ptr = np.searchsorted(g.pp_pose[order], np.arange(g.n_poses + 1))
out = {}
for name, api in (("gpu", GpuGraphAPI()), ("cpu_port", OracleAPI())):
    if name == "cpu_port":
        api.set_jacobian_mode(0)
    gg.configure(api, g)
    pose_ids, plane_ids = [], {}
    t_solve, t_build = [], []
    for i in range(g.n_poses):
        t0 = time.perf_counter()
        pose_ids.append(api.add_pose(None))
        if i == 0:
            api.add_pose_prior(pose_ids[0], g.prior_meas, g.prior_sqrtinf)
        else:
            api.add_odometry(pose_ids[i - 1], pose_ids[i], g.odo_meas[i - 1], g.odo_sqrtinf[i - 1])
        for e in order[ptr[i]:ptr[i + 1]]:
            k = int(g.pp_plane[e])
            if k not in plane_ids:
                plane_ids[k] = api.add_plane(None)
                if k == 0:
                    api.init_plane(plane_ids[k], geo.plane_to_global(geo.pose7_to_T(api.get_pose(pose_ids[i])), g.pp_meas[e]))
                    api.add_plane_prior(plane_ids[k], g.ground_meas, g.ground_sqrtinf)
            api.add_pose_plane(pose_ids[i], plane_ids[k], g.pp_meas[e], g.pp_sqrtinf[e])
        t1 = time.perf_counter()
        if i % 5 == 0:
            api.batch_optimize()
        else:
            api.update()
        t2 = time.perf_counter()
        t_build.append(t1 - t0); t_solve.append(t2 - t1)
    if name == "gpu":
        st = api.stats()
        print("  last frame (update): kernel %.3f ms, h2d %.3f ms (%d bytes), d2h %.3f ms, grid %d CTAs" %
              (st["kernel_ms"], st["h2d_ms"], st["h2d_bytes"], st["d2h_ms"], st["grid_ctas"]))
    ts = np.array(t_solve) * 1e3
    out[name] = ts
    print("%-9s frames %d: back-end call per frame: mean %.2f ms, median %.2f, last-50 mean %.2f, max %.2f; total %.1f ms (graph building %.1f ms)"
          % (name, g.n_poses, ts.mean(), np.median(ts), ts[-50:].mean(), ts.max(), ts.sum(), 1e3 * sum(t_build)))
print("speed-up (sum of back-end calls): %.2fx" % (out["cpu_port"].sum() / out["gpu"].sum()))

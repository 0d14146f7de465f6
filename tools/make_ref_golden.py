"""Generate tests/golden/reference_build.json from oracle/_ref -- the UNMODIFIED reference sources of the hot path
(isam::Slam / Optimizer / Cholesky / numericalDiff, slam3d.h, isam_plane3d.{h,cpp}) compiled against the API shims of
oracle/ref_shim (`make -C oracle ref`, needs the reference checkout at /root/reference).

The fixtures are outputs of the reference's own code: per-factor error() and numericalDiff Jacobians on random inputs, and
whole Levenberg-Marquardt runs (iteration count, lambda / accept trace, chi2, estimates) on BASELINE configs 1, 2 and on
the full-size bench workload (config 3, 20 iterations).  tests/test_reference_golden.py checks the oracle restatement
(CPU) and the CUDA path (GPU) against them without needing the reference at run time.
Regenerate with:  python tools/make_ref_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_api as O  # noqa: E402  (only to draw valid random inputs)
import ref_api as R  # noqa: E402
from pop_up_slam_b200 import geometry as geo, graphgen as gg  # noqa: E402


def rand_pose(rng):
    v = np.concatenate([rng.uniform(-5, 5, 3), [rng.uniform(-3, 3), rng.uniform(-1.2, 1.2), rng.uniform(-3, 3)]])
    return R.pose_from_xyzypr(v)


def rand_plane(rng):
    n = rng.normal(size=3)
    n /= np.linalg.norm(n)
    return geo.plane_normalize(np.append(n, -rng.uniform(0.5, 8.0)))


def ut(rng, n, lo=1.0, hi=30.0):
    A = np.triu(rng.uniform(0.5, 2.0, size=(n, n)))
    A[np.diag_indices(n)] = rng.uniform(lo, hi, n)
    return A[np.triu_indices(n)]


out = {"generator": "tools/make_ref_golden.py", "source": "oracle/_ref (unmodified reference sources + API shims)"}

# ---- per-factor vectors ----
rng = np.random.default_rng(2024)
cases = []
for robust in (0, 1):
    api = R.RefAPI()
    if robust:
        api.set_robust(1, 1.0)
    for i in range(24):
        p, q, l = rand_pose(rng), rand_pose(rng), rand_plane(rng)
        pid, qid, lid = api.add_pose(p), api.add_pose(q), api.add_plane(l)
        meas = R.plane_exmap(R.plane_transform(R.pose_wTo(p), l), rng.normal(0, 0.02, 3))
        si3, si6 = ut(rng, 3, 1.0, 60.0), ut(rng, 6, 1.0, 60.0)
        f = api.add_pose_plane(pid, lid, meas, si3)
        J, r = api.factor_jacobian(f)
        cases.append(dict(kind="pose_plane", robust=robust, b=1.0, pose=p.tolist(), plane=l.tolist(), meas=meas.tolist(), sqrtinf=si3.tolist(),
                          error=api.factor_error(f).tolist(), jacobian=J.tolist()))
        m = R.pose_vector(R.pose_ominus(q, p)) + rng.normal(0, 0.01, 6)
        f = api.add_odometry(pid, qid, m, si6)
        J, r = api.factor_jacobian(f)
        cases.append(dict(kind="odometry", robust=robust, b=1.0, pose=p.tolist(), pose2=q.tolist(), meas=m.tolist(), sqrtinf=si6.tolist(),
                          error=api.factor_error(f).tolist(), jacobian=J.tolist()))
        m = R.pose_vector(p) + rng.normal(0, 0.01, 6)
        f = api.add_pose_prior(pid, m, si6)
        J, r = api.factor_jacobian(f)
        cases.append(dict(kind="pose_prior", robust=robust, b=1.0, pose=p.tolist(), meas=m.tolist(), sqrtinf=si6.tolist(),
                          error=api.factor_error(f).tolist(), jacobian=J.tolist()))
        m = R.plane_exmap(l, rng.normal(0, 0.02, 3))
        f = api.add_plane_prior(lid, m, si3)
        J, r = api.factor_jacobian(f)
        cases.append(dict(kind="plane_prior", robust=robust, b=1.0, plane=l.tolist(), meas=m.tolist(), sqrtinf=si3.tolist(),
                          error=api.factor_error(f).tolist(), jacobian=J.tolist()))
        d6, d3 = rng.normal(0, 0.2, 6), rng.normal(0, 0.2, 3)
        cases.append(dict(kind="exmap", pose=p.tolist(), d6=d6.tolist(), pose_out=R.pose_exmap(p, d6).tolist(), plane=l.tolist(), d3=d3.tolist(),
                          plane_out=R.plane_exmap(l, d3).tolist(), oplus=R.pose_oplus(p, q).tolist(), ominus=R.pose_ominus(q, p).tolist(), pose2=q.tolist()))
out["factors"] = cases

# ---- whole solves ----
runs = {}
for name, cfg, seed, kw, builder, stride in [("config1_seed0", 1, 0, {}, gg.build_interleaved, 1), ("config2_seed0", 2, 0, {}, gg.build_interleaved, 1),
                                             ("config3_small_huber", 3, 0, dict(n_poses=600, n_planes=60), gg.build_bulk, 1),
                                             ("config3_full_20it", 3, 0, {}, gg.build_bulk, 50)]:
    g = gg.make_config(cfg, seed=seed, **kw)
    api = R.RefAPI()
    ids = builder(api, g)
    gg.configure(api, g)
    chi2_0 = api.chi2()
    it = api.batch_optimize()
    tr = api.trace()
    P, L = api.get_poses(ids["pose_ids"]), api.get_planes(ids["plane_ids"])
    runs[name] = dict(config=cfg, seed=seed, kw=kw, builder=builder.__name__, dims=g.dims(), chi2_initial=chi2_0, chi2_final=api.chi2(), iterations=it,
                      accepted=tr["accepted"].tolist(), lambda_trace=tr["lam"].tolist(),
                      chi2_trace=[None if np.isnan(v) else float(v) for v in tr["chi2_new"]],
                      stride=stride, poses=P[::stride].tolist(), planes=L[::max(1, stride // 5)].tolist(), plane_stride=max(1, stride // 5),
                      node_starts=[api.node_start(int(i)) for i in list(ids["pose_ids"][:6]) + list(ids["plane_ids"][:6])])
    print(name, it, chi2_0, api.chi2())
out["runs"] = runs
path = os.path.join(ROOT, "tests", "golden", "reference_build.json")
json.dump(out, open(path, "w"))
print("wrote", path, os.path.getsize(path), "bytes")

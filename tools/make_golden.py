"""Freeze the oracle's outputs for BASELINE configs 1 and 2 as repo goldens.

The reference ships no golden vectors and cannot be built here, so these fixtures pin the ORACLE (CPU restatement
of the reference's iSAM path, numeric Jacobians as upstream) against regressions, and give the GPU parity tests a
reference that does not need the oracle binary.  Regenerate with:  python tools/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle_api import OracleAPI  # noqa: E402
from pop_up_slam_b200 import graphgen as gg  # noqa: E402

out = {}
for cfg, seed in [(1, 0), (1, 1), (2, 0), (2, 3)]:
    g = gg.make_config(cfg, seed=seed)
    api = OracleAPI()
    api.set_jacobian_mode(0)
    ids = gg.build_interleaved(api, g)
    gg.configure(api, g)
    chi2_0 = api.chi2()
    it = api.batch_optimize()
    tr = api.trace()
    out[f"config{cfg}_seed{seed}"] = dict(
        config=cfg, seed=seed, dims=g.dims(), chi2_initial=chi2_0, chi2_final=api.chi2(), iterations=it,
        accepted=tr["accepted"].tolist(), chi2_trace=tr["chi2_new"].tolist(), lambda_trace=tr["lam"].tolist(),
        pose_ids=ids["pose_ids"].tolist(), plane_ids=ids["plane_ids"].tolist(),
        node_starts=[api.node_start(int(i)) for i in list(ids["pose_ids"][:8]) + list(ids["plane_ids"][:8])],
        factor_rows=[api.factor_row(int(f)) for f in ids["pp_fids"][:16]],
        poses=api.get_poses(ids["pose_ids"]).tolist(), planes=api.get_planes(ids["plane_ids"]).tolist())
    print(cfg, seed, it, chi2_0, api.chi2())
path = os.path.join(ROOT, "tests", "golden", "oracle_configs_1_2.json")
json.dump(out, open(path, "w"))
print("wrote", path, os.path.getsize(path), "bytes")

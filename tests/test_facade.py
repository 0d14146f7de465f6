"""The header-only C++ facade (include/isam_facade.hpp) that lets pop_planar_slam's Mapping.cpp call sites compile
against the GPU library: it must compile and link on the CPU box, fail loudly without a GPU, and -- on the GPU --
reproduce the oracle when replaying a Mapper_mono::processFrame-style frame sequence."""
import os
import subprocess

import numpy as np
import pytest

from oracle_api import OracleAPI
from pop_up_slam_b200 import geometry as geo, graphgen as gg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "facade_replay")


def build_replay(eigen=False):
    """eigen=True: the -DISAM_FACADE_USE_EIGEN flavour of include/isam_facade.hpp (the one INTEGRATION.md tells a maintainer to
    use), compiled against the Eigen API stand-in of oracle/ref_shim -- Eigen itself is not installed in the build container."""
    src = os.path.join(ROOT, "tests", "facade_replay.cpp")
    libdir = os.path.join(ROOT, "pop_up_slam_b200")
    out = BIN + ("_eigen" if eigen else "")
    extra = ["-DISAM_FACADE_USE_EIGEN", "-I" + os.path.join(ROOT, "oracle", "ref_shim")] if eigen else []
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall"] + extra + ["-o", out, src, "-L" + libdir, "-lpopup_gpu",
                           "-Wl,-rpath," + libdir])
    return out


def write_frames(g, path, sig_of_edge):
    order = np.argsort(g.pp_pose, kind="stable")
    ptr = np.searchsorted(g.pp_pose[order], np.arange(g.n_poses + 1))
    with open(path, "w") as f:
        f.write(f"{g.n_poses} {g.n_planes} {g.ground_plane}\n")
        f.write(" ".join(["2.0"] * 6) + " 0.05\n")
        for i in range(g.n_poses):
            o = g.prior_meas if i == 0 else g.odo_meas[i - 1]
            f.write(" ".join(repr(float(x)) for x in o) + "\n")
            es = order[ptr[i]:ptr[i + 1]]
            f.write(f"{len(es)}\n")
            for e in es:
                f.write(f"{int(g.pp_plane[e])} " + " ".join(repr(float(x)) for x in g.pp_meas[e]) + f" {float(sig_of_edge[e])!r}\n")


def test_facade_compiles_and_links():
    assert os.path.exists(build_replay())


def _no_gpu():
    import torch
    return not torch.cuda.is_available()


@pytest.mark.skipif(not _no_gpu(), reason="a CUDA device is present")
def test_facade_fails_loudly_without_gpu(tmp_path):
    build_replay()
    g = gg.make_config(1, seed=0)
    path = tmp_path / "frames.txt"
    write_frames(g, path, np.full(g.n_pose_plane, 9.0))
    r = subprocess.run([BIN, str(path)], capture_output=True, text=True)
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)


def test_facade_eigen_flavour_compiles_and_fails_loudly_without_a_gpu(tmp_path):
    """the ISAM_FACADE_USE_EIGEN branch builds (against the Eigen API stand-in) and, like the built-in flavour, has no CPU fallback"""
    import ctypes
    exe = build_replay(eigen=True)
    try:
        ctypes.CDLL("libcuda.so.1")
        has_gpu = subprocess.run(["nvidia-smi", "-L"], capture_output=True).returncode == 0
    except OSError:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present: covered by test_facade_replay_matches_oracle[eigen]")
    g = gg.make_config(1, seed=0, sigma_mode="reference")
    path = tmp_path / "frames.txt"
    write_frames(g, path, (1.0 / g.pp_sqrtinf[:, 0]).tolist())
    r = subprocess.run([exe, str(path)], capture_output=True, text=True)
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)


@pytest.mark.gpu
@pytest.mark.parametrize("eigen", [False, True], ids=["builtin", "eigen"])
def test_facade_replay_matches_oracle(tmp_path, eigen):
    BIN = build_replay(eigen)
    g = gg.make_config(2, seed=6, n_poses=30, n_planes=10, sigma_mode="reference")
    sig = (1.0 / g.pp_sqrtinf[:, 0])
    path = tmp_path / "frames.txt"
    write_frames(g, path, sig.tolist())
    r = subprocess.run([BIN, str(path)], capture_output=True, text=True, check=True)
    chi2 = None
    poses, planes = [], []
    for line in r.stdout.splitlines():
        tok = line.split()
        if tok[0] == "chi2":
            chi2 = float(tok[1])
        elif tok[0] == "pose":
            poses.append([float(x) for x in tok[1:]])
        elif tok[0] == "plane":
            planes.append([float(x) for x in tok[1:]])
    # the same sequence through the oracle (Mapping.cpp:31-43 properties, batch every 5th frame, update otherwise)
    o = OracleAPI()
    o.set_jacobian_mode(1)
    o.set_properties(**dict(gg.PPS_PROPERTIES))
    order = np.argsort(g.pp_pose, kind="stable")
    ptr = np.searchsorted(g.pp_pose[order], np.arange(g.n_poses + 1))
    pose_ut = gg.diag_ut([0.5] * 6)
    pid, lid = [], {}
    for i in range(g.n_poses):
        pid.append(o.add_pose(None))
        if i == 0:
            o.add_pose_prior(pid[0], g.prior_meas, pose_ut)
        else:
            o.add_odometry(pid[i - 1], pid[i], g.odo_meas[i - 1], pose_ut)
        for e in order[ptr[i]:ptr[i + 1]]:
            k = int(g.pp_plane[e])
            if k not in lid:
                lid[k] = o.add_plane(None)
                o.init_plane(lid[k], geo.plane_to_global(geo.pose7_to_T(o.get_pose(pid[i])), g.pp_meas[e]))
                if k == g.ground_plane:
                    o.add_plane_prior(lid[k], [0, 0, -1, 0], gg.diag_ut([20.0] * 3))
            o.add_pose_plane(pid[i], lid[k], g.pp_meas[e], gg.diag_ut([1.0 / sig[e]] * 3))
        if i % 5 == 0:
            o.batch_optimize()
        else:
            o.update()
    c_o = o.chi2()
    assert abs(chi2 - c_o) <= 1e-4 * max(c_o, 1e-9)
    P_o = o.get_poses(pid)
    P = np.array(poses)
    assert np.abs(P[:, :3] - P_o[:, :3]).max() < 1e-4 * max(1.0, np.abs(P_o[:, :3]).max())
    L_o = o.get_planes([lid[k] for k in sorted(lid, key=lambda k: lid[k])])
    L = np.array(planes)
    sgn = np.sign(np.sum(L * L_o, axis=1))[:, None]
    assert np.abs(L * sgn - L_o).max() < 1e-4

"""CPU tests: the __host__ __device__ math the CUDA kernels run (csrc/pus_math.cuh, compiled for the
host through tests/hostmath_shim.cpp) against the oracle's restatement of the reference formulas."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle_api import OracleAPI
import oracle_api as O
from pop_up_slam_b200.capi import _dp
from pop_up_slam_b200 import geometry as geo, graphgen as gg

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hm():
    src = os.path.join(HERE, "hostmath_shim.cpp")
    out = os.path.join(HERE, "libhostmath.so")
    hdr = os.path.join(HERE, "..", "pop_up_slam_b200", "csrc", "pus_math.cuh")
    if not os.path.exists(out) or max(os.path.getmtime(src), os.path.getmtime(hdr)) > os.path.getmtime(out):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", "-o", out, src])
    lib = C.CDLL(out)
    dbl = C.c_double
    P = C.POINTER(C.c_double)
    lib.hm_pose_plane_linearize.argtypes = [P, P, P, P, C.c_int, dbl, P, P, P]
    lib.hm_pose_plane2_linearize.argtypes = [P, P, P, P, C.c_int, dbl, P, P, P]
    lib.hm_plane_prior_linearize.argtypes = [P, P, P, C.c_int, dbl, P, P]
    lib.hm_pose_plane_residual.argtypes = [P, P, P, P, C.c_int, dbl, P]
    lib.hm_pose_factor_linearize.argtypes = [P, P, P, P, C.c_int, dbl, P, P, P]
    lib.hm_pose_factor_residual.argtypes = [P, P, P, P, C.c_int, dbl, P]
    return lib


def rand_pose(rng, scale=5.0):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    return np.concatenate([rng.uniform(-scale, scale, 3), q])


def rand_pose_level(rng, scale=5.0):
    """pose with |pitch| < ~60deg (Euler-difference factors are singular at +-90deg)."""
    v = np.concatenate([rng.uniform(-scale, scale, 3), [rng.uniform(-3, 3), rng.uniform(-1.0, 1.0), rng.uniform(-3, 3)]])
    return O.pose_from_xyzypr(v)


def rand_plane(rng):
    n = rng.normal(size=3)
    n /= np.linalg.norm(n)
    return geo.plane_normalize(np.append(n, -rng.uniform(0.5, 8.0)))


def ut(rng, n):
    A = rng.uniform(0.5, 2.0, size=(n, n))
    A = np.triu(A)
    A[np.diag_indices(n)] = rng.uniform(1.0, 30.0, n)
    return A[np.triu_indices(n)]


@pytest.mark.parametrize("robust", [(0, 1.0), (1, 1.0), (2, 0.5)])
def test_pose_plane_matches_oracle(hm, robust):
    rng = np.random.default_rng(1)
    rk, rb = robust
    for trial in range(200):
        pose, plane = rand_pose(rng), rand_plane(rng)
        T = geo.pose7_to_T(pose)
        meas = geo.plane_exmap(geo.plane_to_local(T, plane), rng.normal(0, 0.3 if trial % 2 else 0.01, 3))
        sinf = ut(rng, 3)
        api = OracleAPI()
        api.set_robust(rk, rb)
        pid, lid = api.add_pose(pose), api.add_plane(plane)
        fid = api.add_pose_plane(pid, lid, meas, sinf)
        Ha, ra = api.factor_jacobian(fid, 1)
        Hn, rn = api.factor_jacobian(fid, 0)
        r, Jp, Jl = np.zeros(3), np.zeros(18), np.zeros(9)
        hm.hm_pose_plane_linearize(_dp(pose), _dp(plane), _dp(meas), _dp(sinf), rk, rb, _dp(r), _dp(Jp), _dp(Jl))
        J = np.hstack([Jp.reshape(3, 6), Jl.reshape(3, 3)])
        scale = max(1.0, np.abs(Ha).max())
        assert np.allclose(r, ra, atol=1e-12 * max(1, np.abs(ra).max())), (r, ra)
        assert np.abs(J - Ha).max() <= 1e-10 * scale
        # the reference's eps=1e-4 central differences agree to their truncation error
        if rk == 0:  # (central differences straddle the Huber kink otherwise)
            assert np.abs(J - Hn).max() <= 1e-4 * scale
        r2 = np.zeros(3)
        hm.hm_pose_plane_residual(_dp(pose), _dp(plane), _dp(meas), _dp(sinf), rk, rb, _dp(r2))
        assert np.array_equal(r, r2)


def test_plane_prior_matches_oracle(hm):
    rng = np.random.default_rng(2)
    for _ in range(100):
        plane = rand_plane(rng)
        meas = geo.plane_exmap(plane, rng.normal(0, 0.2, 3))
        sinf = ut(rng, 3)
        api = OracleAPI()
        lid = api.add_plane(plane)
        fid = api.add_plane_prior(lid, meas, sinf)
        Ha, ra = api.factor_jacobian(fid, 1)
        r, Jl = np.zeros(3), np.zeros(9)
        hm.hm_plane_prior_linearize(_dp(plane), _dp(meas), _dp(sinf), 0, 1.0, _dp(r), _dp(Jl))
        assert np.allclose(r, ra, atol=1e-12)
        assert np.abs(Jl.reshape(3, 3) - Ha).max() <= 1e-10 * max(1.0, np.abs(Ha).max())


@pytest.mark.parametrize("robust", [(0, 1.0), (1, 1.0)])
def test_odometry_and_prior_match_oracle(hm, robust):
    rng = np.random.default_rng(3)
    rk, rb = robust
    for _ in range(200):
        p1 = rand_pose_level(rng)
        rel = O.pose_from_xyzypr(np.concatenate([rng.uniform(-1, 1, 3), rng.uniform(-0.5, 0.5, 3)]))
        p2 = O.pose_oplus(p1, rel)
        meas = O.pose_vector(rel) + rng.normal(0, 0.05, 6)
        sinf = ut(rng, 6)
        api = OracleAPI()
        api.set_robust(rk, rb)
        a, b = api.add_pose(p1), api.add_pose(p2)
        fid = api.add_odometry(a, b, meas, sinf)
        Ha, ra = api.factor_jacobian(fid, 1)
        Hn, rn = api.factor_jacobian(fid, 0)
        r, J1, J2 = np.zeros(6), np.zeros(36), np.zeros(36)
        hm.hm_pose_factor_linearize(_dp(p1), _dp(p2), _dp(meas), _dp(sinf), rk, rb, _dp(r), _dp(J1), _dp(J2))
        J = np.hstack([J1.reshape(6, 6), J2.reshape(6, 6)])
        scale = max(1.0, np.abs(Ha).max())
        assert np.allclose(r, rn, atol=1e-10 * max(1, np.abs(rn).max()))   # residual vs the reference path (matrix->quat->euler)
        assert np.abs(J - Ha).max() <= 1e-9 * scale
        if rk == 0:
            assert np.abs(J - Hn).max() <= 1e-4 * scale
        # prior
        fid2 = api.add_pose_prior(a, meas, sinf)
        Hp, rp = api.factor_jacobian(fid2, 1)
        r3, J3 = np.zeros(6), np.zeros(36)
        hm.hm_pose_factor_linearize(_dp(p1), None, _dp(meas), _dp(sinf), rk, rb, _dp(r3), _dp(J3), None)
        assert np.allclose(r3, rp, atol=1e-12 * max(1, np.abs(rp).max()))
        assert np.abs(J3.reshape(6, 6) - Hp).max() <= 1e-10 * max(1.0, np.abs(Hp).max())


def test_exmaps_match_oracle(hm):
    rng = np.random.default_rng(4)
    for i in range(200):
        p = rand_pose(rng)
        d = rng.normal(0, 1e-5 if i % 3 == 0 else 0.3, 6)
        out = np.zeros(7)
        hm.hm_pose_exmap(_dp(p), _dp(d), _dp(out))
        assert np.allclose(out, O.pose_exmap(p, d), atol=1e-15)
        pl = rand_plane(rng)
        d3 = rng.normal(0, 1e-9 if i % 3 == 0 else 0.3, 3)
        o4 = np.zeros(4)
        hm.hm_plane_exmap(_dp(pl), _dp(d3), _dp(o4))
        assert np.allclose(o4, O.plane_exmap(pl, d3), atol=1e-15)
        a, b = rand_pose(rng), rand_pose(rng)
        o7 = np.zeros(7)
        hm.hm_pose_oplus(_dp(a), _dp(b), _dp(o7))
        assert np.allclose(o7, O.pose_oplus(a, b), atol=1e-14)
        hm.hm_pose_ominus(_dp(a), _dp(b), _dp(o7))
        assert np.allclose(o7, O.pose_ominus(a, b), atol=1e-14)


def test_pose_plane_factor2_matches_oracle(hm):
    """Pose3d_Plane3d_Factor2 (measurement re-popped from two ground-edge rays inside the residual): the closed-form
    pose Jacobian of the product header (incl. d measurement / d pose) against the oracle's numericalDiff of its own,
    independent restatement of isam_plane3d.h:375-419 / isam_plane3d.cpp:20-55."""
    rng = np.random.default_rng(5)
    for trial in range(150):
        R = geo.euler_to_R(rng.uniform(-3, 3), rng.uniform(-0.15, 0.15), rng.uniform(-0.15, 0.15)) @ gg.R_BASE
        T = np.eye(4); T[:3, :3] = R; T[:3, 3] = [rng.uniform(-3, 3), rng.uniform(-3, 3), rng.uniform(0.8, 1.8)]
        pose = geo.T_to_pose7(T)
        th = rng.uniform(-np.pi, np.pi)
        wall = geo.plane_normalize([np.cos(th), np.sin(th), 0.0, -rng.uniform(2.0, 6.0)])   # a vertical wall
        plane = geo.plane_exmap(wall, rng.normal(0, 0.05, 3))
        meas = geo.plane_exmap(geo.plane_to_local(T, wall), rng.normal(0, 0.02, 3))
        rays = gg.rays_from_measurement(T, meas)
        sinf = ut(rng, 3)
        api = OracleAPI()
        pid, lid = api.add_pose(pose), api.add_plane(plane)
        fid = api.add_pose_plane2(pid, lid, meas, rays, sinf)
        Hn, rn = api.factor_jacobian(fid, 0)
        r, Jp, Jl = np.zeros(3), np.zeros(18), np.zeros(9)
        hm.hm_pose_plane2_linearize(_dp(pose), _dp(plane), _dp(rays), _dp(sinf), 0, 1.0, _dp(r), _dp(Jp), _dp(Jl))
        J = np.hstack([Jp.reshape(3, 6), Jl.reshape(3, 3)])
        scale = max(1.0, np.abs(Hn).max())
        assert np.allclose(r, rn, atol=1e-11 * max(1, np.abs(rn).max())), (r, rn)
        assert np.abs(J - Hn).max() <= 2e-5 * scale, (trial, np.abs(J - Hn).max(), scale)

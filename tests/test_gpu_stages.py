"""GPU parity tests, stage by stage: every kernel phase of the LM loop is run through the debug hooks of the
C-ABI and compared with the oracle (CPU restatement of the reference) on the same graph."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spl

from oracle_api import OracleAPI
from pop_up_slam_b200 import graphgen as gg
from pop_up_slam_b200.capi import GpuGraphAPI

pytestmark = pytest.mark.gpu

NB, SP = 16, 16


def make_pair(g, robust=True):
    gpu, orc = GpuGraphAPI(), OracleAPI()
    orc.set_jacobian_mode(1)
    ig, io = gg.build_bulk(gpu, g), gg.build_bulk(orc, g)
    gg.configure(gpu, g)
    gg.configure(orc, g)
    if not robust:
        gpu.set_robust(0, 1.0)
        orc.set_robust(0, 1.0)
    return gpu, orc, ig, io


def small_graph(seed=0, **kw):
    args = dict(n_poses=150, n_planes=30, obs_per_pose=6, step=0.1, outlier_frac=0.03, robust_kind=1, robust_b=1.0,
                odo_noise=(0.003, 0.0005), aisle=10.0, radius=1.5)
    args.update(kw)
    return gg.make_corridor(seed=seed, **args)


def relerr(a, b):
    return np.abs(a - b).max() / max(1e-300, np.abs(b).max())


@pytest.mark.parametrize("robust", [False, True])
def test_linearize_and_assembly(robust):
    g = small_graph()
    gpu, orc, ig, io = make_pair(g, robust)
    N, M = g.n_poses, g.n_planes
    A, b = orc.normal_equations(0.0)
    A = A.toarray()
    gpu.upload()
    gpu.debug_run_stage(0)
    Hpp = gpu.debug_fetch("Hpp", N * 36).reshape(N, 6, 6)
    gp = gpu.debug_fetch("gp", N * 6).reshape(N, 6)
    Hll = gpu.debug_fetch("Hll", M * 9).reshape(M, 3, 3)
    gl = gpu.debug_fetch("gl", M * 3).reshape(M, 3)
    dims = gpu.debug_fetch("dims", 12).astype(int)
    nslot, ntile_pl = dims[10], dims[11]
    W = gpu.debug_fetch("W", nslot * 18).reshape(-1, 6, 3)            # pose-major slots (block-padded)
    Wt = gpu.debug_fetch("Wt", ntile_pl * 32 * 18).reshape(-1, 6, 3)  # plane-major slots
    pp_pose = gpu.debug_fetch("pp_pose", nslot).astype(int)
    pp_plane = gpu.debug_fetch("pp_plane", nslot).astype(int)
    pl2pm = gpu.debug_fetch("pl2pm", 1 << 20).astype(int)
    for p in range(N):
        assert relerr(Hpp[p], A[6 * p:6 * p + 6, 6 * p:6 * p + 6]) < 1e-10
    assert relerr(gp.reshape(-1), -b[:6 * N]) < 1e-10
    for l in range(M):
        s = 6 * N + 3 * l
        assert relerr(Hll[l], A[s:s + 3, s:s + 3]) < 1e-10
    assert relerr(gl.reshape(-1), -b[6 * N:]) < 1e-10
    assert (pp_pose >= 0).sum() == g.n_pose_plane
    for e in range(nslot):   # each (pose, plane) pair is observed once in the generator
        p, l = pp_pose[e], pp_plane[e]
        if p < 0:
            assert np.all(W[e] == 0)
            continue
        ref = A[6 * p:6 * p + 6, 6 * N + 3 * l:6 * N + 3 * l + 3]
        assert np.abs(W[e] - ref).max() <= 1e-10 * max(1.0, np.abs(ref).max())
    assert np.array_equal(Wt[:g.n_pose_plane], W[pl2pm[:g.n_pose_plane]])


def schur_reference(orc, g, lam):
    N, M = g.n_poses, g.n_planes
    A, b = orc.normal_equations(lam)
    A = A.tocsc()
    n_p = 6 * N
    App, Apl, All = A[:n_p, :n_p], A[:n_p, n_p:], A[n_p:, n_p:]
    S = App.toarray() - (Apl @ spl.spsolve(All.tocsc(), Apl.T.tocsc())).toarray()
    return A, b, S


def test_schur_operator_and_preconditioner():
    g = small_graph(seed=1)
    gpu, orc, ig, io = make_pair(g)
    N, M = g.n_poses, g.n_planes
    lam = 1e-3
    A, b, S = schur_reference(orc, g, lam)
    gpu.upload()
    gpu.debug_run_stage(1, lam)
    # operator
    rng = np.random.default_rng(0)
    x = rng.normal(size=6 * N)
    gpu.debug_store("pv0", x)
    gpu.debug_run_stage(3, lam)
    q = gpu.debug_fetch("q", 6 * N)
    ref = S @ x
    assert relerr(q, ref) < 1e-9
    # dense diagonal blocks
    nblk = (N + NB - 1) // NB
    packed = gpu.debug_fetch("Binv", nblk * 4656).reshape(nblk, 4656)    # upper triangles, row by row
    iu = np.triu_indices(96)
    for k in range(nblk):
        lo, hi = 6 * NB * k, min(6 * NB * (k + 1), 6 * N)
        ref = np.linalg.inv(S[lo:hi, lo:hi])
        Bk = np.zeros((96, 96))
        Bk[iu] = packed[k]
        Bk = Bk + np.triu(Bk, 1).T
        assert relerr(Bk[:hi - lo, :hi - lo], ref) < 1e-7
    # coarse Galerkin operator
    nc = (N - 1 + SP - 1) // SP + 1
    P = np.zeros((6 * N, 6 * nc))
    for p in range(N):
        c0, t = p // SP, (p % SP) / SP
        for d in range(6):
            P[6 * p + d, 6 * c0 + d] = 1 - t
            if t > 0:
                P[6 * p + d, 6 * (c0 + 1) + d] = t
    Ac = P.T @ S @ P
    ncp = (nc + 7) // 8 * 8       # A_c is padded to whole 48-wide pivot blocks (identity on the padding)
    Acinv = gpu.debug_fetch("Acinv", 36 * ncp * ncp).reshape(6 * ncp, 6 * ncp)
    assert relerr(Acinv[:6 * nc, :6 * nc], np.linalg.inv(Ac)) < 1e-7
    assert np.allclose(Acinv[6 * nc:, 6 * nc:], np.eye(6 * (ncp - nc))) and np.all(Acinv[:6 * nc, 6 * nc:] == 0)


@pytest.mark.parametrize("lam", [1e-6, 1e-2])
def test_damped_step_matches_direct_solve(lam):
    g = small_graph(seed=2)
    gpu, orc, ig, io = make_pair(g)
    N, M = g.n_poses, g.n_planes
    ref = orc.solve_step(lam)
    gpu.upload()
    gpu.debug_run_stage(2, lam)
    x = gpu.debug_fetch("x", 6 * N)
    dl = gpu.debug_fetch("dl", 3 * M)
    got = np.concatenate([x, dl])
    assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 1e-7
    st = gpu.stats()
    assert 0 < st["pcg_iterations"] < 400


def test_chi2_matches_oracle():
    g = small_graph(seed=3)
    gpu, orc, ig, io = make_pair(g)
    c_g, c_o = gpu.chi2(), orc.chi2()
    assert abs(c_g - c_o) <= 1e-11 * c_o


def hat_matrix(N, spc):
    nc = (N - 1 + spc - 1) // spc + 1
    P = np.zeros((6 * N, 6 * nc))
    for p in range(N):
        c0, t = p // spc, (p % spc) / spc
        for d in range(6):
            P[6 * p + d, 6 * c0 + d] = 1 - t
            if t > 0:
                P[6 * p + d, 6 * (c0 + 1) + d] = t
    return P, nc


def test_three_level_preconditioner_pieces():
    """the large-graph preconditioner forced on a small corridor (solver option reserved[2] bit 3): the level-2 groups are
    the inverted 96 x 96 diagonal blocks of P2^T S P2 (hat nodes every 16 poses, 16 nodes per group) and level 3 is the
    inverse of P3^T S P3 (hat nodes every 128 poses), both against numpy on the oracle's normal equations."""
    import ctypes
    g = small_graph(seed=4, n_poses=610, n_planes=60)
    gpu, orc, ig, io = make_pair(g)
    o = gpu.get_solver_options()
    o.reserved[2] = 8
    gpu._chk(gpu.lib.pus_set_solver_options(gpu.h, ctypes.byref(o)))
    N = g.n_poses
    lam = 1e-3
    A, b, S = schur_reference(orc, g, lam)
    gpu.upload()
    gpu.debug_run_stage(1, lam)
    dims = gpu.debug_fetch("dims", 18).astype(int)
    assert dims[14] == 3 and dims[12] == 128
    P2, nc2 = hat_matrix(N, 16)
    assert dims[15] == nc2
    A2 = P2.T @ S @ P2
    ng = (nc2 + 15) // 16
    D2inv = gpu.debug_fetch("D2inv", ng * 96 * 96).reshape(ng, 96, 96)
    for k in range(ng):
        lo, hi = 96 * k, min(96 * (k + 1), 6 * nc2)
        ref = np.linalg.inv(A2[lo:hi, lo:hi])
        assert relerr(D2inv[k][:hi - lo, :hi - lo], ref) < 1e-7, k
        assert np.allclose(D2inv[k][hi - lo:, hi - lo:], np.eye(96 - (hi - lo)))
    P3, nc3 = hat_matrix(N, 128)
    A3 = P3.T @ S @ P3
    ncp = (nc3 + 7) // 8 * 8
    Acinv = gpu.debug_fetch("Acinv", 36 * ncp * ncp).reshape(6 * ncp, 6 * ncp)
    assert relerr(Acinv[:6 * nc3, :6 * nc3], np.linalg.inv(A3)) < 1e-7
    # and the whole solve with it
    ref = orc.solve_step(lam)
    gpu.debug_run_stage(2, lam)
    got = np.concatenate([gpu.debug_fetch("x", 6 * N), gpu.debug_fetch("dl", 3 * g.n_planes)])
    assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 1e-7
    print("three-level PCG iterations on the 610-pose corridor:", gpu.stats()["pcg_iterations"])


def test_coarse_inverse_with_a_heavy_plane():
    """A_c^-1 once the ground plane touches more than 16 coarse nodes (420 poses) and enters the Galerkin operator as a
    dense rank-3 update: against numpy."""
    g = small_graph(seed=5, n_poses=420, n_planes=52)
    gpu, orc, ig, io = make_pair(g)
    N = g.n_poses
    lam = 1e-3
    A, b, S = schur_reference(orc, g, lam)
    gpu.upload()
    gpu.debug_run_stage(1, lam)
    P, nc = hat_matrix(N, 16)
    Ac = P.T @ S @ P
    ncp = (nc + 7) // 8 * 8
    Acinv = gpu.debug_fetch("Acinv", 36 * ncp * ncp).reshape(6 * ncp, 6 * ncp)
    assert relerr(Acinv[:6 * nc, :6 * nc], np.linalg.inv(Ac)) < 1e-7
    assert np.allclose(Acinv[6 * nc:, 6 * nc:], np.eye(6 * (ncp - nc))) and np.all(Acinv[:6 * nc, 6 * nc:] == 0)

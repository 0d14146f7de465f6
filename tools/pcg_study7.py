"""CPU study for large graphs (BASELINE config 5 structure: 20 observations per pose, ~200 observers per plane), reduced to
N poses: PCG iterations of the implicit-Schur system with (a) the two-level preconditioner at several coarse spacings (the
dense A_c^-1 caps the coarse dimension, so 50 k poses means spacing 160 today) and (b) three-level additive variants whose
spacing-16 level is solved by its own block-Jacobi plus a dense coarser level.
usage: python tools/pcg_study7.py [N=8000] [lambda=1e-4]"""
import sys, time, warnings
warnings.filterwarnings("ignore")
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spl
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
from oracle_api import OracleAPI
from pop_up_slam_b200 import graphgen as gg

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8000
lam = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-4
g = gg.make_config(5, seed=0, n_poses=N, n_planes=N // 10)
api = OracleAPI(); api.set_jacobian_mode(1)
gg.build_bulk(api, g); gg.configure(api, g)
A, b = api.normal_equations(lam)
M = g.n_planes
np_ = 6 * N
App = A[:np_, :np_].tocsr(); Apl = A[:np_, np_:].tocsr(); All = A[np_:, np_:].tocsc()
bp, bl = b[:np_], b[np_:]
Alli = sp.block_diag([sp.coo_matrix(np.linalg.inv(All[3*k:3*k+3, 3*k:3*k+3].toarray())) for k in range(M)]).tocsr()
B = (Alli @ Apl.T.tocsr()).tocsc()
rhs = bp - Apl @ (Alli @ bl)
S_mv = lambda x: App @ x - Apl @ (B @ x)
def pcg(Minv, tol=1e-8, maxit=3000):
    x = np.zeros_like(rhs); r = rhs.copy(); z = Minv(r); p = z.copy(); rz = r @ z; rz0 = rz
    for k in range(maxit):
        q = S_mv(p); alpha = rz / (p @ q); x += alpha * p; r -= alpha * q
        z = Minv(r); rz_new = r @ z
        if np.sqrt(abs(rz_new) / rz0) < tol: return k + 1
        p = z + (rz_new / rz) * p; rz = rz_new
    return maxit
def hatP(n, sp_, dof=6):
    nc = (n - 1 + sp_ - 1) // sp_ + 1
    rows, cols, vals = [], [], []
    for p in range(n):
        c0 = p // sp_; t = (p - c0 * sp_) / sp_
        for d in range(dof):
            rows.append(dof*p+d); cols.append(dof*c0+d); vals.append(1 - t)
            if t > 0 and c0 + 1 < nc:
                rows.append(dof*p+d); cols.append(dof*(c0+1)+d); vals.append(t)
    return sp.csr_matrix((vals, (rows, cols)), shape=(dof*n, dof*nc)), nc
bs = 16
nb = (N + bs - 1) // bs
blocks = []
for k in range(nb):
    lo, hi = 6*k*bs, min(6*(k+1)*bs, np_)
    Skk = App[lo:hi, lo:hi].toarray() - (Apl[lo:hi, :] @ B[:, lo:hi]).toarray()
    blocks.append(sp.coo_matrix(np.linalg.inv(Skk)))
Binv = sp.block_diag(blocks).tocsr()
t0 = time.time()
P16, nc16 = hatP(N, 16)
SP = np.column_stack([S_mv(P16[:, j].toarray().ravel()) for j in range(P16.shape[1])])
A16 = np.asarray(P16.T @ SP); A16 = 0.5 * (A16 + A16.T)
print("N", N, "lambda", lam, "A16 dim", A16.shape[0], "galerkin %.1fs" % (time.time() - t0), flush=True)
bw = max(abs(i - j) for i, j in zip(*np.nonzero(np.abs(A16) > 1e-12 * np.abs(A16).max()))) 
print("A16 half-bandwidth (scalars):", bw, "=", bw / 6, "nodes", flush=True)
print("block-16 only:", pcg(lambda r: Binv @ r, maxit=1500), flush=True)
A16inv = np.linalg.inv(A16)
print("2-level hat16 exact:", pcg(lambda r: Binv @ r + P16 @ (A16inv @ (P16.T @ r))), flush=True)
for spc in (48, 96, 160):
    Pc, ncc = hatP(N, spc)
    # spacing multiple of 16: hat_spc = P16 * hat(nc16, spc/16)
    Pcc, _ = hatP(nc16, spc // 16)
    Ac = Pcc.T @ A16 @ Pcc
    Aci = np.linalg.inv(Ac)
    print("2-level hat%d exact (dim %d):" % (spc, Ac.shape[0]), pcg(lambda r: Binv @ r + P16 @ (Pcc @ (Aci @ (Pcc.T @ (P16.T @ r))))), flush=True)
for grp, cs in ((16, 8), (16, 16), (8, 8), (32, 16), (16, 4)):
    ng = (nc16 + grp - 1) // grp
    D = sp.block_diag([sp.coo_matrix(np.linalg.inv(A16[6*grp*k:6*grp*(k+1), 6*grp*k:6*grp*(k+1)])) for k in range(ng)]).tocsr()
    P2, nc2 = hatP(nc16, cs)
    A2 = P2.T @ A16 @ P2; A2inv = np.linalg.inv(A2)
    def Minv(r, D=D, P2=P2, A2inv=A2inv):
        rc = P16.T @ r
        zc = D @ rc + P2 @ (A2inv @ (P2.T @ rc))
        return Binv @ r + P16 @ zc
    print("3-level additive: hat16 in blocks of %d nodes (%d-dim) + hat%d (dim %d):" % (grp, 6*grp, 16*cs, 6*nc2), pcg(Minv), flush=True)
    # multiplicative inside the coarse level: zc = D rc + P2 A2^-1 P2^T (rc - A16 D rc)  (one extra banded mat-vec on the coarse level)
    def Minv2(r, D=D, P2=P2, A2inv=A2inv):
        rc = P16.T @ r
        z1 = D @ rc
        z2 = P2 @ (A2inv @ (P2.T @ (rc - A16 @ z1)))
        zc = z1 + z2
        zc = zc + D @ (rc - A16 @ zc)     # symmetric: post-smoothing
        return Binv @ r + P16 @ zc
    print("   with a symmetric multiplicative V-cycle on the coarse level:", pcg(Minv2), flush=True)

"""CPU study: three-level additive preconditioner (dense 16-pose blocks + hat16 level solved approximately by
its own block-Jacobi + hat128 coarse level) vs the exact hat16 / hat32 coarse solves."""
import sys, time, warnings
warnings.filterwarnings("ignore")
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spl
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
from oracle_api import OracleAPI
from pop_up_slam_b200 import graphgen as gg

g = gg.make_config(3, seed=0)
api = OracleAPI(); api.set_jacobian_mode(1)
gg.build_bulk(api, g); gg.configure(api, g)
A, b = api.normal_equations(1e-6)
N, M = g.n_poses, g.n_planes
np_ = 6 * N
App = A[:np_, :np_].tocsr(); Apl = A[:np_, np_:].tocsr(); All = A[np_:, np_:].tocsc()
bp, bl = b[:np_], b[np_:]
Alli = sp.block_diag([sp.coo_matrix(np.linalg.inv(All[3*k:3*k+3, 3*k:3*k+3].toarray())) for k in range(M)]).tocsr()
B = (Alli @ Apl.T.tocsr()).tocsc()
rhs = bp - Apl @ (Alli @ bl)
S_mv = lambda x: App @ x - Apl @ (B @ x)
x_ref = spl.spsolve(A.tocsc(), b)[:np_]
def pcg(Minv, tol=1e-10, maxit=3000):
    x = np.zeros_like(rhs); r = rhs.copy(); z = Minv(r); p = z.copy(); rz = r @ z; rz0 = rz
    for k in range(maxit):
        q = S_mv(p); alpha = rz / (p @ q); x += alpha * p; r -= alpha * q
        z = Minv(r); rz_new = r @ z
        if np.sqrt(abs(rz_new) / rz0) < tol: return k + 1, np.linalg.norm(x - x_ref) / np.linalg.norm(x_ref)
        p = z + (rz_new / rz) * p; rz = rz_new
    return maxit, -1
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
def galerkin(P):
    SP = np.column_stack([S_mv(P[:, j].toarray().ravel()) for j in range(P.shape[1])])
    return P.T @ SP
for sp1 in (16, 32):
    P1, nc1 = hatP(N, sp1)
    A1 = np.asarray(galerkin(P1)); A1 = 0.5 * (A1 + A1.T)
    A1inv = np.linalg.inv(A1)
    print("hat%d exact (dim %d): its, err ="%(sp1, 6*nc1), pcg(lambda r: Binv @ r + P1 @ (A1inv @ (P1.T @ r))))
    # three-level: A1^-1 ~= blockdiag_g(A1)^-1 + P2 A2^-1 P2^T, coarse-of-coarse spacing cs nodes
    for grp, cs in ((8, 8), (4, 8), (8, 4), (16, 8)):
        ng = (nc1 + grp - 1) // grp
        D = sp.block_diag([sp.coo_matrix(np.linalg.inv(A1[6*grp*k:6*grp*(k+1), 6*grp*k:6*grp*(k+1)])) for k in range(ng)]).tocsr()
        P2, nc2 = hatP(nc1, cs)
        A2 = P2.T @ A1 @ P2; A2inv = np.linalg.inv(A2)
        def Minv(r, D=D, P2=P2, A2inv=A2inv):
            rc = P1.T @ r
            zc = D @ rc + P2 @ (A2inv @ (P2.T @ rc))
            return Binv @ r + P1 @ zc
        print("  3-level: hat%d blocks of %d nodes (%d-dim) + hat x%d (dim %d): its, err ="%(sp1, grp, 6*grp, cs, 6*nc2), pcg(Minv))

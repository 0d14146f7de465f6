"""Scratch study (CPU, scipy): PCG iteration counts on the plane-eliminated (Schur) pose system for
candidate preconditioners, using the oracle's assembled normal equations.  Developer tool only."""
import sys, time
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spl
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
from oracle_api import OracleAPI
from pop_up_slam_b200 import graphgen as gg

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 3
lam = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-6
g = gg.make_config(cfg, seed=0)
api = OracleAPI(); api.set_jacobian_mode(1)
gg.build_bulk(api, g); gg.configure(api, g)
t0 = time.time()
A, b = api.normal_equations(lam)
print("assembled", A.shape, A.nnz, "%.1fs" % (time.time() - t0))
N, M = g.n_poses, g.n_planes
np_, nl = 6 * N, 3 * M
App = A[:np_, :np_].tocsr(); Apl = A[:np_, np_:].tocsr(); All = A[np_:, np_:].tocsc()
bp, bl = b[:np_], b[np_:]
# block-diagonal inverse of All
Alli = sp.block_diag([np.linalg.inv(All[3*k:3*k+3, 3*k:3*k+3].toarray()) for k in range(M)]).tocsr()
Alp = Apl.T.tocsr()
B = (Alli @ Alp).tocsc()          # 3M x 6N
rhs = bp - Apl @ (Alli @ bl)
def S_mv(x):
    return App @ x - Apl @ (B @ x)
S = spl.LinearOperator((np_, np_), matvec=S_mv)
# exact reference solution
xs = spl.spsolve(A.tocsc(), b)
x_ref = xs[:np_]
print("direct solved; |x|", np.linalg.norm(x_ref))
# block-Jacobi of S
t0 = time.time()
Dinv = []
Aplc = Apl.tocsr()
for p in range(N):
    rows = slice(6*p, 6*p+6)
    Sp = App[rows, rows].toarray() - (Aplc[rows, :] @ B[:, rows]).toarray()
    Dinv.append(np.linalg.inv(Sp))
Dinv = sp.block_diag(Dinv).tocsr()
print("block-jacobi built %.1fs" % (time.time() - t0))

def pcg(matvec, rhs, Minv, tol, maxit, x_ref=None):
    x = np.zeros_like(rhs); r = rhs.copy(); z = Minv(r); p = z.copy(); rz = r @ z; rz0 = rz
    hist = []
    for k in range(maxit):
        q = matvec(p); alpha = rz / (p @ q); x += alpha * p; r -= alpha * q
        z = Minv(r); rz_new = r @ z
        err = np.linalg.norm(x - x_ref) / np.linalg.norm(x_ref) if x_ref is not None else 0
        hist.append((np.sqrt(rz_new / rz0), np.linalg.norm(r) / np.linalg.norm(rhs), err))
        if np.sqrt(rz_new / rz0) < tol: break
        p = z + (rz_new / rz) * p; rz = rz_new
    return x, hist

def report(name, hist):
    h = np.array(hist)
    def first(col, t):
        idx = np.nonzero(h[:, col] < t)[0]
        return int(idx[0]) + 1 if len(idx) else None
    print(name, "iters", len(h), " M-res<1e-3:", first(0, 1e-3), " <1e-6:", first(0, 1e-6), " <1e-10:", first(0, 1e-10),
          " | x-err<1e-3:", first(2, 1e-3), " <1e-6:", first(2, 1e-6), " <1e-8:", first(2, 1e-8))

x, h = pcg(S_mv, rhs, lambda r: Dinv @ r, 1e-12, 3000, x_ref)
report("block-Jacobi", h)

# two-level additive: piecewise-constant aggregates of `agg` consecutive poses (6 dof each)
for agg in (25, 50, 100):
    nc = (N + agg - 1) // agg
    rows = np.arange(np_); cols = (rows // 6 // agg) * 6 + rows % 6
    P = sp.csr_matrix((np.ones(np_), (rows, cols)), shape=(np_, 6 * nc))
    SP = np.column_stack([S_mv(P[:, j].toarray().ravel()) for j in range(6 * nc)])
    Ac = P.T @ SP
    Aci = np.linalg.inv(Ac)
    Minv = lambda r: Dinv @ r + P @ (Aci @ (P.T @ r))
    x, h = pcg(S_mv, rhs, Minv, 1e-12, 3000, x_ref)
    report("2-level additive agg=%d (coarse dim %d)" % (agg, 6 * nc), h)

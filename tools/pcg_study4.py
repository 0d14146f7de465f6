import sys, time, warnings
warnings.filterwarnings("ignore")
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spl
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
from oracle_api import OracleAPI
from pop_up_slam_b200 import graphgen as gg

def study(g, lam, label, maxit=3000):
    api = OracleAPI(); api.set_jacobian_mode(1)
    gg.build_bulk(api, g); gg.configure(api, g)
    A, b = api.normal_equations(lam)
    N, M = g.n_poses, g.n_planes
    np_, nl = 6 * N, 3 * M
    App = A[:np_, :np_].tocsr(); Apl = A[:np_, np_:].tocsr(); All = A[np_:, np_:].tocsc()
    bp, bl = b[:np_], b[np_:]
    Alli = sp.block_diag([sp.coo_matrix(np.linalg.inv(All[3*k:3*k+3, 3*k:3*k+3].toarray())) for k in range(M)]).tocsr()
    B = (Alli @ Apl.T.tocsr()).tocsc()
    rhs = bp - Apl @ (Alli @ bl)
    S_mv = lambda x: App @ x - Apl @ (B @ x)
    x_ref = spl.spsolve(A.tocsc(), b)[:np_]
    Cd = sp.block_diag([sp.coo_matrix((Apl[6*p:6*p+6, :] @ B[:, 6*p:6*p+6]).toarray()) for p in range(N)]).tocsr()
    Dblk = sp.block_diag([sp.coo_matrix(App[6*p:6*p+6, 6*p:6*p+6].toarray()) for p in range(N)]).tocsr() - Cd
    Dinv = sp.block_diag([sp.coo_matrix(np.linalg.inv(Dblk[6*p:6*p+6, 6*p:6*p+6].toarray())) for p in range(N)]).tocsr()
    lu = spl.splu((App - Cd).tocsc())
    def pcg(Minv, tol=1e-11):
        x = np.zeros_like(rhs); r = rhs.copy(); z = Minv(r); p = z.copy(); rz = r @ z; rz0 = rz
        hist = []
        for k in range(maxit):
            q = S_mv(p); alpha = rz / (p @ q); x += alpha * p; r -= alpha * q
            z = Minv(r); rz_new = r @ z
            hist.append((np.sqrt(abs(rz_new) / rz0), np.linalg.norm(x - x_ref) / np.linalg.norm(x_ref)))
            if np.sqrt(abs(rz_new) / rz0) < tol: break
            p = z + (rz_new / rz) * p; rz = rz_new
        return np.array(hist)
    def rep(name, h):
        f = lambda c, t: (int(np.nonzero(h[:, c] < t)[0][0]) + 1) if (h[:, c] < t).any() else None
        print("  %-36s its=%5d  Mres<1e-3:%s <1e-6:%s <1e-10:%s | xerr<1e-3:%s <1e-6:%s" % (name, len(h), f(0,1e-3), f(0,1e-6), f(0,1e-10), f(1,1e-3), f(1,1e-6)))
    print(label, "N=%d M=%d" % (N, M))
    def hatP(sp_):
        nc = (N - 1 + sp_ - 1) // sp_ + 1
        rows, cols, vals = [], [], []
        for p in range(N):
            c0 = p // sp_; t = (p - c0 * sp_) / sp_
            for d in range(6):
                rows.append(6*p+d); cols.append(6*c0+d); vals.append(1 - t)
                if t > 0 and c0 + 1 < nc:
                    rows.append(6*p+d); cols.append(6*(c0+1)+d); vals.append(t)
        return sp.csr_matrix((vals, (rows, cols)), shape=(np_, 6*nc)), nc
    for sp_ in (50, 100, 200):
        P, nc = hatP(sp_)
        SP = np.column_stack([S_mv(P[:, j].toarray().ravel()) for j in range(6 * nc)])
        Aci = np.linalg.inv(P.T @ SP)
        rep("jacobi + hat coarse sp=%d (dim %d)" % (sp_, 6*nc), pcg(lambda r: Dinv @ r + P @ (Aci @ (P.T @ r))))
        rep("tridiag + hat coarse sp=%d (dim %d)" % (sp_, 6*nc), pcg(lambda r: lu.solve(r) + P @ (Aci @ (P.T @ r))))
        # multiplicative (symmetric): coarse, then tridiag on residual, then coarse
        def mult(r):
            z = P @ (Aci @ (P.T @ r))
            z = z + lu.solve(r - S_mv(z))
            z = z + P @ (Aci @ (P.T @ (r - S_mv(z))))
            return z
        rep("mult: coarse-tridiag-coarse sp=%d" % sp_, pcg(mult))

study(gg.make_config(3, seed=0), 1e-6, "C3 robust")

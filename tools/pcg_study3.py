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
    App = A[:np_, :np_].tocsc(); Apl = A[:np_, np_:].tocsr(); All = A[np_:, np_:].tocsr()
    Alp = Apl.T.tocsr()
    bp, bl = b[:np_], b[np_:]
    xs = spl.spsolve(A.tocsc(), b)
    xl_ref = xs[np_:]
    lu = spl.splu(App)
    rhs = bl - Alp @ lu.solve(bp)
    Sl_mv = lambda x: All @ x - Alp @ lu.solve(Apl @ x)
    # exact block diag of S_l (study only)
    t0=time.time()
    X = lu.solve(Apl.toarray()) if nl <= 2000 else None
    if X is not None:
        Sl = All.toarray() - Alp @ X
        ev = np.linalg.eigvalsh(Sl)
        print("  S_l eig min/max", ev[0], ev[-1], "cond %.3g" % (ev[-1]/ev[0]))
        Dinv = sp.block_diag([sp.coo_matrix(np.linalg.inv(Sl[3*k:3*k+3, 3*k:3*k+3])) for k in range(M)]).tocsr()
        Dl = np.sqrt(np.diag(Sl)); ev2 = np.linalg.eigvalsh(Sl / Dl[:,None] / Dl[None,:]); print("  diag-scaled cond %.3g" % (ev2[-1]/ev2[0]))
    Alli = sp.block_diag([sp.coo_matrix(np.linalg.inv(All[3*k:3*k+3, 3*k:3*k+3].toarray())) for k in range(M)]).tocsr()
    def pcg(Minv, tol=1e-11):
        x = np.zeros_like(rhs); r = rhs.copy(); z = Minv(r); p = z.copy(); rz = r @ z; rz0 = rz
        hist = []
        for k in range(maxit):
            q = Sl_mv(p); alpha = rz / (p @ q); x += alpha * p; r -= alpha * q
            z = Minv(r); rz_new = r @ z
            hist.append((np.sqrt(abs(rz_new) / rz0), np.linalg.norm(x - xl_ref) / np.linalg.norm(xl_ref)))
            if np.sqrt(abs(rz_new) / rz0) < tol: break
            p = z + (rz_new / rz) * p; rz = rz_new
        return np.array(hist)
    def rep(name, h):
        f = lambda c, t: (int(np.nonzero(h[:, c] < t)[0][0]) + 1) if (h[:, c] < t).any() else None
        print("  %-28s its=%5d  Mres<1e-3:%s <1e-6:%s <1e-10:%s | xerr<1e-3:%s <1e-6:%s" % (name, len(h), f(0,1e-3), f(0,1e-6), f(0,1e-10), f(1,1e-3), f(1,1e-6)))
    print(label, "N=%d M=%d" % (N, M))
    rep("plane-space, All^-1 precond", pcg(lambda r: Alli @ r))
    if X is not None: rep("plane-space, diag(S_l)^-1", pcg(lambda r: Dinv @ r))

study(gg.make_config(2, seed=0), 1e-6, "C2")
study(gg.make_config(3, seed=0), 1e-6, "C3 robust")
study(gg.make_config(3, seed=0, robust_kind=0, outlier_frac=0.0), 1e-6, "C3 plain LS")

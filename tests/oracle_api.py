"""Test-side binding of the CPU oracle (oracle/liboracle.so, orc_* entry points).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use
this module: the oracle is the checker, never the product path.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from pop_up_slam_b200.capi import GraphAPI, _dp, _ip, c_double_p, c_int_p

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "liboracle.so")

_lib = None


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR])


def oracle_lib():
    global _lib
    if _lib is None:
        srcs = [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR) if f.endswith((".hpp", ".cpp"))]
        if not os.path.exists(ORACLE_LIB) or any(os.path.getmtime(s) > os.path.getmtime(ORACLE_LIB) for s in srcs):
            build_oracle()
        _lib = C.CDLL(ORACLE_LIB)
        _lib.orc_standard_rad.restype = C.c_double
        _lib.orc_standard_rad.argtypes = [C.c_double]
        _lib.orc_solve_step.argtypes = [C.c_void_p, C.c_double, c_double_p, C.c_int]
        _lib.orc_normal_equations.restype = C.c_longlong
        _lib.orc_normal_equations.argtypes = [C.c_void_p, C.c_double, c_int_p, c_int_p, c_double_p, c_double_p, C.c_longlong]
    return _lib


class OracleAPI(GraphAPI):
    def __init__(self):
        super().__init__(oracle_lib(), "orc_", 0)

    def set_jacobian_mode(self, mode):
        """0 = numeric central differences (the reference's numericalDiff), 1 = analytic."""
        self.lib.orc_set_jacobian_mode(self.h, int(mode))

    def set_reuse_ordering(self, on):
        self.lib.orc_set_reuse_ordering(self.h, int(on))

    def timers(self):
        out = np.zeros(8)
        self.lib.orc_get_timers(self.h, _dp(out))
        return dict(linearize=out[0], solve=out[1], chi2=out[2], order=out[3], total=out[4],
                    n_linearize=int(out[5]), n_solve=int(out[6]), n_chi2=int(out[7]))

    def reset_timers(self):
        self.lib.orc_reset_timers(self.h)

    def factor_jacobian(self, fid, mode):
        H = np.zeros(6 * 12)
        r = np.zeros(6)
        ncols = self._chk(self.lib.orc_factor_jacobian(self.h, int(fid), int(mode), _dp(H), _dp(r)))
        dim = 3 if ncols in (3, 9) else 6
        return H[:dim * ncols].reshape(dim, ncols).copy(), r[:dim].copy()

    def factor_error(self, fid):
        r = np.zeros(6)
        d = self._chk(self.lib.orc_factor_error(self.h, int(fid), _dp(r)))
        return r[:d].copy()

    def normal_equations(self, lam):
        import scipy.sparse as sp
        nnz = self.lib.orc_normal_equations(self.h, float(lam), None, None, None, None, 0)
        n = self.state_dim()
        Ap = np.zeros(n + 1, dtype=np.int32)
        Ai = np.zeros(nnz, dtype=np.int32)
        Ax = np.zeros(nnz)
        b = np.zeros(n)
        self.lib.orc_normal_equations(self.h, float(lam), _ip(Ap), _ip(Ai), _dp(Ax), _dp(b), nnz)
        U = sp.csc_matrix((Ax, Ai, Ap), shape=(n, n))
        A = U + sp.triu(U, 1).T
        return A.tocsc(), b

    def state_dim(self):
        d = np.zeros(1)
        return self.lib.orc_solve_step(self.h, 0.0, _dp(d), 0)

    def solve_step(self, lam):
        n = self.state_dim()
        d = np.zeros(n)
        self.lib.orc_solve_step(self.h, float(lam), _dp(d), n)
        return d

    def apply_delta(self, delta):
        d = np.ascontiguousarray(delta, dtype=np.float64)
        self.lib.orc_apply_delta(self.h, _dp(d), len(d))


def _call7(fn, *args):
    lib = oracle_lib()
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in args]
    return lib, arrs


def pose_from_xyzypr(v):
    out = np.zeros(7)
    oracle_lib().orc_pose_from_xyzypr(_dp(np.ascontiguousarray(v, dtype=np.float64)), _dp(out))
    return out


def pose_vector(p7):
    out = np.zeros(6)
    oracle_lib().orc_pose_vector(_dp(np.ascontiguousarray(p7, dtype=np.float64)), _dp(out))
    return out


def pose_exmap(p7, d6):
    out = np.zeros(7)
    oracle_lib().orc_pose_exmap(_dp(np.ascontiguousarray(p7, dtype=np.float64)), _dp(np.ascontiguousarray(d6, dtype=np.float64)), _dp(out))
    return out


def pose_oplus(a7, b7):
    out = np.zeros(7)
    oracle_lib().orc_pose_oplus(_dp(np.ascontiguousarray(a7, dtype=np.float64)), _dp(np.ascontiguousarray(b7, dtype=np.float64)), _dp(out))
    return out


def pose_ominus(a7, b7):
    out = np.zeros(7)
    oracle_lib().orc_pose_ominus(_dp(np.ascontiguousarray(a7, dtype=np.float64)), _dp(np.ascontiguousarray(b7, dtype=np.float64)), _dp(out))
    return out


def pose_wTo(p7):
    out = np.zeros(16)
    oracle_lib().orc_pose_wTo(_dp(np.ascontiguousarray(p7, dtype=np.float64)), _dp(out))
    return out.reshape(4, 4)


def pose_oTw(p7):
    out = np.zeros(16)
    oracle_lib().orc_pose_oTw(_dp(np.ascontiguousarray(p7, dtype=np.float64)), _dp(out))
    return out.reshape(4, 4)


def pose_from_mat4(T):
    out = np.zeros(7)
    oracle_lib().orc_pose_from_mat4(_dp(np.ascontiguousarray(T, dtype=np.float64).reshape(-1)), _dp(out))
    return out


def plane_exmap(p4, d3):
    out = np.zeros(4)
    oracle_lib().orc_plane_exmap(_dp(np.ascontiguousarray(p4, dtype=np.float64)), _dp(np.ascontiguousarray(d3, dtype=np.float64)), _dp(out))
    return out


def plane_transform(T, p4):
    out = np.zeros(4)
    oracle_lib().orc_plane_transform(_dp(np.ascontiguousarray(T, dtype=np.float64).reshape(-1)), _dp(np.ascontiguousarray(p4, dtype=np.float64)), _dp(out))
    return out


def plane_log_error(l4, m4):
    out = np.zeros(3)
    oracle_lib().orc_plane_log_error(_dp(np.ascontiguousarray(l4, dtype=np.float64)), _dp(np.ascontiguousarray(m4, dtype=np.float64)), _dp(out))
    return out

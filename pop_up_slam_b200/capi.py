"""ctypes binding of the C-ABI declared in include/popup_gpu.h.

`GraphAPI` is a thin, prefix-generic wrapper: the product binds it to libpopup_gpu.so with the
`pus_` prefix (see `load_library` / `Solver` in slam.py); the test-suite binds the same class to the CPU
oracle's `orc_` entry points, so parity tests drive both through identical calls.  Nothing here
imports or loads anything under oracle/.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpopup_gpu.so")

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
c_float_p = C.POINTER(C.c_float)


class Properties(C.Structure):
    """pus_properties (isam::Properties, Properties.h:37-110)."""
    _fields_ = [("method", C.c_int), ("epsilon2", C.c_double), ("epsilon_abs", C.c_double),
                ("epsilon_rel", C.c_double), ("max_iterations", C.c_int), ("lm_lambda0", C.c_double),
                ("lm_lambda_factor", C.c_double), ("mod_update", C.c_int), ("mod_batch", C.c_int),
                ("mod_solve", C.c_int)]


class SolverOptions(C.Structure):
    _fields_ = [("pcg_rel_tol", C.c_double), ("pcg_max_iter", C.c_int), ("ctas_per_sm", C.c_int),
                ("team_ctas", C.c_int), ("reserved", C.c_int * 4)]


class Stats(C.Structure):
    _fields_ = [("lm_iterations", C.c_int), ("accepted", C.c_int), ("relinearizations", C.c_int),
                ("chi2_evals", C.c_int), ("pcg_iterations", C.c_longlong), ("chi2_initial", C.c_double),
                ("chi2_final", C.c_double), ("kernel_ms", C.c_double), ("h2d_ms", C.c_double),
                ("d2h_ms", C.c_double), ("h2d_bytes", C.c_longlong), ("d2h_bytes", C.c_longlong),
                ("n_poses", C.c_int), ("n_planes", C.c_int), ("n_pose_plane", C.c_int), ("n_odometry", C.c_int),
                ("n_pose_prior", C.c_int), ("n_plane_prior", C.c_int), ("gpu_launches", C.c_int),
                ("grid_ctas", C.c_int), ("block_threads", C.c_int), ("phase_ms", C.c_double * 24)]

    def as_dict(self):
        d = {}
        for name, _ in self._fields_:
            v = getattr(self, name)
            d[name] = list(v) if hasattr(v, "__len__") else v
        return d


def _dp(a):
    return a.ctypes.data_as(c_double_p)


def _ip(a):
    return a.ctypes.data_as(c_int_p)


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        assert a.shape == shape, (a.shape, shape)
    return a


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class ApiError(RuntimeError):
    pass


def load_library(path=None):
    """Load libpopup_gpu.so. Fails loudly when the CUDA library has not been built -- there is no
    CPU fallback behind this ABI."""
    path = path or os.environ.get("PUS_LIBRARY", LIB_PATH)   # (PUS_LIBRARY: A/B builds of the same ABI)
    if not os.path.exists(path):
        raise ApiError(f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(nvcc, sm_100a). There is no CPU fallback.")
    return C.CDLL(path)


class GraphAPI:
    """One solver handle behind the C-ABI (prefix `pus_` = CUDA library, `orc_` = test oracle)."""

    def __init__(self, lib, prefix="pus_", device=0):
        self.lib = lib
        self.prefix = prefix
        self._bind()
        h = C.c_void_p()
        self._chk(self._f("create")(int(device), C.byref(h)))
        self.h = h

    # -- plumbing --
    def _f(self, name):
        return getattr(self.lib, self.prefix + name)

    def _bind(self):
        f = self._f
        f("last_error").restype = C.c_char_p
        for name in ("debug_fetch", "debug_store", "normal_equations"):
            if hasattr(self.lib, self.prefix + name):
                f(name).restype = C.c_longlong
        if hasattr(self.lib, self.prefix + "set_robust"):
            f("set_robust").argtypes = [C.c_void_p, C.c_int, C.c_double]
        if hasattr(self.lib, self.prefix + "debug_run_stage"):
            f("debug_run_stage").argtypes = [C.c_void_p, C.c_int, C.c_double]

    def _chk(self, rc):
        if rc < 0:
            msg = self._f("last_error")()
            raise ApiError((msg or b"?").decode())
        return rc

    def close(self):
        if self.h is not None:
            self._f("destroy")(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- vertices --
    def add_pose(self, init7=None):
        p = _dp(_f64(init7, (7,))) if init7 is not None else None
        return self._chk(self._f("add_pose")(self.h, p))

    def add_plane(self, abcd=None):
        p = _dp(_f64(abcd, (4,))) if abcd is not None else None
        return self._chk(self._f("add_plane")(self.h, p))

    def add_poses(self, init7s):
        v = _f64(init7s)
        ids = np.empty(len(v), dtype=np.int32)
        self._chk(self._f("add_poses")(self.h, len(v), _dp(v), _ip(ids)))
        return ids

    def add_planes(self, abcds):
        v = _f64(abcds)
        ids = np.empty(len(v), dtype=np.int32)
        self._chk(self._f("add_planes")(self.h, len(v), _dp(v), _ip(ids)))
        return ids

    def init_pose(self, id_, init7):
        self._chk(self._f("init_pose")(self.h, int(id_), _dp(_f64(init7, (7,)))))

    def init_plane(self, id_, abcd):
        self._chk(self._f("init_plane")(self.h, int(id_), _dp(_f64(abcd, (4,)))))

    def init_poses(self, ids, init7s):
        ids, v = _i32(ids), _f64(init7s, (len(ids), 7))
        self._chk(self._f("init_poses")(self.h, len(ids), _ip(ids), _dp(v)))

    def init_planes(self, ids, abcds):
        ids, v = _i32(ids), _f64(abcds, (len(ids), 4))
        self._chk(self._f("init_planes")(self.h, len(ids), _ip(ids), _dp(v)))

    def get_pose(self, id_):
        out = np.empty(7)
        self._chk(self._f("get_pose")(self.h, int(id_), _dp(out)))
        return out

    def get_plane(self, id_):
        out = np.empty(4)
        self._chk(self._f("get_plane")(self.h, int(id_), _dp(out)))
        return out

    def get_poses(self, ids):
        ids = _i32(ids)
        out = np.empty((len(ids), 7))
        self._chk(self._f("get_poses")(self.h, len(ids), _ip(ids), _dp(out)))
        return out

    def get_planes(self, ids):
        ids = _i32(ids)
        out = np.empty((len(ids), 4))
        self._chk(self._f("get_planes")(self.h, len(ids), _ip(ids), _dp(out)))
        return out

    # -- edges --
    def add_pose_prior(self, pose, xyzypr, sqrtinf21):
        return self._chk(self._f("add_pose_prior")(self.h, int(pose), _dp(_f64(xyzypr, (6,))), _dp(_f64(sqrtinf21, (21,)))))

    def add_odometry(self, p1, p2, xyzypr, sqrtinf21):
        return self._chk(self._f("add_odometry")(self.h, int(p1), int(p2), _dp(_f64(xyzypr, (6,))), _dp(_f64(sqrtinf21, (21,)))))

    def add_pose_plane(self, pose, plane, meas4, sqrtinf6):
        return self._chk(self._f("add_pose_plane")(self.h, int(pose), int(plane), _dp(_f64(meas4, (4,))), _dp(_f64(sqrtinf6, (6,)))))

    def add_pose_plane2(self, pose, plane, meas4, rays6, sqrtinf6):
        """Pose3d_Plane3d_Factor2: measurement re-popped from two ground-edge rays inside the residual."""
        fn = self._f("add_pose_plane2")
        fn.argtypes = [C.c_void_p, C.c_int, C.c_int, c_double_p, c_double_p, c_double_p]
        return self._chk(fn(self.h, int(pose), int(plane), _dp(_f64(meas4, (4,))), _dp(_f64(rays6, (6,))), _dp(_f64(sqrtinf6, (6,)))))

    def add_plane_prior(self, plane, abcd, sqrtinf6):
        return self._chk(self._f("add_plane_prior")(self.h, int(plane), _dp(_f64(abcd, (4,))), _dp(_f64(sqrtinf6, (6,)))))

    def add_odometry_bulk(self, p1, p2, xyzypr, sqrtinf21):
        p1, p2 = _i32(p1), _i32(p2)
        m, s = _f64(xyzypr, (len(p1), 6)), _f64(sqrtinf21, (len(p1), 21))
        out = np.empty(len(p1), dtype=np.int32)
        if len(p1):
            self._chk(self._f("add_odometry_bulk")(self.h, len(p1), _ip(p1), _ip(p2), _dp(m), _dp(s), _ip(out)))
        return out

    def add_pose_plane_bulk(self, pose, plane, meas4, sqrtinf6):
        pose, plane = _i32(pose), _i32(plane)
        m, s = _f64(meas4, (len(pose), 4)), _f64(sqrtinf6, (len(pose), 6))
        out = np.empty(len(pose), dtype=np.int32)
        if len(pose):
            self._chk(self._f("add_pose_plane_bulk")(self.h, len(pose), _ip(pose), _ip(plane), _dp(m), _dp(s), _ip(out)))
        return out

    def set_measurement(self, fid, meas):
        self._chk(self._f("set_measurement")(self.h, int(fid), _dp(_f64(meas))))

    def get_measurement(self, fid, n=4):
        out = np.zeros(6)
        self._chk(self._f("get_measurement")(self.h, int(fid), _dp(out)))
        return out[:n]

    def remove_factor(self, fid):
        self._chk(self._f("remove_factor")(self.h, int(fid)))

    def remove_node(self, id_):
        self._chk(self._f("remove_node")(self.h, int(id_)))

    def num_nodes(self):
        return self._chk(self._f("num_nodes")(self.h))

    def num_factors(self):
        return self._chk(self._f("num_factors")(self.h))

    def factor_nodes(self, fid):
        out = np.zeros(2, dtype=np.int32)
        n = self._chk(self._f("factor_nodes")(self.h, int(fid), _ip(out)))
        return out[:n].tolist()

    def node_factors(self, id_, cap=1 << 16):
        out = np.zeros(cap, dtype=np.int32)
        n = self._chk(self._f("node_factors")(self.h, int(id_), _ip(out), cap))
        return out[:min(n, cap)].tolist()

    def node_start(self, id_):
        return self._f("node_start")(self.h, int(id_))

    def factor_row(self, fid):
        return self._f("factor_row")(self.h, int(fid))

    # -- configuration --
    def get_properties(self):
        p = Properties()
        self._chk(self._f("get_properties")(self.h, C.byref(p)))
        return {n: getattr(p, n) for n, _ in Properties._fields_}

    def set_properties(self, **kw):
        p = Properties()
        self._chk(self._f("get_properties")(self.h, C.byref(p)))
        for k, v in kw.items():
            if not hasattr(p, k):
                raise KeyError(k)
            setattr(p, k, v)
        self._chk(self._f("set_properties")(self.h, C.byref(p)))

    def set_robust(self, kind, b):
        self._chk(self._f("set_robust")(self.h, int(kind), float(b)))

    # -- optimisation --
    def batch_optimize(self):
        it = C.c_int(0)
        self._chk(self._f("batch_optimize")(self.h, C.byref(it)))
        return it.value

    def update(self):
        self._chk(self._f("update")(self.h))

    def chi2(self):
        out = C.c_double(0)
        self._chk(self._f("chi2")(self.h, C.byref(out)))
        return out.value

    def trace(self, cap=4096):
        lam, en, eb, dn = (np.zeros(cap) for _ in range(4))
        acc, pcg = np.zeros(cap, dtype=np.int32), np.zeros(cap, dtype=np.int32)
        n = self._chk(self._f("get_trace")(self.h, cap, _dp(lam), _dp(en), _dp(eb), _dp(dn), _ip(acc), _ip(pcg)))
        n = min(n, cap)
        return dict(lam=lam[:n], chi2_new=en[:n], chi2_before=eb[:n], delta_norm=dn[:n], accepted=acc[:n], pcg=pcg[:n])

    def save_graph(self, path, precision=0):
        """Slam::save text format (include/popup_gpu.h); host only."""
        fn = self._f("save_graph")
        fn.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        self._chk(fn(self.h, str(path).encode(), int(precision)))

    def load_isam_dataset(self, path):
        """3-D part of the iSAM dataset grammar (EDGE3, POSE3D_INIT); returns (poses added, factors added)."""
        fn = self._f("load_isam_dataset")
        fn.argtypes = [C.c_void_p, C.c_char_p, c_int_p, c_int_p]
        a, b = C.c_int(0), C.c_int(0)
        self._chk(fn(self.h, str(path).encode(), C.byref(a), C.byref(b)))
        return a.value, b.value

    def refresh_plane_measurements(self, frame_pose, seg_ptr, segs, invK, map_fid, map_frame, map_row):
        """Mapper_mono::update_plane_measurement on the resident estimates (include/popup_gpu.h); returns the new
        measurements [n_map, 4]."""
        frame_pose, seg_ptr = _i32(frame_pose), _i32(seg_ptr)
        map_fid, map_frame, map_row = _i32(map_fid), _i32(map_frame), _i32(map_row)
        segs = np.ascontiguousarray(segs, dtype=np.float32).reshape(-1, 4)
        invK = np.ascontiguousarray(invK, dtype=np.float32).reshape(3, 3)
        out = np.zeros((len(map_fid), 4))
        fn = self._f("refresh_plane_measurements")
        fn.argtypes = [C.c_void_p, C.c_int, c_int_p, c_int_p, c_float_p, c_float_p, C.c_int, c_int_p, c_int_p, c_int_p, c_double_p]
        self._chk(fn(self.h, len(frame_pose), _ip(frame_pose), _ip(seg_ptr), segs.ctypes.data_as(c_float_p),
                     invK.ctypes.data_as(c_float_p), len(map_fid), _ip(map_fid), _ip(map_frame), _ip(map_row), _dp(out)))
        return out

    def refresh_bind(self, frame_pose, seg_ptr, segs, invK, map_fid, map_frame, map_row):
        """pus_refresh_bind: keep the frames' segment lists / invK / factor map resident on the device."""
        frame_pose, seg_ptr = _i32(frame_pose), _i32(seg_ptr)
        segs = np.ascontiguousarray(segs, dtype=np.float32).reshape(-1, 4)
        invK = np.ascontiguousarray(invK, dtype=np.float32).reshape(3, 3)
        map_fid, map_frame, map_row = _i32(map_fid), _i32(map_frame), _i32(map_row)
        self._n_map = len(map_fid)
        self._chk(self._f("refresh_bind")(self.h, len(frame_pose), _ip(frame_pose), _ip(seg_ptr), segs.ctypes.data_as(c_float_p),
                                          invK.ctypes.data_as(c_float_p), len(map_fid), _ip(map_fid), _ip(map_frame), _ip(map_row)))

    def refresh_run(self, want_output=False):
        """pus_refresh_run: re-pop every bound frame from the resident pose estimates; new measurements stay on the device unless
        `want_output`."""
        out = np.zeros((self._n_map, 4)) if want_output else None
        self._chk(self._f("refresh_run")(self.h, _dp(out) if want_output else None))
        return out

    def project_to_planes(self, plane_ids, pts):
        """Plane3d::project_to_plane of float points onto the current plane estimates (include/popup_gpu.h)."""
        plane_ids = _i32(plane_ids)
        pts = np.ascontiguousarray(pts, dtype=np.float32).reshape(-1, 3)
        out = np.zeros_like(pts)
        fn = self._f("project_to_planes")
        fn.argtypes = [C.c_void_p, C.c_int, c_int_p, c_float_p, c_float_p]
        self._chk(fn(self.h, len(plane_ids), _ip(plane_ids), pts.ctypes.data_as(c_float_p), out.ctypes.data_as(c_float_p)))
        return out


class GpuGraphAPI(GraphAPI):
    """GraphAPI plus the entry points only the CUDA library has (solver options, split
    upload/solve/download, stats, debug hooks)."""

    def __init__(self, lib=None, device=0):
        super().__init__(lib if lib is not None else load_library(), "pus_", device)

    def set_stream(self, cuda_stream_ptr):
        self._chk(self.lib.pus_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    def set_jacobian_mode(self, mode):
        """Same convention as the oracle binding: 0 = the reference's numerical differences (numericalDiff, eps = 1e-4), 1 = closed
        forms (the default of the CUDA library)."""
        self._chk(self.lib.pus_set_jacobian_mode(self.h, 1 if int(mode) == 0 else 0))

    def get_solver_options(self):
        o = SolverOptions()
        self._chk(self.lib.pus_get_solver_options(self.h, C.byref(o)))
        return o

    def set_solver_options(self, **kw):
        o = self.get_solver_options()
        for k, v in kw.items():
            if not hasattr(o, k):
                raise KeyError(k)
            setattr(o, k, v)
        self._chk(self.lib.pus_set_solver_options(self.h, C.byref(o)))

    def upload(self):
        self._chk(self.lib.pus_upload(self.h))

    def solve_resident(self):
        it = C.c_int(0)
        self._chk(self.lib.pus_solve_resident(self.h, C.byref(it)))
        return it.value

    def download(self):
        self._chk(self.lib.pus_download(self.h))

    def stats(self):
        s = Stats()
        self._chk(self.lib.pus_get_stats(self.h, C.byref(s)))
        return s.as_dict()

    def span_export(self):
        buf = C.create_string_buffer(64)
        fn = self.lib.pus_span_export
        fn.argtypes = [C.c_void_p, C.c_void_p]
        self._chk(fn(self.h, buf))
        return bytes(buf.raw)

    def span_connect(self, rank, world, handles):
        blob = b"".join(handles)
        assert len(blob) == 64 * world
        fn = self.lib.pus_span_connect
        fn.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_char_p]
        self._chk(fn(self.h, int(rank), int(world), blob))

    def span_optimize(self):
        it = C.c_int(0)
        fn = self.lib.pus_span_optimize
        fn.argtypes = [C.c_void_p, c_int_p]
        self._chk(fn(self.h, C.byref(it)))
        return it.value

    def span_disconnect(self):
        fn = self.lib.pus_span_disconnect
        fn.argtypes = [C.c_void_p]
        self._chk(fn(self.h))

    def debug_fetch(self, name, cap):
        out = np.zeros(int(cap))
        n = self.lib.pus_debug_fetch(self.h, name.encode(), _dp(out), C.c_longlong(int(cap)))
        if n < 0:
            raise ApiError((self.lib.pus_last_error() or b"?").decode())
        return out[:n]

    def debug_store(self, name, arr):
        a = _f64(arr).reshape(-1)
        n = self.lib.pus_debug_store(self.h, name.encode(), _dp(a), C.c_longlong(len(a)))
        if n < 0:
            raise ApiError((self.lib.pus_last_error() or b"?").decode())

    def debug_run_stage(self, stage, lam=0.0):
        self._chk(self.lib.pus_debug_run_stage(self.h, int(stage), float(lam)))


def _handles(apis):
    arr = (C.c_void_p * len(apis))(*[a.h for a in apis])
    return arr


def batch_optimize_many(apis):
    lib = apis[0].lib
    its = np.zeros(len(apis), dtype=np.int32)
    rc = lib.pus_batch_optimize_many(_handles(apis), len(apis), _ip(its))
    if rc < 0:
        raise ApiError((lib.pus_last_error() or b"?").decode())
    return its


def span_emulate_optimize(apis):
    """one graph spanning len(apis) "ranks" emulated on one device (include/popup_gpu.h); every api holds the same graph"""
    lib = apis[0].lib
    it = C.c_int(0)
    rc = lib.pus_span_emulate_optimize(_handles(apis), len(apis), C.byref(it))
    if rc < 0:
        raise ApiError((lib.pus_last_error() or b"?").decode())
    return it.value


def upload_many(apis):
    rc = apis[0].lib.pus_upload_many(_handles(apis), len(apis))
    if rc < 0:
        raise ApiError((apis[0].lib.pus_last_error() or b"?").decode())


def solve_resident_many(apis):
    lib = apis[0].lib
    its = np.zeros(len(apis), dtype=np.int32)
    rc = lib.pus_solve_resident_many(_handles(apis), len(apis), _ip(its))
    if rc < 0:
        raise ApiError((lib.pus_last_error() or b"?").decode())
    return its


def download_many(apis):
    rc = apis[0].lib.pus_download_many(_handles(apis), len(apis))
    if rc < 0:
        raise ApiError((apis[0].lib.pus_last_error() or b"?").decode())


def popup_fit_frames(lib, seg_ptr, segs, invK, Ts, dist_thre=10.0, mode=0, prefix="pus_", device=0):
    """popup_plane::get_plane_equation batched over frames (include/popup_gpu.h)."""
    seg_ptr = _i32(seg_ptr)
    nf = len(seg_ptr) - 1
    segs = np.ascontiguousarray(segs, dtype=np.float32).reshape(-1, 4)
    invK = np.ascontiguousarray(invK, dtype=np.float32).reshape(3, 3)
    Ts = np.ascontiguousarray(Ts, dtype=np.float32).reshape(nf, 4, 4)
    rows = int(seg_ptr[-1]) + nf
    pw = np.zeros((rows, 4), dtype=np.float32)
    ps = np.zeros((rows, 4), dtype=np.float32)
    dist = np.zeros(rows, dtype=np.float32)
    good = np.zeros(rows, dtype=np.int32)
    fn = getattr(lib, prefix + "popup_fit_frames")
    fn.argtypes = [C.c_int, C.c_int, c_int_p, c_float_p, c_float_p, c_float_p, C.c_float, C.c_int, c_float_p,
                   c_float_p, c_float_p, c_int_p]
    rc = fn(device, nf, _ip(seg_ptr), segs.ctypes.data_as(c_float_p), invK.ctypes.data_as(c_float_p),
            Ts.ctypes.data_as(c_float_p), float(dist_thre), int(mode), pw.ctypes.data_as(c_float_p),
            ps.ctypes.data_as(c_float_p), dist.ctypes.data_as(c_float_p), _ip(good))
    if rc < 0:
        raise ApiError("popup_fit_frames failed")
    return pw, ps, dist, good

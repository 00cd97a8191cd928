"""ctypes binding of the C ABI declared in include/visfs_ba.h.

Python is only the test / benchmark harness language here: the product is the CUDA library
(`visfs_b200/csrc/libvisfs_ba.so`) and the C++17 `VISFS::Optimizer::Optimizer` shim above it
(`visfs_b200/host/`).  There is no CPU fallback: if the shared library is missing, loading
raises, and if no CUDA device is present every compute call returns VISFS_BA_ERR_CUDA.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.environ.get("VISFS_BA_LIB") or os.path.join(ROOT, "visfs_b200", "csrc", "libvisfs_ba.so")   # (override: A/B runs of tools/)

OK, ERR_INVALID, ERR_CUDA, ERR_NUMERIC_PASS1, ERR_NUMERIC_PASS2, ERR_UNSUPPORTED = range(6)
EDGE_STEREO, EDGE_MONO = 0, 1
FLAG_PARTITIONED, FLAG_SINGLE_PASS, FLAG_NO_CULL = 1, 2, 4
STOP_NOT_RUN, STOP_ITERATIONS, STOP_TERMINATE, STOP_EMPTY, STOP_SOLVER_FAIL = range(5)
COMM_ID_BYTES = 128

_dp = C.POINTER(C.c_double)
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_u8p = C.POINTER(C.c_uint8)


class Config(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("device", C.c_int32), ("profile_kernels", C.c_int32),
                ("reserved", C.c_int32 * 5)]


class Problem(C.Structure):
    _fields_ = [("n_poses", C.c_int32), ("n_points", C.c_int32), ("n_edges", C.c_int32), ("flags", C.c_uint32),
                ("pose_tq", _dp), ("pose_id", _i64p), ("pose_fixed", _u8p),
                ("point_xyz", _dp), ("point_id", _i64p), ("point_fixed", _u8p),
                ("edge_obs", _dp), ("edge_pose", _i32p), ("edge_point", _i32p), ("edge_kind", _u8p),
                ("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double), ("bf", C.c_double),
                ("pixel_variance", C.c_double), ("huber_delta", C.c_double),
                ("iterations", C.c_int32), ("solver", C.c_int32), ("trust_region", C.c_int32), ("n_links", C.c_int32),
                ("link_from", _i32p), ("link_to", _i32p), ("link_tq", _dp), ("odometry_variance", C.c_double),
                ("edge_obs_f32", C.POINTER(C.c_float))]


class Result(C.Structure):
    _fields_ = [("pose_tq", _dp), ("point_xyz", _dp), ("edge_level", _u8p),
                ("status", C.c_int32), ("n_outliers", C.c_int32),
                ("iterations_run", C.c_int32 * 2), ("trials_run", C.c_int32 * 2), ("stop_reason", C.c_int32 * 2),
                ("n_free_poses", C.c_int32 * 2), ("n_free_points", C.c_int32 * 2),
                ("chi2_initial", C.c_double), ("chi2_pass1", C.c_double), ("chi2_final", C.c_double),
                ("chi2_last_trial", C.c_double), ("lambda_final", C.c_double * 2)]


class Linearization(C.Structure):
    _fields_ = [("error", _dp), ("chi2", _dp), ("rho", _dp), ("weight", _dp), ("J_point", _dp), ("J_pose", _dp)]


class Structure(C.Structure):
    _fields_ = [("edge_level", _u8p), ("pose_hidx", _i32p), ("point_hidx", _i32p), ("edge_active", _u8p),
                ("hpl_row", _i32p), ("hpl_col", _i32p), ("schur_rows", _i32p), ("schur_cols", _i32p),
                ("schur_capacity", C.c_int32), ("n_schur_blocks", C.c_int32), ("n_free_poses", C.c_int32),
                ("n_free_points", C.c_int32), ("n_active_edges", C.c_int32), ("n_hpl_blocks", C.c_int32)]


class WindowConfig(C.Structure):
    _fields_ = [("max_frames", C.c_int32), ("max_points", C.c_int32), ("max_observations", C.c_int32), ("reserved0", C.c_int32),
                ("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double), ("bf", C.c_double),
                ("pixel_variance", C.c_double), ("huber_delta", C.c_double),
                ("iterations", C.c_int32), ("solver", C.c_int32), ("trust_region", C.c_int32), ("reserved1", C.c_int32)]


class WindowResult(C.Structure):
    _fields_ = [("frame_id", _i64p), ("pose_tq", _dp), ("outlier_point_id", _i64p), ("outlier_frame_id", _i64p),
                ("outlier_capacity", C.c_int32),
                ("n_frames", C.c_int32), ("n_points", C.c_int32), ("n_edges", C.c_int32), ("n_outliers", C.c_int32),
                ("status", C.c_int32), ("iterations_run", C.c_int32 * 2), ("trials_run", C.c_int32 * 2),
                ("stop_reason", C.c_int32 * 2),
                ("chi2_initial", C.c_double), ("chi2_pass1", C.c_double), ("chi2_final", C.c_double),
                ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64)]


class Timing(C.Structure):
    _fields_ = [("total_ms", C.c_double), ("build_ms", C.c_double), ("solve_ms", C.c_double),
                ("update_ms", C.c_double), ("other_ms", C.c_double),
                ("build_launches", C.c_int64), ("solve_launches", C.c_int64), ("update_launches", C.c_int64),
                ("other_launches", C.c_int64), ("lm_iterations", C.c_int64), ("lm_trials", C.c_int64),
                ("edge_trials", C.c_int64), ("alg_bytes_build", C.c_int64), ("alg_bytes_update", C.c_int64),
                ("kernel_launches", C.c_int64), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
                ("solve_clocks", C.c_int64 * 6)]


EXPORTS = ["visfs_ba_abi_version", "visfs_ba_create", "visfs_ba_destroy", "visfs_ba_last_error", "visfs_ba_solve",
           "visfs_ba_solve_batch", "visfs_ba_linearize", "visfs_ba_structure_build", "visfs_ba_upload",
           "visfs_ba_run_resident", "visfs_ba_download", "visfs_ba_get_timing", "visfs_ba_comm_unique_id",
           "visfs_ba_comm_init", "visfs_ba_comm_destroy", "visfs_ba_probe_fp64", "visfs_ba_debug_trial",
           "visfs_ba_host_alloc", "visfs_ba_host_free", "visfs_ba_debug_pose_oplus", "visfs_ba_debug_link_linearize",
           "visfs_ba_window_create", "visfs_ba_window_destroy", "visfs_ba_window_set_points", "visfs_ba_window_insert_frame",
           "visfs_ba_window_remove_frame", "visfs_ba_window_remove_points", "visfs_ba_window_remove_observations",
           "visfs_ba_window_set_poses", "visfs_ba_window_set_links", "visfs_ba_window_solve", "visfs_ba_window_get_points", "visfs_ba_window_h2d_bytes_total"]


def _ptr(a, typ):
    return None if a is None else a.ctypes.data_as(typ)


def with_float_observations(w):
    """The same window with its observations as floats (edge_obs_f32).  Exact for windows whose observations come from float
    key points and depths, which is what the reference has (synth.marshal_observations); refuses to round otherwise."""
    f = np.ascontiguousarray(w["edge_obs"], dtype=np.float32)
    if not np.array_equal(f.astype(np.float64), np.asarray(w["edge_obs"], dtype=np.float64)):
        raise ValueError("edge_obs is not representable in float32")
    out = dict(w)
    out["edge_obs_f32"] = f
    return out


def window_to_problem(w, keep):
    """Fill a Problem from a window dict (visfs_b200.synth.make_window).  `keep` is a list that
    receives the contiguous arrays so they outlive the call."""
    def arr(key, dtype):
        a = w.get(key)
        if a is None:
            return None
        a = np.ascontiguousarray(a, dtype=dtype)
        keep.append(a)
        return a
    p = Problem()
    p.n_poses, p.n_points, p.n_edges = int(w["n_poses"]), int(w["n_points"]), int(w["n_edges"])
    p.flags = int(w.get("flags", 0))
    p.pose_tq = _ptr(arr("pose_tq", np.float64), _dp)
    p.pose_id = _ptr(arr("pose_id", np.int64), _i64p)
    p.pose_fixed = _ptr(arr("pose_fixed", np.uint8), _u8p)
    p.point_xyz = _ptr(arr("point_xyz", np.float64), _dp)
    p.point_id = _ptr(arr("point_id", np.int64), _i64p)
    p.point_fixed = _ptr(arr("point_fixed", np.uint8), _u8p)
    if w.get("edge_obs_f32") is not None:    # observations as floats (include/visfs_ba.h): used instead of edge_obs
        p.edge_obs_f32 = _ptr(arr("edge_obs_f32", np.float32), C.POINTER(C.c_float))
    else:
        p.edge_obs = _ptr(arr("edge_obs", np.float64), _dp)
    p.edge_pose = _ptr(arr("edge_pose", np.int32), _i32p)
    p.edge_point = _ptr(arr("edge_point", np.int32), _i32p)
    p.edge_kind = _ptr(arr("edge_kind", np.uint8), _u8p)
    for k in ("fx", "fy", "cx", "cy", "bf", "pixel_variance", "huber_delta"):
        setattr(p, k, float(w[k]))
    p.iterations, p.solver, p.trust_region = int(w["iterations"]), int(w["solver"]), int(w["trust_region"])
    p.n_links = int(w.get("n_links", 0))
    if p.n_links:
        p.link_from = _ptr(arr("link_from", np.int32), _i32p)
        p.link_to = _ptr(arr("link_to", np.int32), _i32p)
        p.link_tq = _ptr(arr("link_tq", np.float64), _dp)
    p.odometry_variance = float(w.get("odometry_variance", 0.00005))
    return p


class ResultArrays:
    """Owns the output arrays a Result points at."""

    def __init__(self, w):
        self.pose_tq = np.zeros((int(w["n_poses"]), 7))
        self.point_xyz = np.zeros((int(w["n_points"]), 3))
        self.edge_level = np.zeros(int(w["n_edges"]), dtype=np.uint8)

    def bind(self, r: Result):
        r.pose_tq = _ptr(self.pose_tq, _dp)
        r.point_xyz = _ptr(self.point_xyz, _dp)
        r.edge_level = _ptr(self.edge_level, _u8p)


def result_to_dict(r: Result, arrays: ResultArrays):
    d = dict(pose_tq=arrays.pose_tq, point_xyz=arrays.point_xyz, edge_level=arrays.edge_level)
    for name, typ in Result._fields_:
        if name in ("pose_tq", "point_xyz", "edge_level"):
            continue
        v = getattr(r, name)
        d[name] = list(v) if hasattr(v, "__len__") else v
    return d


def new_linearization(E):
    bufs = dict(error=np.zeros((E, 3)), chi2=np.zeros(E), rho=np.zeros(E), weight=np.zeros(E),
                J_point=np.zeros((E, 3, 3)), J_pose=np.zeros((E, 3, 6)))
    lin = Linearization()
    for k, a in bufs.items():
        setattr(lin, k, _ptr(a, _dp))
    return lin, bufs


def new_structure(w, edge_level=None, capacity=None):
    P, L, E = int(w["n_poses"]), int(w["n_points"]), int(w["n_edges"])
    capacity = capacity if capacity is not None else P * (P + 1) // 2
    bufs = dict(pose_hidx=np.zeros(P, np.int32), point_hidx=np.zeros(L, np.int32), edge_active=np.zeros(E, np.uint8),
                hpl_row=np.zeros(E, np.int32), hpl_col=np.zeros(E, np.int32),
                schur_rows=np.zeros(capacity, np.int32), schur_cols=np.zeros(capacity, np.int32))
    s = Structure()
    if edge_level is not None:
        bufs["edge_level"] = np.ascontiguousarray(edge_level, dtype=np.uint8)
        s.edge_level = _ptr(bufs["edge_level"], _u8p)
    s.pose_hidx, s.point_hidx = _ptr(bufs["pose_hidx"], _i32p), _ptr(bufs["point_hidx"], _i32p)
    s.edge_active = _ptr(bufs["edge_active"], _u8p)
    s.hpl_row, s.hpl_col = _ptr(bufs["hpl_row"], _i32p), _ptr(bufs["hpl_col"], _i32p)
    s.schur_rows, s.schur_cols = _ptr(bufs["schur_rows"], _i32p), _ptr(bufs["schur_cols"], _i32p)
    s.schur_capacity = capacity
    return s, bufs


def structure_to_dict(s: Structure, bufs):
    d = {k: v for k, v in bufs.items() if k != "edge_level"}
    n = min(s.n_schur_blocks, s.schur_capacity)
    d["schur_rows"], d["schur_cols"] = bufs["schur_rows"][:n].copy(), bufs["schur_cols"][:n].copy()
    for k in ("n_schur_blocks", "n_free_poses", "n_free_points", "n_active_edges", "n_hpl_blocks"):
        d[k] = getattr(s, k)
    return d


class BAError(RuntimeError):
    pass


def load_library(path=LIB_PATH):
    if not os.path.exists(path):
        raise BAError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                      "(there is no CPU fallback for the bundle adjustment)")
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    lib.visfs_ba_create.argtypes = [C.POINTER(Config), C.POINTER(C.c_void_p)]
    lib.visfs_ba_destroy.argtypes = [C.c_void_p]
    lib.visfs_ba_destroy.restype = None
    lib.visfs_ba_last_error.argtypes = [C.c_void_p]
    lib.visfs_ba_last_error.restype = C.c_char_p
    lib.visfs_ba_solve.argtypes = [C.c_void_p, C.POINTER(Problem), C.POINTER(Result)]
    lib.visfs_ba_solve_batch.argtypes = [C.c_void_p, C.c_int32, C.POINTER(Problem), C.POINTER(Result)]
    lib.visfs_ba_linearize.argtypes = [C.c_void_p, C.POINTER(Problem), C.POINTER(Linearization)]
    lib.visfs_ba_structure_build.argtypes = [C.c_void_p, C.POINTER(Problem), C.POINTER(Structure)]
    lib.visfs_ba_upload.argtypes = [C.c_void_p, C.c_int32, C.POINTER(Problem)]
    lib.visfs_ba_run_resident.argtypes = [C.c_void_p]
    lib.visfs_ba_download.argtypes = [C.c_void_p, C.c_int32, C.POINTER(Result)]
    lib.visfs_ba_get_timing.argtypes = [C.c_void_p, C.POINTER(Timing)]
    lib.visfs_ba_comm_unique_id.argtypes = [C.c_void_p]
    lib.visfs_ba_comm_init.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
    lib.visfs_ba_comm_destroy.argtypes = [C.c_void_p]
    lib.visfs_ba_probe_fp64.argtypes = [C.c_void_p, _dp]
    lib.visfs_ba_debug_trial.argtypes = [C.c_void_p, C.POINTER(Problem), C.c_double, _dp, _dp, _dp, _dp,
                                         C.POINTER(C.c_int32), _dp, _dp, _dp]
    lib.visfs_ba_debug_pose_oplus.argtypes = [C.c_void_p, C.c_int32, _dp, _dp, _dp]
    lib.visfs_ba_debug_link_linearize.argtypes = [C.c_void_p, C.c_int32, _dp, _dp, _dp, _dp, _dp, _dp]
    lib.visfs_ba_window_create.argtypes = [C.c_void_p, C.POINTER(WindowConfig), C.POINTER(C.c_void_p)]
    lib.visfs_ba_window_destroy.argtypes = [C.c_void_p]
    lib.visfs_ba_window_destroy.restype = None
    lib.visfs_ba_window_set_points.argtypes = [C.c_void_p, C.c_int32, _i64p, _dp, _u8p]
    lib.visfs_ba_window_insert_frame.argtypes = [C.c_void_p, C.c_int64, _dp, C.c_int32, _i64p, C.POINTER(C.c_float), _u8p]
    lib.visfs_ba_window_remove_frame.argtypes = [C.c_void_p, C.c_int64]
    lib.visfs_ba_window_remove_points.argtypes = [C.c_void_p, C.c_int32, _i64p]
    lib.visfs_ba_window_remove_observations.argtypes = [C.c_void_p, C.c_int32, _i64p, _i64p]
    lib.visfs_ba_window_set_poses.argtypes = [C.c_void_p, C.c_int32, _i64p, _dp]
    lib.visfs_ba_window_set_links.argtypes = [C.c_void_p, C.c_int32, _i64p, _i64p, _dp, C.c_double]
    lib.visfs_ba_window_solve.argtypes = [C.c_void_p, C.c_int64, C.POINTER(WindowResult)]
    lib.visfs_ba_window_get_points.argtypes = [C.c_void_p, C.c_int32, _i64p, _dp]
    lib.visfs_ba_window_h2d_bytes_total.argtypes = [C.c_void_p]
    lib.visfs_ba_window_h2d_bytes_total.restype = C.c_int64
    lib.visfs_ba_host_alloc.argtypes = [C.c_size_t]
    lib.visfs_ba_host_alloc.restype = C.c_void_p
    lib.visfs_ba_host_free.argtypes = [C.c_void_p]
    lib.visfs_ba_host_free.restype = None
    return lib


class PinnedArena:
    """Page-locked host memory from visfs_ba_host_alloc, handed out as numpy arrays (64-byte aligned).  Arrays placed
    here are moved by DMA straight between the caller's memory and the device (include/visfs_ba.h)."""

    def __init__(self, lib, nbytes):
        self.lib, self.nbytes, self.used = lib, int(nbytes), 0
        self.ptr = lib.visfs_ba_host_alloc(self.nbytes)
        if not self.ptr:
            raise BAError(f"visfs_ba_host_alloc({nbytes}) failed")
        self.buf = np.ctypeslib.as_array((C.c_uint8 * self.nbytes).from_address(self.ptr))

    @staticmethod
    def size_of(a):
        return (int(a.nbytes) + 63) & ~63

    def place(self, a):
        """Copy `a` into the arena; returns the page-locked array."""
        n = int(a.nbytes)
        if self.used + n > self.nbytes:
            raise BAError("pinned arena exhausted")
        out = self.buf[self.used:self.used + n].view(a.dtype).reshape(a.shape)
        out[...] = a
        self.used += (n + 63) & ~63
        return out

    def close(self):
        if getattr(self, "ptr", None):
            self.buf = None
            self.lib.visfs_ba_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class BundleAdjuster:
    """Thin object wrapper over one visfs_ba_handle (one CUDA device, one stream)."""

    def __init__(self, device=0, profile_kernels=False, lib=None):
        self.lib = lib or load_library()
        cfg = Config(abi_version=self.lib.visfs_ba_abi_version(), device=int(device),
                     profile_kernels=int(bool(profile_kernels)))
        h = C.c_void_p()
        st = self.lib.visfs_ba_create(C.byref(cfg), C.byref(h))
        if st != OK or not h:
            msg = self.lib.visfs_ba_last_error(None)
            raise BAError(f"visfs_ba_create failed ({st}): {msg.decode() if msg else ''}")
        self.h = h
        self.device = int(device)
        self._resident = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.visfs_ba_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def error(self):
        m = self.lib.visfs_ba_last_error(self.h)
        return m.decode() if m else ""

    def _check(self, st, allow=(OK,)):
        if st not in allow:
            raise BAError(f"status {st}: {self.error()}")
        return st

    def solve(self, w):
        keep = []
        p = window_to_problem(w, keep)
        out = ResultArrays(w)
        r = Result()
        out.bind(r)
        self._check(self.lib.visfs_ba_solve(self.h, C.byref(p), C.byref(r)),
                    allow=(OK, ERR_NUMERIC_PASS1, ERR_NUMERIC_PASS2))
        return result_to_dict(r, out)

    def _pack(self, windows, pinned=False):
        keep = []
        n = len(windows)
        outs = [ResultArrays(w) for w in windows]
        if pinned:
            # every input and output array moves into one page-locked arena (freed with the packed batch)
            keys = ("pose_tq", "pose_fixed", "point_xyz", "point_fixed", "edge_obs", "edge_obs_f32", "edge_pose", "edge_point", "edge_kind")
            cast = dict(pose_tq=np.float64, pose_fixed=np.uint8, point_xyz=np.float64, point_fixed=np.uint8,
                        edge_obs=np.float64, edge_obs_f32=np.float32, edge_pose=np.int32, edge_point=np.int32, edge_kind=np.uint8)
            arrs = [{k: np.ascontiguousarray(w[k], dtype=cast[k]) for k in keys
                     if w.get(k) is not None and not (k == "edge_obs" and w.get("edge_obs_f32") is not None)} for w in windows]
            total = sum(PinnedArena.size_of(a) for d in arrs for a in d.values())
            total += sum(PinnedArena.size_of(a) for o in outs for a in (o.pose_tq, o.point_xyz, o.edge_level))
            arena = PinnedArena(self.lib, total + 64)
            windows = [dict(w, **{k: arena.place(a) for k, a in d.items()}) for w, d in zip(windows, arrs)]
            for o in outs:
                o.pose_tq, o.point_xyz, o.edge_level = arena.place(o.pose_tq), arena.place(o.point_xyz), arena.place(o.edge_level)
            keep.append(arena)
        probs = (Problem * n)(*[window_to_problem(w, keep) for w in windows])
        res = (Result * n)()
        for r, o in zip(res, outs):
            o.bind(r)
        return n, probs, res, outs, keep

    def solve_batch(self, windows):
        n, probs, res, outs, keep = self._pack(windows)
        self._check(self.lib.visfs_ba_solve_batch(self.h, n, probs, res))
        return [result_to_dict(r, o) for r, o in zip(res, outs)]

    def prepare_batch(self, windows, pinned=False, float_obs=False):
        """Marshal once; returns an opaque packed batch for repeated solve_packed() calls.  pinned=True places every
        array in page-locked memory (visfs_ba_host_alloc): no staging copy on either side of the call."""
        if float_obs:
            windows = [with_float_observations(w) for w in windows]
        return self._pack(windows, pinned)

    def packed_results(self, packed):
        n, probs, res, outs, keep = packed
        return [result_to_dict(r, o) for r, o in zip(res, outs)]

    def solve_packed(self, packed):
        n, probs, res, outs, keep = packed
        self._check(self.lib.visfs_ba_solve_batch(self.h, n, probs, res))
        return res

    def upload(self, windows):
        self._resident = self._pack(windows)
        n, probs, res, outs, keep = self._resident
        self._check(self.lib.visfs_ba_upload(self.h, n, probs))

    def run_resident(self):
        self._check(self.lib.visfs_ba_run_resident(self.h))

    def download(self):
        n, probs, res, outs, keep = self._resident
        self._check(self.lib.visfs_ba_download(self.h, n, res))
        return [result_to_dict(r, o) for r, o in zip(res, outs)]

    def timing(self):
        t = Timing()
        self._check(self.lib.visfs_ba_get_timing(self.h, C.byref(t)))
        return {name: (list(getattr(t, name)) if name == "solve_clocks" else getattr(t, name)) for name, _ in Timing._fields_}

    def linearize(self, w):
        keep = []
        p = window_to_problem(w, keep)
        lin, bufs = new_linearization(int(w["n_edges"]))
        self._check(self.lib.visfs_ba_linearize(self.h, C.byref(p), C.byref(lin)))
        return bufs

    def structure(self, w, edge_level=None, capacity=None):
        keep = []
        p = window_to_problem(w, keep)
        s, bufs = new_structure(w, edge_level, capacity)
        self._check(self.lib.visfs_ba_structure_build(self.h, C.byref(p), C.byref(s)))
        return structure_to_dict(s, bufs)

    def debug_trial(self, w, lam=-1.0):
        keep = []
        p = window_to_problem(w, keep)
        nmax = 6 * int(w["n_poses"])
        S = np.zeros(nmax * nmax)
        bs = np.zeros(nmax)
        xp = np.zeros(nmax)
        tp = np.zeros((int(w["n_points"]), 3))
        n = C.c_int32()
        chi2, lam_out, tchi = C.c_double(), C.c_double(), C.c_double()
        self._check(self.lib.visfs_ba_debug_trial(self.h, C.byref(p), float(lam), _ptr(S, _dp), _ptr(bs, _dp), _ptr(xp, _dp),
                                                  _ptr(tp, _dp), C.byref(n), C.byref(chi2), C.byref(lam_out), C.byref(tchi)))
        nn = n.value
        return dict(n=nn, S=S[: nn * nn].reshape(nn, nn).copy(), b_s=bs[:nn].copy(), x_pose=xp[:nn].copy(), trial_points=tp,
                    chi2=chi2.value, lambda_used=lam_out.value, trial_chi2=tchi.value)

    def debug_pose_oplus(self, tq, delta):
        tq = np.ascontiguousarray(tq, dtype=np.float64)
        delta = np.ascontiguousarray(delta, dtype=np.float64)
        out = np.zeros_like(tq)
        self._check(self.lib.visfs_ba_debug_pose_oplus(self.h, len(tq), _ptr(tq, _dp), _ptr(delta, _dp), _ptr(out, _dp)))
        return out

    def debug_link_linearize(self, from_tq, to_tq, meas_tq):
        a, b, m = (np.ascontiguousarray(x, dtype=np.float64) for x in (from_tq, to_tq, meas_tq))
        n = len(a)
        err, Ji, Jj = np.zeros((n, 6)), np.zeros((n, 6, 6)), np.zeros((n, 6, 6))
        self._check(self.lib.visfs_ba_debug_link_linearize(self.h, n, _ptr(a, _dp), _ptr(b, _dp), _ptr(m, _dp), _ptr(err, _dp),
                                                           _ptr(Ji, _dp), _ptr(Jj, _dp)))
        return dict(error=err, J_from=Ji, J_to=Jj)

    def probe_fp64(self):
        v = C.c_double()
        self._check(self.lib.visfs_ba_probe_fp64(self.h, C.byref(v)))
        return v.value

    def comm_init(self, n_ranks, rank, unique_id: bytes):
        buf = C.create_string_buffer(unique_id, COMM_ID_BYTES)
        self._check(self.lib.visfs_ba_comm_init(self.h, n_ranks, rank, buf))

    def comm_init_torch(self, dist):
        """Bootstrap the library's NCCL communicator over an initialised torch.distributed group (plumbing only)."""
        world, rank = dist.get_world_size(), dist.get_rank()
        box = [self.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        self.comm_init(world, rank, box[0])

    def comm_destroy(self):
        """Tear the communicator down; every rank has to call this at the same point (NCCL finalises collectively)."""
        self._check(self.lib.visfs_ba_comm_destroy(self.h))

    def comm_unique_id(self) -> bytes:
        buf = C.create_string_buffer(COMM_ID_BYTES)
        self._check(self.lib.visfs_ba_comm_unique_id(buf))
        return buf.raw


class ResidentWindow:
    """visfs_ba_window_*: the local map kept in HBM and fed with LocalMap's deltas (include/visfs_ba.h)."""

    def __init__(self, ba: BundleAdjuster, max_frames, max_points, max_observations, *, fx, fy, cx, cy, bf, pixel_variance=1.5,
                 huber_delta=8.0, iterations=10, solver=0, trust_region=0):
        self.ba, self.lib = ba, ba.lib
        self.cfg = WindowConfig(max_frames=max_frames, max_points=max_points, max_observations=max_observations, fx=fx, fy=fy, cx=cx,
                                cy=cy, bf=bf, pixel_variance=pixel_variance, huber_delta=huber_delta, iterations=iterations,
                                solver=solver, trust_region=trust_region)
        self.w = C.c_void_p()
        ba._check(self.lib.visfs_ba_window_create(ba.h, C.byref(self.cfg), C.byref(self.w)))

    def close(self):
        if self.w:
            self.lib.visfs_ba_window_destroy(self.w)
            self.w = C.c_void_p()

    def set_points(self, ids, xyz, fixed=None):
        ids = np.ascontiguousarray(ids, dtype=np.int64)
        xyz = np.ascontiguousarray(xyz, dtype=np.float64)
        fixed = None if fixed is None else np.ascontiguousarray(fixed, dtype=np.uint8)
        self.ba._check(self.lib.visfs_ba_window_set_points(self.w, len(ids), _ptr(ids, _i64p), _ptr(xyz, _dp), _ptr(fixed, _u8p)))

    def insert_frame(self, frame_id, pose_tq, point_ids, obs, kind=None):
        tq = np.ascontiguousarray(pose_tq, dtype=np.float64)
        ids = np.ascontiguousarray(point_ids, dtype=np.int64)
        ob = np.ascontiguousarray(obs, dtype=np.float32)
        kd = None if kind is None else np.ascontiguousarray(kind, dtype=np.uint8)
        self.ba._check(self.lib.visfs_ba_window_insert_frame(self.w, int(frame_id), _ptr(tq, _dp), len(ids), _ptr(ids, _i64p),
                                                             _ptr(ob, C.POINTER(C.c_float)), _ptr(kd, _u8p)))

    def remove_frame(self, frame_id):
        self.ba._check(self.lib.visfs_ba_window_remove_frame(self.w, int(frame_id)))

    def remove_points(self, ids):
        ids = np.ascontiguousarray(ids, dtype=np.int64)
        self.ba._check(self.lib.visfs_ba_window_remove_points(self.w, len(ids), _ptr(ids, _i64p)))

    def remove_observations(self, point_ids, frame_ids):
        a = np.ascontiguousarray(point_ids, dtype=np.int64)
        b = np.ascontiguousarray(frame_ids, dtype=np.int64)
        self.ba._check(self.lib.visfs_ba_window_remove_observations(self.w, len(a), _ptr(a, _i64p), _ptr(b, _i64p)))

    def set_poses(self, frame_ids, pose_tq):
        a = np.ascontiguousarray(frame_ids, dtype=np.int64)
        tq = np.ascontiguousarray(pose_tq, dtype=np.float64)
        self.ba._check(self.lib.visfs_ba_window_set_poses(self.w, len(a), _ptr(a, _i64p), _ptr(tq, _dp)))

    def set_links(self, from_ids, to_ids, link_tq, odometry_variance):
        """Replace the odometry links of the map (by frame id); links whose frames are not both present are skipped by a solve."""
        a = np.ascontiguousarray(from_ids, dtype=np.int64)
        b = np.ascontiguousarray(to_ids, dtype=np.int64)
        tq = np.ascontiguousarray(link_tq, dtype=np.float64).reshape(-1, 7)
        assert len(a) == len(b) == len(tq)
        self.ba._check(self.lib.visfs_ba_window_set_links(self.w, len(a), _ptr(a, _i64p), _ptr(b, _i64p), _ptr(tq, _dp), float(odometry_variance)))

    def get_points(self, ids):
        ids = np.ascontiguousarray(ids, dtype=np.int64)
        out = np.zeros((len(ids), 3))
        self.ba._check(self.lib.visfs_ba_window_get_points(self.w, len(ids), _ptr(ids, _i64p), _ptr(out, _dp)))
        return out

    def solve(self, root_frame_id):
        F, cap = self.cfg.max_frames, self.cfg.max_observations
        fid, tq = np.zeros(F, dtype=np.int64), np.zeros((F, 7))
        op, of = np.zeros(cap, dtype=np.int64), np.zeros(cap, dtype=np.int64)
        r = WindowResult(frame_id=_ptr(fid, _i64p), pose_tq=_ptr(tq, _dp), outlier_point_id=_ptr(op, _i64p),
                         outlier_frame_id=_ptr(of, _i64p), outlier_capacity=cap)
        self.ba._check(self.lib.visfs_ba_window_solve(self.w, int(root_frame_id), C.byref(r)),
                       allow=(OK, ERR_NUMERIC_PASS1, ERR_NUMERIC_PASS2))
        n, k = r.n_frames, min(r.n_outliers, cap)
        return dict(frame_id=fid[:n].copy(), pose_tq=tq[:n].copy(), outliers=list(zip(op[:k].tolist(), of[:k].tolist())),
                    n_frames=n, n_points=r.n_points, n_edges=r.n_edges, n_outliers=r.n_outliers, status=r.status,
                    iterations_run=list(r.iterations_run), trials_run=list(r.trials_run), stop_reason=list(r.stop_reason),
                    chi2_initial=r.chi2_initial, chi2_pass1=r.chi2_pass1, chi2_final=r.chi2_final,
                    h2d_bytes=r.h2d_bytes, d2h_bytes=r.d2h_bytes)

    def h2d_bytes_total(self):
        return int(self.lib.visfs_ba_window_h2d_bytes_total(self.w))

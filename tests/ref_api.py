"""ctypes binding of oracle/_ref/libvisfs_ref.so: the reference's OWN CameraPose / VertexPose / EdgeStereo /
EdgePoseConstraint, compiled from its unmodified sources (oracle/Makefile, oracle/ref_shim.cpp).  Test infrastructure:
the library exists wherever `make -C oracle` ran with /root/reference present (this container) and travels to the GPU
box as a prebuilt file; where it is absent `available()` is False and the tests fall back to the committed fixtures
tests/golden/ref_*.npz that tests/golden/make_ref_golden.py generated from it."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
PATH = os.path.join(ORACLE_DIR, "_ref", "libvisfs_ref.so")
REFERENCE = "/root/reference/corelib/src/Optimizer/g2o/OptimizeTypeDefine.cpp"
_dp = C.POINTER(C.c_double)
_lib = None


def available():
    if not os.path.exists(PATH) and os.path.exists(REFERENCE):
        subprocess.run(["make", "-C", ORACLE_DIR, "-s", "ref"], check=False)
    return os.path.exists(PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libvisfs_ref.so is not built (needs /root/reference)")
        l = C.CDLL(PATH)
        l.visfs_ref_pose_from_matrix.argtypes = [_dp, _dp, _dp]
        l.visfs_ref_pose_normalize.argtypes = [_dp, _dp]
        l.visfs_ref_pose_map.argtypes = [_dp, _dp, _dp, _dp]
        l.visfs_ref_pose_oplus.argtypes = [_dp, _dp, _dp]
        l.visfs_ref_edge_stereo.argtypes = [C.c_int, _dp, _dp, _dp, _dp, _dp, _dp, _dp, C.POINTER(C.c_int)]
        l.visfs_ref_edge_pose_constraint.argtypes = [C.c_int, _dp, _dp, _dp, _dp, _dp, _dp]
        l.visfs_ref_point_oplus.argtypes = [_dp, _dp, _dp]
        l.visfs_ref_unorm3.argtypes = [C.c_double] * 3
        l.visfs_ref_unorm3.restype = C.c_double
        _lib = l
    return _lib


def _p(a):
    return a.ctypes.data_as(_dp)


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def edge_stereo(pose_tq, point, obs, intr):
    """EdgeStereo::computeError + linearizeOplus for n (pose, point, observation) triples."""
    pose_tq, point, obs, intr = _c(pose_tq), _c(point), _c(obs), _c(intr)
    n = len(pose_tq)
    err, Jl, Jp = np.zeros((n, 3)), np.zeros((n, 3, 3)), np.zeros((n, 3, 6))
    dp = np.zeros(n, dtype=np.int32)
    lib().visfs_ref_edge_stereo(n, _p(pose_tq), _p(point), _p(obs), _p(intr), _p(err), _p(Jl), _p(Jp),
                                dp.ctypes.data_as(C.POINTER(C.c_int)))
    return dict(error=err, J_point=Jl, J_pose=Jp, depth_positive=dp)


def edge_stereo_window(w):
    """The same for every edge of a window dict at its input state (mono edges: the caller masks row 2)."""
    intr = np.array([w["fx"], w["fy"], w["cx"], w["cy"], w["bf"]], dtype=np.float64)
    return edge_stereo(w["pose_tq"][w["edge_pose"]], w["point_xyz"][w["edge_point"]], w["edge_obs"], intr)


def edge_pose_constraint(from_tq, to_tq, meas_tq):
    from_tq, to_tq, meas_tq = _c(from_tq), _c(to_tq), _c(meas_tq)
    n = len(from_tq)
    err, Ji, Jj = np.zeros((n, 6)), np.zeros((n, 6, 6)), np.zeros((n, 6, 6))
    lib().visfs_ref_edge_pose_constraint(n, _p(from_tq), _p(to_tq), _p(meas_tq), _p(err), _p(Ji), _p(Jj))
    return dict(error=err, J_from=Ji, J_to=Jj)


def pose_oplus(tq, delta):
    tq, delta = _c(tq), _c(delta)
    out = np.zeros_like(tq)
    for i in range(len(tq)):
        lib().visfs_ref_pose_oplus(_p(tq[i]), _p(delta[i]), _p(out[i]))
    return out


def pose_from_matrix(R, t):
    R, t = _c(R), _c(t)
    out = np.zeros(7)
    lib().visfs_ref_pose_from_matrix(_p(R), _p(t), _p(out))
    return out


def pose_normalize(tq):
    tq = _c(tq)
    out = np.zeros(7)
    lib().visfs_ref_pose_normalize(_p(tq), _p(out))
    return out


def pose_map(tq, pw):
    tq, pw = _c(tq), _c(pw)
    pc, T = np.zeros(3), np.zeros((4, 4))
    lib().visfs_ref_pose_map(_p(tq), _p(pw), _p(pc), _p(T))
    return pc, T


def point_oplus(p, d):
    p, d = _c(p), _c(d)
    out = np.zeros(3)
    lib().visfs_ref_point_oplus(_p(p), _p(d), _p(out))
    return out


def unorm3(x, y, z):
    return float(lib().visfs_ref_unorm3(float(x), float(y), float(z)))

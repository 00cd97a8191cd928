"""ctypes binding of the CPU oracle (oracle/ba_oracle.cpp).  Test infrastructure: imported only by
tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from visfs_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_libs = {}


def build():
    subprocess.run(["make", "-C", ORACLE_DIR, "-s"], check=True)


def lib(omp=False):
    name = "liboracle_omp.so" if omp else "liboracle.so"
    if name not in _libs:
        path = os.path.join(ORACLE_DIR, "_build", name)
        if not os.path.exists(path):
            build()
        l = C.CDLL(path)
        l.oracle_solve.argtypes = [C.POINTER(capi.Problem), C.POINTER(capi.Result), C.c_int]
        l.oracle_solve_timed.argtypes = [C.POINTER(capi.Problem), C.POINTER(capi.Result), C.c_int,
                                         C.POINTER(C.c_double), C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_int32)]
        l.oracle_linearize.argtypes = [C.POINTER(capi.Problem), C.POINTER(capi.Linearization)]
        l.oracle_structure.argtypes = [C.POINTER(capi.Problem), C.POINTER(capi.Structure)]
        l.oracle_reduced_system.argtypes = [C.POINTER(capi.Problem), C.c_double, capi._dp, capi._dp, capi._dp,
                                            C.POINTER(C.c_int32), capi._dp, capi._dp]
        l.oracle_link_linearize.argtypes = [C.POINTER(capi.Problem), capi._dp, capi._dp, capi._dp]
        l.oracle_pose_oplus.argtypes = [C.c_int, capi._dp, capi._dp, capi._dp]
        _libs[name] = l
    return _libs[name]


def threads():
    return lib(True).oracle_threads()


def solve(w, threads=1, omp=False):
    keep = []
    p = capi.window_to_problem(w, keep)
    out = capi.ResultArrays(w)
    r = capi.Result()
    out.bind(r)
    lib(omp).oracle_solve(C.byref(p), C.byref(r), threads)
    return capi.result_to_dict(r, out)


def solve_timed(w, threads=1, omp=False):
    keep = []
    p = capi.window_to_problem(w, keep)
    out = capi.ResultArrays(w)
    r = capi.Result()
    out.bind(r)
    sec = C.c_double()
    log = (C.c_int32 * 256)()
    n = C.c_int32()
    lib(omp).oracle_solve_timed(C.byref(p), C.byref(r), threads, C.byref(sec), log, 256, C.byref(n))
    d = capi.result_to_dict(r, out)
    d["seconds"] = sec.value
    d["trials_per_iteration"] = list(log[: n.value])
    return d


def linearize(w):
    keep = []
    p = capi.window_to_problem(w, keep)
    lin, bufs = capi.new_linearization(int(w["n_edges"]))
    lib().oracle_linearize(C.byref(p), C.byref(lin))
    return bufs


def structure(w, edge_level=None, capacity=None):
    keep = []
    p = capi.window_to_problem(w, keep)
    s, bufs = capi.new_structure(w, edge_level, capacity)
    lib().oracle_structure(C.byref(p), C.byref(s))
    return capi.structure_to_dict(s, bufs)


def reduced_system(w, lam):
    keep = []
    p = capi.window_to_problem(w, keep)
    nmax = 6 * int(w["n_poses"])
    S = np.zeros((nmax, nmax))
    bs = np.zeros(nmax)
    x = np.zeros(nmax + 3 * int(w["n_points"]))
    n = C.c_int32()
    chi2 = C.c_double()
    lam0 = C.c_double()
    # S is written with leading dimension n, so hand over a flat buffer and reshape afterwards
    flat = np.zeros(nmax * nmax)
    st = lib().oracle_reduced_system(C.byref(p), float(lam), capi._ptr(flat, capi._dp), capi._ptr(bs, capi._dp),
                                     capi._ptr(x, capi._dp), C.byref(n), C.byref(chi2), C.byref(lam0))
    nn = n.value
    return dict(status=st, n=nn, S=flat[: nn * nn].reshape(nn, nn).copy(), b_s=bs[:nn].copy(), x=x, chi2=chi2.value,
                lambda_init=lam0.value)


def link_linearize(w):
    keep = []
    p = capi.window_to_problem(w, keep)
    K = int(w.get("n_links", 0))
    err, Ji, Jj = np.zeros((K, 6)), np.zeros((K, 6, 6)), np.zeros((K, 6, 6))
    lib().oracle_link_linearize(C.byref(p), capi._ptr(err, capi._dp), capi._ptr(Ji, capi._dp), capi._ptr(Jj, capi._dp))
    return dict(error=err, J_from=Ji, J_to=Jj)


def pose_oplus(tq, delta):
    """CameraPose::update of every row (the oracle's restatement of OptimizeTypeDefine.cpp:7-14)."""
    tq = np.ascontiguousarray(tq, dtype=np.float64)
    delta = np.ascontiguousarray(delta, dtype=np.float64)
    out = np.zeros_like(tq)
    lib().oracle_pose_oplus(len(tq), capi._ptr(tq, capi._dp), capi._ptr(delta, capi._dp), capi._ptr(out, capi._dp))
    return out

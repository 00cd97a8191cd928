"""Per-call host time of the resident local map over a 40-frame sequence: python tools/window_phases.py"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from visfs_b200 import capi, synth  # noqa: E402

big = len(sys.argv) > 1 and sys.argv[1] == "c2"
n_frames, views, per_new = (24, 10, 1000) if big else (40, 6, 50)
seq = synth.make_window(n_frames, per_new * n_frames, views=views, layout="consecutive", seed=synth.BASE_SEED + 11)
max_pts, max_obs = (32768, 262144) if big else (4096, 32768)
first_seen = np.full(seq["n_points"], 10**9)
np.minimum.at(first_seen, seq["edge_point"], seq["edge_pose"])
ba = capi.BundleAdjuster(0)
if 'ctx' in sys.argv:   # the handle first runs a pipelined batch (20 sub-handles, streams, page-locked staging)
    _p = ba.prepare_batch(synth.config_c3_windows(512), pinned=True, float_obs=True)
    ba.solve_packed(_p); ba.solve_packed(_p)
win = capi.ResidentWindow(ba, views + 1, max_pts, max_obs, fx=seq["fx"], fy=seq["fy"], cx=seq["cx"], cy=seq["cy"], bf=seq["bf"],
                          pixel_variance=seq["pixel_variance"], huber_delta=seq["huber_delta"], iterations=seq["iterations"])
lib, W = ba.lib, win.w
I64, F64, F32, U8 = capi._i64p, capi._dp, C.POINTER(C.c_float), capi._u8p
cap = max_obs
r_fid, r_tq = np.zeros(views + 1, dtype=np.int64), np.zeros((views + 1, 7))
r_op, r_of = np.zeros(cap, dtype=np.int64), np.zeros(cap, dtype=np.int64)
res = capi.WindowResult(frame_id=capi._ptr(r_fid, I64), pose_tq=capi._ptr(r_tq, F64), outlier_point_id=capi._ptr(r_op, I64),
                        outlier_frame_id=capi._ptr(r_of, I64), outlier_capacity=cap)
acc = {k: 0.0 for k in ("set_points", "insert_frame", "remove_frame", "solve", "remove_observations")}
frames, solves = [], 0
for f in range(n_frames):
    sel = np.nonzero(seq["edge_pose"] == f)[0]
    new = np.unique(seq["edge_point"][sel][first_seen[seq["edge_point"][sel]] == f])
    fid = int(seq["pose_id"][f])
    tq = np.ascontiguousarray(seq["pose_tq"][f])
    new_id = np.ascontiguousarray(seq["point_id"][new], dtype=np.int64); new_xyz = np.ascontiguousarray(seq["point_xyz"][new])
    pid = np.ascontiguousarray(seq["point_id"][seq["edge_point"][sel]], dtype=np.int64)
    ob = np.ascontiguousarray(seq["edge_obs"][sel], dtype=np.float32); kind = np.ascontiguousarray(seq["edge_kind"][sel], dtype=np.uint8)
    full = len(frames) >= views - 1
    t = time.perf_counter()
    if len(new_id):
        lib.visfs_ba_window_set_points(W, len(new_id), capi._ptr(new_id, I64), capi._ptr(new_xyz, F64), None)
    t1 = time.perf_counter()
    lib.visfs_ba_window_insert_frame(W, fid, capi._ptr(tq, F64), len(pid), capi._ptr(pid, I64), capi._ptr(ob, F32), capi._ptr(kind, U8))
    t2 = time.perf_counter()
    frames.append(f)
    if len(frames) > views:
        old = frames.pop(0)
        lib.visfs_ba_window_remove_frame(W, int(seq["pose_id"][old]))
    t3 = time.perf_counter()
    t4 = t5 = t3
    if len(frames) >= 2:
        lib.visfs_ba_window_solve(W, fid - 1, C.byref(res))
        t4 = time.perf_counter()
        if os.environ.get('VISFS_BA_WIN_TRACE'): print('py solve ms', round(1e3 * (t4 - t3), 3), 'n_out', res.n_outliers, 'E', res.n_edges, file=sys.stderr)
        k = min(res.n_outliers, cap)
        if k:
            lib.visfs_ba_window_remove_observations(W, k, capi._ptr(r_op, I64), capi._ptr(r_of, I64))
        t5 = time.perf_counter()
    if full and f >= views + 2:
        acc["set_points"] += t1 - t; acc["insert_frame"] += t2 - t1; acc["remove_frame"] += t3 - t2; acc["solve"] += t4 - t3
        acc["remove_observations"] += t5 - t4
        solves += 1
print({k: round(1e3 * v / solves, 4) for k, v in acc.items()}, "ms per frame;", solves, "frames; device LM ms of the last solve:", ba.timing()["total_ms"])

"""Window files for visfs_b200/host/optimizer_selftest (the C++ mirror of VISFS::Optimizer::Optimizer)."""
from __future__ import annotations

import os
import struct
import subprocess

import numpy as np

from visfs_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "visfs_b200", "host", "optimizer_selftest")


def write_window(path, w, feature_id_offset=0, extra_points=0, n_cameras=2, pose_subset=None):
    """Reference-form inputs of window `w` (synth.make_window): robot poses, world points, float key points
    and depths.  `extra_points` adds points3D entries that have no observation (they must come back NaN)."""
    P, L, E = int(w["n_poses"]), int(w["n_points"]), int(w["n_edges"])
    pose_rows = list(range(P)) if pose_subset is None else list(pose_subset)   # observations of the others stay in the file
    with open(path, "wb") as f:
        f.write(struct.pack("<8q", len(pose_rows), L + extra_points, E, int(w["ref_root_id"]) if w["ref_root_id"] >= 0 else 10**9, n_cameras,
                            int(w["iterations"]), int(w["solver"]), int(w["trust_region"])))
        f.write(struct.pack("<7d", synth.FX, synth.FY, synth.CX, synth.CY, synth.BASELINE, float(w["pixel_variance"]),
                            float(w["huber_delta"])))
        for i in pose_rows:
            f.write(struct.pack("<q", int(w["pose_id"][i])))
            f.write(np.ascontiguousarray(w["ref_T_wr"][i], dtype="<f8").tobytes())
        for l in range(L):
            f.write(struct.pack("<q3dq", int(w["point_id"][l]) + feature_id_offset, *map(float, w["point_xyz"][l]), int(w["point_fixed"][l])))
        for k in range(extra_points):
            f.write(struct.pack("<q3dq", 10**6 + k, 1.0, 2.0, 3.0, 0))
        for e in range(E):
            f.write(struct.pack("<2q4f", int(w["point_id"][w["edge_point"][e]]) + feature_id_offset, int(w["pose_id"][w["edge_pose"][e]]),
                                float(w["ref_kpt"][e, 0]), float(w["ref_kpt"][e, 1]), float(w["ref_depth"][e]), 0.0))
        K = int(w.get("n_links", 0))
        f.write(struct.pack("<q", K))
        if K:
            f.write(struct.pack("<d", float(w["odometry_variance"])))
            for k in range(K):
                f.write(struct.pack("<2q", int(w["pose_id"][w["link_from"][k]]), int(w["pose_id"][w["link_to"][k]])))
                f.write(np.ascontiguousarray(w["ref_link_T"][k], dtype="<f8").tobytes())


def _rdv(buf, off, dtype):
    n = struct.unpack_from("<q", buf, off)[0]
    off += 8
    a = np.frombuffer(buf, dtype=dtype, count=n, offset=off).copy()
    return a, off + n * np.dtype(dtype).itemsize


def run_marshal(in_path, out_path):
    subprocess.run([EXE, "marshal", in_path, out_path], check=True)
    buf = open(out_path, "rb").read()
    ok = struct.unpack_from("<q", buf, 0)[0]
    off = 8
    out = {"ok": ok}
    for name, dt in (("pose_tq", "<f8"), ("pose_id", "<i8"), ("pose_fixed", "u1"), ("point_xyz", "<f8"), ("point_id", "<i8"),
                     ("point_fixed", "u1"), ("edge_obs", "<f8"), ("edge_pose", "<i4"), ("edge_point", "<i4"), ("edge_kind", "u1"),
                     ("intr", "<f8")):
        out[name], off = _rdv(buf, off, dt)
    return out


def run_solve(in_path, out_path, mode="solve"):
    """mode "solve": VISFS::Optimizer::Optimizer::localOptimize; "resident": the same window through ResidentLocalMap."""
    p = subprocess.run([EXE, mode, in_path, out_path], check=True, capture_output=True, text=True)
    buf = open(out_path, "rb").read()
    off = 0
    n = struct.unpack_from("<q", buf, off)[0]; off += 8
    poses = {}
    for _ in range(n):
        pid = struct.unpack_from("<q", buf, off)[0]; off += 8
        poses[pid] = np.frombuffer(buf, dtype="<f8", count=16, offset=off).reshape(4, 4).copy(); off += 128
    n = struct.unpack_from("<q", buf, off)[0]; off += 8
    points = {}
    for _ in range(n):
        fid = struct.unpack_from("<q", buf, off)[0]; off += 8
        points[fid] = np.frombuffer(buf, dtype="<f8", count=3, offset=off).copy(); off += 24
    n = struct.unpack_from("<q", buf, off)[0]; off += 8
    outliers = np.frombuffer(buf, dtype="<i8", count=2 * n, offset=off).reshape(n, 2).copy()
    return dict(poses=poses, points=points, outliers=outliers, stdout=p.stdout, stderr=p.stderr)


def run_time(in_path, out_path):
    """localOptimize 20 times through the C++ mirror: {'local_optimize_ms_mean', 'local_optimize_ms_best', 'marshal_ms', ...}."""
    import json
    p = subprocess.run([EXE, "time", in_path, out_path], check=True, capture_output=True, text=True)
    return json.loads(p.stdout.strip().splitlines()[-1])

"""Landmark partition of one global BA problem for the multi-GPU path (BASELINE config C4).

Every rank keeps ALL poses and a contiguous range of landmarks together with their edges; the C library sums the
per-rank reduced camera systems with one ncclAllReduce per LM trial (visfs_ba_comm_init + VISFS_BA_FLAG_PARTITIONED).
Harness-side helper: the reference has no multi-process path, so there is no reference interface to mirror here.
"""
from __future__ import annotations

import numpy as np

from .capi import FLAG_PARTITIONED


def landmark_range(n_points: int, n_ranks: int, rank: int):
    return (n_points * rank) // n_ranks, (n_points * (rank + 1)) // n_ranks


def partition_window(w: dict, n_ranks: int, rank: int) -> dict:
    """The rank-local problem: all poses, landmarks [l0, l1) with local indices, their edges in the original order."""
    l0, l1 = landmark_range(int(w["n_points"]), n_ranks, rank)
    ep = np.asarray(w["edge_point"])
    keep = (ep >= l0) & (ep < l1)
    out = {k: v for k, v in w.items() if not k.startswith("ref_")}
    out["n_points"] = l1 - l0
    out["point_xyz"] = np.ascontiguousarray(w["point_xyz"][l0:l1])
    out["point_id"] = np.ascontiguousarray(w["point_id"][l0:l1])
    out["point_fixed"] = np.ascontiguousarray(w["point_fixed"][l0:l1])
    out["edge_obs"] = np.ascontiguousarray(w["edge_obs"][keep])
    out["edge_pose"] = np.ascontiguousarray(w["edge_pose"][keep])
    out["edge_point"] = np.ascontiguousarray(ep[keep] - l0).astype(np.int32)
    out["edge_kind"] = np.ascontiguousarray(w["edge_kind"][keep])
    out["n_edges"] = int(keep.sum())
    # odometry links are pose-pose constraints: every rank passes all of them, the library adds them on one rank only
    out["flags"] = int(w.get("flags", 0)) | FLAG_PARTITIONED
    out["part_range"] = (l0, l1)
    out["part_edge_index"] = np.nonzero(keep)[0]
    return out


def merge_results(w: dict, parts: list, results: list) -> dict:
    """Stitch per-rank results back into one result of the global problem (poses are identical on every rank)."""
    merged = {k: v for k, v in results[0].items() if k not in ("point_xyz", "edge_level", "n_outliers")}
    pts = np.zeros((int(w["n_points"]), 3))
    lev = np.zeros(int(w["n_edges"]), dtype=np.uint8)
    n_out = 0
    for p, r in zip(parts, results):
        l0, l1 = p["part_range"]
        pts[l0:l1] = r["point_xyz"]
        lev[p["part_edge_index"]] = r["edge_level"]
        n_out += int(r["n_outliers"])
    merged["point_xyz"], merged["edge_level"], merged["n_outliers"] = pts, lev, n_out
    return merged

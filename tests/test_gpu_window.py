"""The resident local map (visfs_ba_window_*) against the oracle re-solving every frame from scratch.

A 36-frame sequence is replayed the way VISFS runs (corelib/src/Estimator.cpp:216-254, 391-395; corelib/src/LocalMap.cpp):
every frame adds its new features and its observations, the oldest frame leaves when more than six are in the map, features
nobody observes any more are erased, the two-pass BA runs on what is left (features seen more than once), its poses and
points are written back (points only when displaced by less than 5 m), and the observations it culled are removed.
`Mirror` does the same with plain dicts + the CPU oracle; the window on the GPU must agree with it frame by frame."""
import numpy as np
import pytest

from tests import oracle_api as O
from visfs_b200 import capi, synth

pytestmark = pytest.mark.gpu

WINDOW = 6


class Mirror:
    """LocalMap's bookkeeping in Python; solve() = getSignaturePoses + getFeaturePosesAndObservations + localOptimize (oracle)."""

    def __init__(self, seq):
        self.seq, self.frames, self.points, self.obs, self.links = seq, {}, {}, {}, []

    def window(self, root):
        fids = sorted(self.frames)
        fidx = {f: i for i, f in enumerate(fids)}
        cnt = {}
        for (p, f) in self.obs:
            cnt[p] = cnt.get(p, 0) + 1
        pids = sorted(p for p, c in cnt.items() if c > 1)       # getObservedTimes() > 1 (LocalMap.cpp:277)
        pidx = {p: i for i, p in enumerate(pids)}
        keys = sorted(k for k in self.obs if k[0] in pidx)
        w = {k: self.seq[k] for k in ("fx", "fy", "cx", "cy", "bf", "pixel_variance", "huber_delta", "iterations", "solver", "trust_region")}
        w.update(n_poses=len(fids), n_points=len(pids), n_edges=len(keys),
                 pose_tq=np.array([self.frames[f] for f in fids]).reshape(-1, 7), pose_id=np.array(fids, dtype=np.int64),
                 pose_fixed=np.array([1 if f == root else 0 for f in fids], dtype=np.uint8),
                 point_xyz=np.array([self.points[p][0] for p in pids]).reshape(-1, 3), point_id=np.array(pids, dtype=np.int64),
                 point_fixed=np.array([self.points[p][1] for p in pids], dtype=np.uint8),
                 edge_obs=np.array([self.obs[k][0] for k in keys], dtype=np.float64).reshape(-1, 3),
                 edge_pose=np.array([fidx[k[1]] for k in keys], dtype=np.int32),
                 edge_point=np.array([pidx[k[0]] for k in keys], dtype=np.int32),
                 edge_kind=np.array([self.obs[k][1] for k in keys], dtype=np.uint8))
        if self.links:                                           # LocalMap::getSignatureLinks: consecutive signatures of the map
            lk = [(a, b, tq) for (a, b, tq) in self.links if a in fidx and b in fidx]
            w.update(n_links=len(lk), link_from=np.array([fidx[a] for a, _, _ in lk], dtype=np.int32),
                     link_to=np.array([fidx[b] for _, b, _ in lk], dtype=np.int32),
                     link_tq=np.array([tq for _, _, tq in lk], dtype=np.float64).reshape(-1, 7), odometry_variance=self.seq["odometry_variance"])
        return w, fids, pids, keys

    def solve(self, root):
        w, fids, pids, keys = self.window(root)
        r = O.solve(w)
        if r["status"] == 0:
            for i, f in enumerate(fids):
                self.frames[f] = r["pose_tq"][i].copy()
            for i, p in enumerate(pids):                          # Optimizer.cpp:343-351
                old, fixed = self.points[p]
                if np.linalg.norm(old - r["point_xyz"][i]) < 5.0:
                    self.points[p] = (r["point_xyz"][i].copy(), fixed)
        r["outliers"] = sorted(keys[e] for e in np.nonzero(r["edge_level"])[0])
        r["frame_id"], r["window"] = fids, w
        return r


def replay(ba, seed, stable_after=None, links=False, max_obs=6000):
    seq = synth.make_window(36, 700, views=5, layout="consecutive", seed=seed, outlier_frac=0.08, links="chain" if links else None)
    first_seen = {}
    for e in range(seq["n_edges"]):
        p, f = int(seq["edge_point"][e]), int(seq["edge_pose"][e])
        first_seen[p] = min(first_seen.get(p, 10**9), f)
    win = capi.ResidentWindow(ba, WINDOW + 1, 600, max_obs, fx=seq["fx"], fy=seq["fy"], cx=seq["cx"], cy=seq["cy"], bf=seq["bf"],
                              pixel_variance=seq["pixel_variance"], huber_delta=seq["huber_delta"], iterations=seq["iterations"])
    mir = Mirror(seq)
    if links:   # declared once, before any frame exists: a solve uses the links whose two frames are in the window at that time
        a, b = seq["pose_id"][seq["link_from"]], seq["pose_id"][seq["link_to"]]
        win.set_links(a, b, seq["link_tq"], seq["odometry_variance"])
        mir.links = [(int(x), int(y), tq.copy()) for x, y, tq in zip(a, b, seq["link_tq"])]
    solved, h2d_solves, full_bytes = 0, 0, 0
    for f in range(seq["n_poses"]):
        fid = int(seq["pose_id"][f])
        sel = np.nonzero(seq["edge_pose"] == f)[0]
        new = [p for p in np.unique(seq["edge_point"][sel]) if first_seen[int(p)] == f]
        if new:
            ids = seq["point_id"][new]
            win.set_points(ids, seq["point_xyz"][new], np.zeros(len(new), np.uint8))
            for p in new:
                mir.points[int(seq["point_id"][p])] = (seq["point_xyz"][p].copy(), 0)
        pid = seq["point_id"][seq["edge_point"][sel]]
        win.insert_frame(fid, seq["pose_tq"][f], pid, seq["edge_obs"][sel].astype(np.float32), seq["edge_kind"][sel])
        mir.frames[fid] = seq["pose_tq"][f].copy()
        for e, p in zip(sel, pid):
            mir.obs[(int(p), fid)] = (seq["edge_obs"][e].astype(np.float32).astype(np.float64), int(seq["edge_kind"][e]))
        if len(mir.frames) > WINDOW:                              # LocalMap::removeSignature, key-frame case: the oldest goes
            old = min(mir.frames)
            win.remove_frame(old)
            del mir.frames[old]
            mir.obs = {k: v for k, v in mir.obs.items() if k[1] != old}
            seen = {k[0] for k in mir.obs}
            gone = [p for p in mir.points if p not in seen]
            if gone:
                win.remove_points(gone)
                for p in gone:
                    del mir.points[p]
        if len(mir.frames) < 2:
            continue
        root = fid - 1                                            # Estimator.cpp:252
        want = mir.solve(root)
        got = win.solve(root)
        what = f"frame {fid}"
        assert got["status"] == want["status"] == 0, what
        assert (got["n_frames"], got["n_points"], got["n_edges"]) == (want["window"]["n_poses"], want["window"]["n_points"], want["window"]["n_edges"]), what
        assert got["frame_id"].tolist() == want["frame_id"], what
        assert got["iterations_run"] == want["iterations_run"] and got["trials_run"] == want["trials_run"], what
        assert got["outliers"] == want["outliers"], what
        assert abs(got["chi2_final"] - want["chi2_final"]) <= 1e-6 * want["chi2_final"], what
        np.testing.assert_allclose(got["pose_tq"], want["pose_tq"], rtol=1e-6, atol=1e-8, err_msg=what)
        ids = sorted(mir.points)
        np.testing.assert_allclose(win.get_points(ids), np.array([mir.points[p][0] for p in ids]), rtol=1e-6, atol=1e-7, err_msg=what)
        if want["outliers"]:                                      # LocalMap::updateLocalMap: culled observations leave the map
            win.remove_observations([k[0] for k in want["outliers"]], [k[1] for k in want["outliers"]])
            for k in want["outliers"]:
                del mir.obs[k]
        if stable_after is not None:                              # features seen often become STABLE = fixed (LocalMap.cpp:84-88)
            cnt = {}
            for (p, _f) in mir.obs:
                cnt[p] = cnt.get(p, 0) + 1
            st = [p for p, c in cnt.items() if c >= stable_after and mir.points[p][1] == 0]
            if st:
                win.set_points(st, np.array([mir.points[p][0] for p in st]), np.ones(len(st), np.uint8))
                for p in st:
                    mir.points[p] = (mir.points[p][0], 1)
        solved += 1
        h2d_solves += got["h2d_bytes"]
        w = want["window"]
        full_bytes += 56 * w["n_poses"] + 24 * w["n_points"] + 20 * w["n_edges"]
    win.close()
    return solved, h2d_solves, full_bytes


def test_sliding_window_equals_resolving_from_scratch_every_frame(ba):
    solved, h2d, full = replay(ba, seed=501)
    assert solved >= 30
    # a solve sends the frame table and the batch descriptors only: far less than the window the reference re-marshals
    assert h2d < 0.25 * full, (h2d, full)


def test_sliding_window_with_a_small_observation_pool(ba):
    # the append-only pool holds 900 observations, the map about 600 live ones: the pool is compacted on the device every few frames
    solved, _, _ = replay(ba, seed=504, max_obs=900)
    assert solved >= 30


def test_sliding_window_with_odometry_links(ba):
    # Optimizer.cpp:116-150 on the resident map: EdgePoseConstraint between consecutive frames of the window
    solved, _, _ = replay(ba, seed=503, links=True)
    assert solved >= 30


def test_sliding_window_with_stable_features(ba):
    solved, _, _ = replay(ba, seed=502, stable_after=4)
    assert solved >= 30


def test_window_rejects_bad_deltas_and_survives(ba):
    seq = synth.make_window(4, 40, layout="all", seed=503)
    win = capi.ResidentWindow(ba, 4, 64, 400, fx=seq["fx"], fy=seq["fy"], cx=seq["cx"], cy=seq["cy"], bf=seq["bf"])
    win.set_points(seq["point_id"], seq["point_xyz"])
    sel = np.nonzero(seq["edge_pose"] == 0)[0]
    pid = seq["point_id"][seq["edge_point"][sel]]
    ob = seq["edge_obs"][sel].astype(np.float32)
    win.insert_frame(1, seq["pose_tq"][0], pid, ob)
    with pytest.raises(capi.BAError, match="already in the window"):
        win.insert_frame(1, seq["pose_tq"][0], pid, ob)
    with pytest.raises(capi.BAError, match="never set"):
        win.insert_frame(2, seq["pose_tq"][1], [10**9], ob[:1])
    with pytest.raises(capi.BAError, match="two observations"):
        win.insert_frame(2, seq["pose_tq"][1], [pid[0], pid[0]], ob[:2])
    with pytest.raises(capi.BAError, match="no such frame"):
        win.remove_frame(77)
    for f in (1, 2, 3):
        s = np.nonzero(seq["edge_pose"] == f)[0]
        win.insert_frame(f + 1, seq["pose_tq"][f], seq["point_id"][seq["edge_point"][s]], seq["edge_obs"][s].astype(np.float32))
    with pytest.raises(capi.BAError, match="window full"):
        win.insert_frame(9, seq["pose_tq"][0], pid, ob)
    got = win.solve(3)
    ref = O.solve(seq)
    assert got["status"] == 0 and got["trials_run"] == ref["trials_run"]
    np.testing.assert_allclose(got["pose_tq"], ref["pose_tq"], rtol=1e-6, atol=1e-8)
    win.close()

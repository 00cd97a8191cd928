"""GPU test of the reference-facing plugin: VISFS::Optimizer::Optimizer::localOptimize (C++17, through the
C ABI) against the oracle plus the reference's write-back rules (Optimizer.cpp:320-358)."""
import numpy as np
import pytest

from tests import host_io
from tests import oracle_api as O
from visfs_b200 import synth

pytestmark = pytest.mark.gpu


def test_local_optimize_matches_oracle_and_write_back_rules(built, tmp_path):
    w = synth.make_window(6, 300, layout="consecutive", views=4, seed=81, mono_frac=0.2, fixed_point_frac=0.1, first_id=5)
    fin, fout = str(tmp_path / "w.bin"), str(tmp_path / "o.bin")
    host_io.write_window(fin, w, feature_id_offset=40, extra_points=3)
    got = host_io.run_solve(fin, fout)
    ref = O.solve(w)
    assert ref["status"] == 0
    # poses: T_wr = T_cw^-1 * T_rc^-1
    T_ref = synth.camera_state_to_robot(ref["pose_tq"])
    assert sorted(got["poses"]) == list(map(int, w["pose_id"]))
    for i, pid in enumerate(w["pose_id"]):
        assert np.allclose(got["poses"][int(pid)], T_ref[i], rtol=1e-6, atol=1e-8)
    # points: accepted when moved < 5 m, NaN for entries that never became a vertex
    for l, fid in enumerate(w["point_id"]):
        new, old = ref["point_xyz"][l], w["point_xyz"][l]
        want = new if np.linalg.norm(old - new) < 5.0 else old
        assert np.allclose(got["points"][int(fid) + 40], want, rtol=1e-6, atol=1e-8)
    for k in range(3):
        assert np.all(np.isnan(got["points"][10**6 + k]))
    # outliers: (feature id, pose id) of culled edges, in insertion order
    lvl = ref["edge_level"].astype(bool)
    want = np.stack([w["point_id"][w["edge_point"][lvl]] + 40, w["pose_id"][w["edge_pose"][lvl]]], axis=1)
    assert np.array_equal(got["outliers"], want)


def test_local_optimize_failure_returns_empty_map(built, tmp_path):
    w = synth.make_window(4, 50, layout="all", seed=82)
    w["point_xyz"][3] = np.nan
    fin, fout = str(tmp_path / "w.bin"), str(tmp_path / "o.bin")
    host_io.write_window(fin, w)
    got = host_io.run_solve(fin, fout)
    assert got["poses"] == {} and len(got["outliers"]) == 0
    assert "NANs" in got["stdout"] or "NANs" in got["stderr"]


def test_local_optimize_with_odometry_links(built, tmp_path):
    # the deployed configuration of the reference (SensorStrategy 3): visual edges + EdgePoseConstraint between the frames
    w = synth.make_window(6, 250, layout="all", seed=83, links="chain", first_id=3)
    fin, fout = str(tmp_path / "w.bin"), str(tmp_path / "o.bin")
    host_io.write_window(fin, w)
    got = host_io.run_solve(fin, fout)
    ref = O.solve(w)
    assert ref["status"] == 0
    T_ref = synth.camera_state_to_robot(ref["pose_tq"])
    assert sorted(got["poses"]) == list(map(int, w["pose_id"]))
    for i, pid in enumerate(w["pose_id"]):
        assert np.allclose(got["poses"][int(pid)], T_ref[i], rtol=1e-6, atol=1e-8)
    bare = {k: v for k, v in w.items() if not k.startswith("link") and k != "n_links"}
    assert not np.allclose(O.solve(bare)["pose_tq"], ref["pose_tq"], atol=1e-9), "the links do not matter: test is void"


def test_local_optimize_repeated_calls_reuse_their_buffers(built, tmp_path):
    # the Optimizer keeps its (page-locked) marshalling buffers from call to call: 23 calls on one object must all
    # succeed, and a C1-sized call has to stay in the low milliseconds (maps in, maps out)
    w = synth.config_c1()
    fin, fout = str(tmp_path / "w.bin"), str(tmp_path / "o.bin")
    host_io.write_window(fin, w)
    t = host_io.run_time(fin, fout)
    assert t["poses"] == 10 and t["edges"] == 20000
    assert 0.0 < t["marshal_ms"] < t["local_optimize_ms_best"] < 50.0


@pytest.mark.parametrize("links", [None, "chain"])
def test_resident_local_map_class_equals_local_optimize(built, tmp_path, links):
    # VISFS::Optimizer::ResidentLocalMap: the same window fed signature by signature as LocalMap's deltas,
    # then one localOptimize on the resident map, must return what Optimizer::localOptimize returns for the maps the reference
    # would have built: poses, culled (feature, signature) pairs, and the points after the 5 m write-back rule
    w = synth.make_window(6, 300, layout="consecutive", views=4, seed=84, mono_frac=0.2, fixed_point_frac=0.1, first_id=5, links=links)
    fin, f1, f2 = str(tmp_path / "w.bin"), str(tmp_path / "o1.bin"), str(tmp_path / "o2.bin")
    host_io.write_window(fin, w, feature_id_offset=40)
    a = host_io.run_solve(fin, f1, mode="solve")
    b = host_io.run_solve(fin, f2, mode="resident")
    assert sorted(a["poses"]) == sorted(b["poses"]) == list(map(int, w["pose_id"]))
    for pid in a["poses"]:
        assert np.allclose(a["poses"][pid], b["poses"][pid], rtol=1e-9, atol=1e-11)
    assert np.array_equal(a["outliers"], b["outliers"]) and len(a["outliers"]) > 0
    for fid in a["points"]:
        assert np.allclose(a["points"][fid], b["points"][fid], rtol=1e-9, atol=1e-11)

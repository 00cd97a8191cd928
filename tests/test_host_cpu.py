"""CPU tests of the C++17 host mirror (visfs_b200/host): map-of-maps -> flat arrays marshalling of
VISFS::Optimizer::Optimizer (reference Optimizer.cpp:100-223), checked bit-exactly against the python
marshalling that produced the synthetic windows."""
import os

import numpy as np

from tests import host_io
from visfs_b200 import synth


def test_marshal_matches_python(built, tmp_path):
    w = synth.make_window(6, 150, layout="consecutive", views=4, seed=71, mono_frac=0.3, fixed_point_frac=0.2, first_id=3)
    fin, fout = str(tmp_path / "w.bin"), str(tmp_path / "m.bin")
    host_io.write_window(fin, w, feature_id_offset=100, extra_points=2)
    m = host_io.run_marshal(fin, fout)
    assert m["ok"] == 1
    assert np.array_equal(m["pose_id"], w["pose_id"]) and np.array_equal(m["pose_fixed"], w["pose_fixed"])
    assert np.allclose(m["pose_tq"].reshape(-1, 7), w["pose_tq"], rtol=0, atol=1e-14)
    # points without observations get no vertex (Optimizer.cpp:158 iterates _wordReferences)
    assert np.array_equal(m["point_id"], w["point_id"] + 100)
    assert np.array_equal(m["point_xyz"].reshape(-1, 3), w["point_xyz"]) and np.array_equal(m["point_fixed"], w["point_fixed"])
    assert np.array_equal(m["edge_pose"], w["edge_pose"]) and np.array_equal(m["edge_point"], w["edge_point"])
    assert np.array_equal(m["edge_kind"], w["edge_kind"])
    assert np.array_equal(m["edge_obs"].reshape(-1, 3), w["edge_obs"])       # float-rounded disparity, bit-exact
    assert np.array_equal(m["intr"], [w["fx"], w["fy"], w["cx"], w["cy"], w["bf"]])


def test_marshal_single_camera_means_mono(built, tmp_path):
    w = synth.make_window(3, 20, layout="all", seed=72)
    fin, fout = str(tmp_path / "w.bin"), str(tmp_path / "m.bin")
    host_io.write_window(fin, w, n_cameras=1)           # baseLine stays 0 -> no stereo edge (Optimizer.cpp:181-184)
    m = host_io.run_marshal(fin, fout)
    assert np.all(m["edge_kind"] == 1) and m["intr"][4] == 0.0


def test_marshal_skips_observations_of_unknown_poses(built, tmp_path):
    w = synth.make_window(4, 30, layout="all", seed=73)
    fin, fout = str(tmp_path / "w.bin"), str(tmp_path / "m.bin")
    host_io.write_window(fin, w, pose_subset=[0, 2, 3])     # pose index 1 is not in _poses, its observations remain
    m = host_io.run_marshal(fin, fout)
    keep = w["edge_pose"] != 1                              # Optimizer.cpp:172 drops them
    assert np.array_equal(m["pose_id"], w["pose_id"][[0, 2, 3]])
    assert len(m["edge_pose"]) == int(keep.sum())
    remap = np.array([0, -1, 1, 2])
    assert np.array_equal(m["edge_pose"], remap[w["edge_pose"][keep]])
    assert np.array_equal(m["edge_obs"].reshape(-1, 3), w["edge_obs"][keep])

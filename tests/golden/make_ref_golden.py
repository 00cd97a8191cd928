"""Golden vectors taken from the REFERENCE'S OWN CODE (not from a restatement).

oracle/_ref/libvisfs_ref.so is the reference's unmodified corelib/src/Optimizer/g2o/OptimizeTypeDefine.cpp (+ its header and
utilite/include/Math.h) compiled in this container (oracle/Makefile).  This script runs it on seeded inputs and stores
inputs + outputs as tests/golden/ref_*.npz, so that the pin travels to machines without /root/reference (the GPU box):

  ref_edges.npz     EdgeStereo::computeError / linearizeOplus on the edges of three synthetic windows
                    (stereo, mixed with far points, loop trajectory 48 m from the origin) -> error, J_point, J_pose
  ref_oplus.npz     VertexPose::oplusImpl -> CameraPose::update + deltaQ on random poses / steps, including large steps and
                    negative-w inputs; CameraPose(R, t) (rotation matrix -> normalised quaternion, w >= 0)
  ref_links.npz     EdgePoseConstraint::computeError / linearizeOplus on random pose pairs and measurements

Run `python tests/golden/make_ref_golden.py` (needs /root/reference).  tests/test_ref_pin.py re-runs the library against
these files whenever it is present, so a stale fixture cannot go unnoticed."""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from tests import ref_api as R  # noqa: E402
from visfs_b200 import synth  # noqa: E402

EDGE_WINDOWS = {   # name -> make_window arguments (the windows are regenerated from these by the tests)
    "stereo_all": dict(n_poses=6, n_points=120, layout="all", seed=101),
    "mixed_consecutive": dict(n_poses=12, n_points=150, views=5, layout="consecutive", seed=102, mono_frac=0.3),
    "loop_far": dict(n_poses=40, n_points=200, views=4, layout="consecutive", trajectory="loop", seed=103),
}


def edge_cases():
    out = {}
    for name, kw in EDGE_WINDOWS.items():
        w = synth.make_window(**kw)
        r = R.edge_stereo_window(w)
        out[name + "_error"], out[name + "_J_point"], out[name + "_J_pose"] = r["error"], r["J_point"], r["J_pose"]
        out[name + "_edge_pose"], out[name + "_edge_point"] = w["edge_pose"], w["edge_point"]   # guards against a synth change
    return out


def random_poses(rng, n, flip=False):
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    q[:, 3] = np.abs(q[:, 3])
    if flip:
        q[::3] *= -1.0        # negative-w representatives of the same rotations
    t = rng.normal(scale=3.0, size=(n, 3))
    return np.concatenate([t, q], axis=1)


def oplus_cases():
    rng = np.random.default_rng(20261018)
    tq = random_poses(rng, 64)
    delta = rng.normal(scale=0.05, size=(64, 6))
    delta[48:] *= 40.0   # steps far outside the small-angle regime: deltaQ is first order and then normalised
    out = dict(tq=tq, delta=delta, tq_out=R.pose_oplus(tq, delta))
    # CameraPose(R, t): every branch of the rotation-matrix -> quaternion conversion (trace > 0 and the three others)
    raw = random_poses(rng, 48, flip=True)
    Rm = np.stack([synth.R_from_quat(q) for q in raw[:, 3:7]])
    out["R"], out["t"] = Rm, raw[:, :3]
    out["tq_from_R"] = np.stack([R.pose_from_matrix(Rm[i], raw[i, :3]) for i in range(len(raw))])
    unnorm = raw.copy()
    unnorm[:, 3:7] *= rng.uniform(0.5, 2.0, size=(len(raw), 1))
    out["tq_unnormalised"] = unnorm
    out["tq_normalised"] = np.stack([R.pose_normalize(unnorm[i]) for i in range(len(raw))])
    pw = rng.normal(scale=4.0, size=(len(raw), 3))
    out["pw"] = pw
    out["pc"] = np.stack([R.pose_map(out["tq_normalised"][i], pw[i])[0] for i in range(len(raw))])
    return out


def link_cases():
    rng = np.random.default_rng(20261019)
    a, b = random_poses(rng, 64), random_poses(rng, 64)
    b[:32] = a[:32] + np.concatenate([rng.normal(scale=0.2, size=(32, 3)), rng.normal(scale=0.02, size=(32, 4))], axis=1)
    b[:32, 3:7] /= np.linalg.norm(b[:32, 3:7], axis=1, keepdims=True)     # neighbouring key frames: the realistic case
    m = random_poses(rng, 64, flip=True)
    r = R.edge_pose_constraint(a, b, m)
    return dict(from_tq=a, to_tq=b, meas_tq=m, error=r["error"], J_from=r["J_from"], J_to=r["J_to"])


def main():
    np.savez_compressed(os.path.join(HERE, "ref_edges.npz"), **edge_cases())
    np.savez_compressed(os.path.join(HERE, "ref_oplus.npz"), **oplus_cases())
    np.savez_compressed(os.path.join(HERE, "ref_links.npz"), **link_cases())
    print("wrote ref_edges.npz ref_oplus.npz ref_links.npz")


if __name__ == "__main__":
    main()

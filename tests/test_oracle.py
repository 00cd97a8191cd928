"""CPU tests of the oracle (oracle/ba_oracle.cpp): hand-computed known answers, finite differences, the
reference's quirks, and the golden fixtures produced by the independent dense
numpy implementation in tests/golden/make_golden.py.  No GPU needed."""
import glob
import os

import numpy as np
import pytest

from tests import oracle_api as O
from tests.golden import make_golden as G
from visfs_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))


def one_edge_window(tq, point, obs, kind=0, delta=8.0):
    return dict(n_poses=1, n_points=1, n_edges=1, pose_tq=np.array([tq], float), pose_id=np.array([1]), pose_fixed=np.array([0], np.uint8),
                point_xyz=np.array([point], float), point_id=np.array([0]), point_fixed=np.array([0], np.uint8),
                edge_obs=np.array([obs], float), edge_pose=np.array([0], np.int32), edge_point=np.array([0], np.int32),
                edge_kind=np.array([kind], np.uint8), fx=420.0, fy=420.0, cx=320.0, cy=240.0, bf=21.0, pixel_variance=1.5,
                huber_delta=delta, iterations=10, solver=0, trust_region=0, flags=0)


# ------------------------------------------------------------------ known answers (computed by hand)
def test_known_answer_point_on_optical_axis():
    # identity pose, point (0,0,2): proj = (cx, cy, cx - bf/2) = (320, 240, 309.5)
    w = one_edge_window([0, 0, 0, 0, 0, 0, 1], [0, 0, 2], [321.0, 238.0, 310.0])
    lin = O.linearize(w)
    assert np.allclose(lin["error"][0], [1.0, -2.0, 0.5], atol=1e-13)
    assert lin["chi2"][0] == pytest.approx((1 + 4 + 0.25) / 1.5, rel=1e-14)
    assert lin["rho"][0] == lin["chi2"][0] and lin["weight"][0] == 1.0
    Jp = np.array([[-210.0, 0, 0, 0, -420.0, 0], [0, -210.0, 0, 420.0, 0, 0], [-210.0, 0, -5.25, 0, -420.0, 0]])
    Jl = np.array([[-210.0, 0, 0], [0, -210.0, 0], [-210.0, 0, -5.25]])
    assert np.allclose(lin["J_pose"][0], Jp, atol=1e-12)
    assert np.allclose(lin["J_point"][0], Jl, atol=1e-12)


def test_known_answer_huber_branch():
    # error (30, 0, 0): chi2 = 600 > delta^2 = 64 -> rho = 2*8*sqrt(600) - 64, w = 8 / sqrt(600)
    w = one_edge_window([0, 0, 0, 0, 0, 0, 1], [0, 0, 2], [350.0, 240.0, 309.5])
    lin = O.linearize(w)
    assert lin["chi2"][0] == pytest.approx(600.0, rel=1e-14)
    assert lin["rho"][0] == pytest.approx(16 * np.sqrt(600.0) - 64.0, rel=1e-14)
    assert lin["weight"][0] == pytest.approx(8.0 / np.sqrt(600.0), rel=1e-14)
    w0 = one_edge_window([0, 0, 0, 0, 0, 0, 1], [0, 0, 2], [350.0, 240.0, 309.5], delta=0.0)   # no kernel
    lin0 = O.linearize(w0)
    assert lin0["rho"][0] == lin0["chi2"][0] and lin0["weight"][0] == 1.0


def test_known_answer_mono_edge_has_two_rows():
    w = one_edge_window([0, 0, 0, 0, 0, 0, 1], [0.2, -0.1, 2], [321.0, 238.0, 12345.0], kind=1)
    lin = O.linearize(w)
    assert lin["error"][0][2] == 0.0
    assert np.all(lin["J_pose"][0][2] == 0.0) and np.all(lin["J_point"][0][2] == 0.0)
    ws = one_edge_window([0, 0, 0, 0, 0, 0, 1], [0.2, -0.1, 2], [321.0, 238.0, 300.0], kind=0)
    ls = O.linearize(ws)
    assert np.array_equal(lin["J_pose"][0][:2], ls["J_pose"][0][:2]) and np.array_equal(lin["error"][0][:2], ls["error"][0][:2])


# ------------------------------------------------------------------ finite differences and quirk B-1
def _err(w, tq=None, pt=None):
    w2 = dict(w)
    if tq is not None:
        w2["pose_tq"] = np.array([tq], float)
    if pt is not None:
        w2["point_xyz"] = np.array([pt], float)
    return O.linearize(w2)["error"][0].copy()


def test_jacobians_finite_differences_and_rotation_quirk():
    rng = np.random.default_rng(3)
    q = rng.normal(size=4); q /= np.linalg.norm(q); q *= np.sign(q[3])
    tq = np.array([0.3, -0.2, 0.5, *q])
    R = synth.R_from_quat(q)
    pc = np.array([0.4, -0.3, 3.0])
    pt = R.T @ (pc - tq[:3])
    w = one_edge_window(tq, pt, [300.0, 200.0, 290.0])
    lin = O.linearize(w)
    h = 1e-6
    # d e / d point and d e / d t agree with central differences
    for k in range(3):
        d = np.zeros(3); d[k] = h
        num = (_err(w, pt=pt + d) - _err(w, pt=pt - d)) / (2 * h)
        assert np.allclose(lin["J_point"][0][:, k], num, rtol=1e-6, atol=1e-5)
        tp, tm = tq.copy(), tq.copy()
        tp[k] += h; tm[k] -= h
        num = (_err(w, tq=tp) - _err(w, tq=tm)) / (2 * h)
        assert np.allclose(lin["J_pose"][0][:, k], num, rtol=1e-6, atol=1e-5)
    # the rotation columns are d e/d pc * (-[pc]x) (SE(3) left perturbation), which is NOT the
    # derivative of the reference's own oplus (t additive, q <- dq * q).  Both facts are asserted so nobody "fixes" it.
    fx = fy = 420.0; bf = 21.0
    x, y, z = pc
    de_dpc = -np.array([[fx / z, 0, -fx * x / z**2], [0, fy / z, -fy * y / z**2], [fx / z, 0, -fx * x / z**2 + bf / z**2]])
    skew = np.array([[0, -z, y], [z, 0, -x], [-y, x, 0]])
    assert np.allclose(lin["J_pose"][0][:, 3:], de_dpc @ (-skew), rtol=1e-12, atol=1e-9)
    num_rot = np.zeros((3, 3))
    for k in range(3):
        d = np.zeros(6); d[3 + k] = h
        num_rot[:, k] = (_err(w, tq=G.pose_oplus(tq, d)) - _err(w, tq=G.pose_oplus(tq, -d))) / (2 * h)
    assert not np.allclose(lin["J_pose"][0][:, 3:], num_rot, rtol=1e-2, atol=1e-2)


def test_pose_oplus_is_first_order_delta_q():
    # Math.h:277-287 + OptimizeTypeDefine.cpp:7-14, checked through one Gauss-Newton-free LM step is overkill;
    # the oracle's oplus is exercised against the numpy one through the golden fixtures.  Here: the numpy one.
    tq = np.array([1.0, 2.0, 3.0, 0.0, 0.0, 0.0, 1.0])
    out = G.pose_oplus(tq, np.array([0.1, 0.2, 0.3, 0.2, 0.0, 0.0]))
    assert np.allclose(out[:3], [1.1, 2.2, 3.3])
    assert np.allclose(out[3:], np.array([0.1, 0, 0, 1.0]) / np.sqrt(1.01))   # (w=1, x=0.1) normalised, not (sin, cos)


# ------------------------------------------------------------------ structure semantics (g2o index mapping)
def test_structure_small_example():
    # 3 poses (pose 1 fixed), 3 points (point 2 fixed); edges: p0:{0,1,2} p1:{1,2} p2(fixed):{1(fixed pose) ,2}
    w = synth.make_window(3, 3, layout="all", seed=1, root=None)
    w["pose_fixed"] = np.array([0, 1, 0], np.uint8)
    w["point_fixed"] = np.array([0, 0, 1], np.uint8)
    keep = np.array([1, 1, 1, 0, 1, 1, 0, 1, 1], bool)   # edges (point, pose): 00 01 02 | 11 12 | 21 22
    for k in ("edge_obs", "edge_pose", "edge_point", "edge_kind"):
        w[k] = np.ascontiguousarray(w[k][keep])
    w["n_edges"] = int(keep.sum())
    s = O.structure(w)
    assert list(s["pose_hidx"]) == [0, -1, 1]
    assert list(s["point_hidx"]) == [2, 3, -1]
    assert list(s["edge_active"]) == [1, 1, 1, 1, 1, 0, 1]           # (point 2, pose 1): both fixed -> inactive
    assert list(s["hpl_row"]) == [0, -1, 1, -1, 1, -1, -1]
    assert list(s["hpl_col"]) == [0, -1, 0, -1, 1, -1, -1]
    assert list(zip(s["schur_cols"], s["schur_rows"])) == [(0, 0), (1, 0), (1, 1)]
    assert (s["n_free_poses"], s["n_free_points"], s["n_active_edges"], s["n_hpl_blocks"]) == (2, 2, 6, 3)


def test_structure_pass2_pattern_keeps_level1_edges():
    # g2o's Schur pattern walks v->edges(), which still contains the culled (level 1) edges
    w = synth.make_window(3, 1, layout="all", seed=2, root=None)
    level = np.array([0, 1, 0], np.uint8)     # the point keeps poses 0 and 2
    s = O.structure(w, level)
    assert list(s["pose_hidx"]) == [0, -1, 1]
    assert list(zip(s["schur_cols"], s["schur_rows"])) == [(0, 0), (1, 0), (1, 1)]
    assert s["n_active_edges"] == 2


# ------------------------------------------------------------------ solver semantics
def test_schur_solution_equals_dense_normal_equations():
    w = synth.make_window(4, 30, layout="consecutive", views=3, seed=5, fixed_point_frac=0.2, mono_frac=0.2)
    lam = 2.5
    red = O.reduced_system(w, lam)
    # dense system from the numpy restatement
    level = np.zeros(w["n_edges"], np.uint8)
    s = O.structure(w)
    lin = O.linearize(w)
    F, NL = s["n_free_poses"], s["n_free_points"]
    n = 6 * F + 3 * NL
    H, b = np.zeros((n, n)), np.zeros(n)
    for e in range(w["n_edges"]):
        if not s["edge_active"][e]:
            continue
        cols, blocks = [], []
        hp, hl = s["pose_hidx"][w["edge_pose"][e]], s["point_hidx"][w["edge_point"][e]]
        if hp >= 0:
            cols.append(np.arange(6) + 6 * hp); blocks.append(lin["J_pose"][e])
        if hl >= 0:
            cols.append(np.arange(3) + 3 * (hl - F) + 6 * F); blocks.append(lin["J_point"][e])
        if not cols:
            continue
        J, c = np.concatenate(blocks, axis=1), np.concatenate(cols)
        wo = lin["weight"][e] / w["pixel_variance"]
        H[np.ix_(c, c)] += wo * J.T @ J
        b[c] -= wo * J.T @ lin["error"][e]
    x = np.linalg.solve(H + lam * np.eye(n), b)
    assert np.allclose(red["x"][:n], x, rtol=1e-9, atol=1e-12)
    # and the reduced system is the Schur complement of that matrix
    Hd = H + lam * np.eye(n)
    App, Apl, All = Hd[: 6 * F, : 6 * F], Hd[: 6 * F, 6 * F:], Hd[6 * F:, 6 * F:]
    S = App - Apl @ np.linalg.solve(All, Apl.T)
    assert np.allclose(red["S"], S, rtol=1e-9, atol=1e-9 * np.abs(S).max())
    assert red["lambda_init"] == pytest.approx(1e-5 * np.abs(np.diag(H)).max(), rel=1e-12)
    del level


def test_lm_converges_on_noise_free_window_and_lambda_shrinks_by_a_third():
    w = synth.make_window(4, 60, layout="all", seed=6, pixel_noise=0.0, outlier_frac=0.0, iterations=20)
    # float key points / depths leave ~1e-5 px of quantisation noise
    r = O.solve(w)
    assert r["status"] == 0 and r["chi2_final"] < 1e-3 * r["chi2_initial"]
    lam0 = O.reduced_system(w, 0.0)["lambda_init"]
    assert r["trials_run"][0] == r["iterations_run"][0] == 10
    # every accepted step scales lambda by a factor in [1/3, 2/3] (g2o's goodStepLowerScale / UpperScale)
    assert lam0 / 3.0 ** 10 * (1 - 1e-9) <= r["lambda_final"][0] <= lam0 * (2.0 / 3.0) ** 10 * (1 + 1e-9)


def test_outlier_rule_uses_delta_not_delta_squared():
    # culling threshold is chi2 > delta (8), Huber corner is chi2 = delta^2 (64)
    w = synth.make_window(4, 80, layout="all", seed=7, iterations=2)   # 1 + 1 iterations
    r = O.solve(w)
    wn = dict(w); wn["flags"] = 4   # NO_CULL
    rn = O.solve(wn)
    assert rn["n_outliers"] == 0 and r["n_outliers"] > 0
    # recompute the plain chi2 at the state after pass 1 (= final state of a single-pass run with the same iterations)
    w1 = dict(w); w1["flags"] = 2; w1["iterations"] = 1
    r1 = O.solve(w1)
    w_at = dict(w); w_at["pose_tq"], w_at["point_xyz"] = r1["pose_tq"], r1["point_xyz"]
    chi2 = O.linearize(w_at)["chi2"]
    assert np.array_equal(r["edge_level"].astype(bool), chi2 > 8.0)
    assert ((chi2 > 8.0) & (chi2 <= 64.0)).any()


def test_integer_division_of_iterations():
    w = synth.make_window(4, 40, layout="all", seed=8, iterations=7)
    assert O.solve(w)["iterations_run"] == [3, 3]
    w["iterations"] = 1
    r = O.solve(w)
    assert r["iterations_run"] == [0, 0] and r["n_outliers"] > 0


def test_pcg_matches_cholesky():
    w = synth.make_window(6, 120, layout="consecutive", views=4, seed=9)
    a = O.solve(w)
    w["solver"] = 2
    b = O.solve(w)
    assert a["iterations_run"] == b["iterations_run"] and np.array_equal(a["edge_level"], b["edge_level"])
    # g2o's PCG stops its first solve after init() at a relative residual of 1e-6, so the two differ a little
    assert np.allclose(a["pose_tq"], b["pose_tq"], rtol=1e-4, atol=5e-5)


def test_openmp_build_matches_sequential_build():
    w = synth.make_window(6, 200, layout="all", seed=10)
    a, b = O.solve(w), O.solve(w, threads=4, omp=True)
    assert a["iterations_run"] == b["iterations_run"] and np.array_equal(a["edge_level"], b["edge_level"])
    assert np.allclose(a["pose_tq"], b["pose_tq"], rtol=1e-9, atol=1e-12)
    assert a["chi2_final"] == pytest.approx(b["chi2_final"], rel=1e-9)


# ------------------------------------------------------------------ golden fixtures
def _load(path):
    z = np.load(path)
    w = {k[3:]: (z[k] if z[k].ndim else z[k].item()) for k in z.files if k.startswith("in_")}
    out = {k[4:]: (z[k] if z[k].ndim else z[k].item()) for k in z.files if k.startswith("out_")}
    return w, out


GOLDEN = sorted(p for p in glob.glob(os.path.join(HERE, "golden", "*.npz")) if not os.path.basename(p).startswith("ref_"))


def test_golden_fixtures_exist():
    assert len(GOLDEN) >= 5


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_matches_golden(path):
    w, out = _load(path)
    r = O.solve(w)
    assert r["status"] == out["status"]
    assert r["iterations_run"] == list(out["iterations_run"]) and r["trials_run"] == list(out["trials_run"])
    assert r["stop_reason"] == list(out["stop_reason"]) and r["n_outliers"] == out["n_outliers"]
    assert np.array_equal(r["edge_level"], out["edge_level"])
    assert r["chi2_pass1"] == pytest.approx(out["chi2_pass1"], rel=1e-9)
    assert r["chi2_final"] == pytest.approx(out["chi2_final"], rel=1e-8)
    assert np.allclose(r["pose_tq"], out["pose_tq"], rtol=1e-8, atol=1e-9)
    assert np.allclose(r["point_xyz"], out["point_xyz"], rtol=1e-8, atol=1e-9)
    assert np.allclose(r["lambda_final"], out["lambda_final"], rtol=1e-8)


@pytest.mark.parametrize("name", ["stereo_4x40", "rejecting_4x50"])
def test_golden_generator_reproduces_committed_fixture(name):
    w, out = _load(os.path.join(HERE, "golden", name + ".npz"))
    again = G.dense_lm(w)
    assert np.array_equal(again["edge_level"], out["edge_level"])
    assert np.allclose(again["pose_tq"], out["pose_tq"], rtol=1e-12, atol=1e-13)
    if name == "rejecting_4x50":
        assert sum(out["trials_run"]) > sum(out["iterations_run"])


# ---------------------------------------------------------------- odometry links (EdgePoseConstraint)
def _oplus(tq, d):
    """CameraPose::update (OptimizeTypeDefine.cpp:7-14) in numpy."""
    t = tq[:3] + d[:3]
    ax, ay, az, aw = d[3] / 2, d[4] / 2, d[5] / 2, 1.0
    bx, by, bz, bw = tq[3:7]
    q = np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz,
                  aw * bz + az * bw + ax * by - ay * bx, aw * bw - ax * bx - ay * by - az * bz])
    return np.concatenate([t, q / np.linalg.norm(q)])


def test_link_jacobians_are_the_derivatives_of_the_error_under_the_reference_oplus():
    # unlike the visual edge's rotation columns (Appendix B-1) the live "Left update" Jacobians of EdgePoseConstraint
    # (OptimizeTypeDefine.cpp:53-72) are exact: central differences through CameraPose::update pin the transcription
    w = synth.make_window(5, 40, layout="all", seed=5, links="chain")
    base = O.link_linearize(w)
    assert np.abs(base["error"]).max() > 1e-3
    h = 1e-6
    for k in range(w["n_links"]):
        for which, key in ((int(w["link_from"][k]), "J_from"), (int(w["link_to"][k]), "J_to")):
            J = np.zeros((6, 6))
            for a in range(6):
                d = np.zeros(6); d[a] = h
                wp = dict(w); wp["pose_tq"] = w["pose_tq"].copy(); wp["pose_tq"][which] = _oplus(w["pose_tq"][which], d)
                wm = dict(w); wm["pose_tq"] = w["pose_tq"].copy(); wm["pose_tq"][which] = _oplus(w["pose_tq"][which], -d)
                J[:, a] = (O.link_linearize(wp)["error"][k] - O.link_linearize(wm)["error"][k]) / (2 * h)
            assert np.abs(J - base[key][k]).max() < 1e-8, (k, key)


def test_link_error_is_zero_at_its_own_measurement_and_links_pull_the_solution():
    w = synth.make_window(4, 60, layout="all", seed=6, links="chain", link_noise=(0.0, 0.0), pose_noise=(0.0, 0.0))
    assert np.abs(O.link_linearize(w)["error"]).max() < 1e-12      # ground-truth poses, noise-free measurement
    w = synth.make_window(6, 200, layout="all", seed=7, links="chain")
    with_links = O.solve(w)
    bare = {k: v for k, v in w.items() if not k.startswith("link") and k != "n_links"}
    without = O.solve(bare)
    assert with_links["status"] == 0 and with_links["chi2_initial"] > without["chi2_initial"]
    assert not np.allclose(with_links["pose_tq"], without["pose_tq"], atol=1e-9)


def test_links_keep_a_pose_without_visual_edges_active():
    w = synth.make_window(5, 80, layout="all", seed=8, links="chain")
    keep = w["edge_pose"] != 1
    for k in ("edge_obs", "edge_pose", "edge_point", "edge_kind"):
        w[k] = np.ascontiguousarray(w[k][keep])
    w["n_edges"] = int(keep.sum())
    s = O.structure(w)
    assert s["pose_hidx"][1] >= 0
    bare = {k: v for k, v in w.items() if not k.startswith("link") and k != "n_links"}
    assert O.structure(bare)["pose_hidx"][1] == -1


def test_float_observations_are_the_same_problem():
    # edge_obs_f32 (include/visfs_ba.h): the reference's observations are floats widened to double, so handing them over as
    # floats is exact — the oracle must return bit-identical results
    from visfs_b200 import capi
    w = synth.make_window(6, 200, layout="consecutive", views=4, seed=77, mono_frac=0.2)
    a, b = O.solve(w), O.solve(capi.with_float_observations(w))
    assert a["trials_run"] == b["trials_run"] and a["chi2_final"] == b["chi2_final"]
    assert np.array_equal(a["pose_tq"], b["pose_tq"]) and np.array_equal(a["edge_level"], b["edge_level"])
    w2 = dict(w); w2["edge_obs"] = w["edge_obs"] + 1e-9
    with pytest.raises(ValueError):
        capi.with_float_observations(w2)

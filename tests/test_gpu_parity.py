"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded windows.

Gates (BASELINE.json): index maps and block patterns bit-exact; per-edge residuals and Jacobians
1e-9 relative; final chi2, poses and points 1e-6 relative after the same iteration / trial counts.
"""
import numpy as np
import pytest

from tests import oracle_api as O
from visfs_b200 import capi, synth

pytestmark = pytest.mark.gpu

RTOL_EDGE = 1e-9
RTOL_FINAL = 1e-6


def rejecting_window(seed, pose_noise, iterations=20):
    """Close landmarks and a bad initial guess: LM has to reject steps and raise the damping."""
    return synth.make_window(5, 150, layout="all", seed=seed, pose_noise=pose_noise, point_noise=0.5, iterations=iterations,
                             depth_range=(1.0, 6.0))


def rel_close(a, b, rtol, what=""):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = max(float(np.max(np.abs(b))) if b.size else 0.0, 1e-300)
    err = float(np.max(np.abs(a - b))) / scale if b.size else 0.0
    assert err <= rtol, f"{what}: max abs diff / max|ref| = {err:.3e} > {rtol:.1e}"


def elem_close(a, b, rtol, atol, what=""):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    bad = np.abs(a - b) > atol + rtol * np.abs(b)
    assert not bad.any(), f"{what}: {int(bad.sum())} of {bad.size} entries differ, worst {np.max(np.abs(a - b)):.3e}"


def check_solution(got, ref, what=""):
    assert got["status"] == ref["status"], (what, got["status"], ref["status"])
    for k in ("iterations_run", "trials_run", "stop_reason", "n_free_poses", "n_free_points"):
        assert got[k] == ref[k], f"{what}: {k} {got[k]} != {ref[k]}"
    assert got["n_outliers"] == ref["n_outliers"], what
    assert np.array_equal(got["edge_level"], ref["edge_level"]), f"{what}: outlier sets differ"
    for k in ("chi2_initial", "chi2_pass1", "chi2_final"):
        assert abs(got[k] - ref[k]) <= RTOL_FINAL * abs(ref[k]) + 1e-12, f"{what}: {k} {got[k]!r} vs {ref[k]!r}"
    # poses / points: 1e-6 relative, with an absolute floor at 1e-6 of the coordinate scale
    elem_close(got["pose_tq"], ref["pose_tq"], RTOL_FINAL, RTOL_FINAL * max(1.0, float(np.abs(ref["pose_tq"]).max())) * 1e-2,
               what + " poses")
    elem_close(got["point_xyz"], ref["point_xyz"], RTOL_FINAL, RTOL_FINAL * max(1.0, float(np.abs(ref["point_xyz"]).max())) * 1e-2,
               what + " points")


# ---------------------------------------------------------------- per-edge residuals and Jacobians
@pytest.mark.parametrize("mono_frac", [0.0, 0.3])
def test_linearize_matches_oracle(ba, mono_frac):
    w = synth.make_window(7, 500, layout="consecutive", views=4, seed=11, mono_frac=mono_frac)
    got, ref = ba.linearize(w), O.linearize(w)
    for k in ("error", "chi2", "rho", "weight", "J_point", "J_pose"):
        elem_close(got[k], ref[k], RTOL_EDGE, 1e-9 * float(np.abs(ref[k]).max()) * 1e-3, k)
    assert (ref["weight"] < 1.0).any(), "test window has no Huber-downweighted edge"


def test_linearize_unsorted_edges_keep_caller_order(ba):
    w = synth.make_window(5, 200, layout="all", seed=12, shuffle_edges=True)
    got, ref = ba.linearize(w), O.linearize(w)
    elem_close(got["error"], ref["error"], RTOL_EDGE, 1e-12, "error")
    elem_close(got["J_pose"], ref["J_pose"], RTOL_EDGE, 1e-12, "J_pose")


# ---------------------------------------------------------------- structure: bit-exact
def _structure_case(ba, w, level=None):
    got, ref = ba.structure(w, level), O.structure(w, level)
    for k in ("pose_hidx", "point_hidx", "edge_active", "hpl_row", "hpl_col", "schur_rows", "schur_cols"):
        assert np.array_equal(got[k], ref[k]), f"{k} differs"
    for k in ("n_schur_blocks", "n_free_poses", "n_free_points", "n_active_edges", "n_hpl_blocks"):
        assert got[k] == ref[k], (k, got[k], ref[k])


def test_structure_all_views(ba):
    _structure_case(ba, synth.make_window(10, 400, layout="all", seed=21))


def test_structure_consecutive_views_fixed_points(ba):
    _structure_case(ba, synth.make_window(20, 3000, layout="consecutive", views=6, seed=22, fixed_point_frac=0.3))


def test_structure_with_culled_edges_and_dead_vertices(ba):
    w = synth.make_window(8, 600, layout="random", views=3, seed=23, fixed_point_frac=0.2)
    rng = np.random.default_rng(5)
    level = (rng.random(w["n_edges"]) < 0.4).astype(np.uint8)
    level[w["edge_pose"] == 2] = 1                  # a pose that loses every edge
    level[w["edge_point"] < 20] = 1                 # landmarks that lose every edge
    _structure_case(ba, w, level)


def test_structure_no_fixed_pose(ba):
    _structure_case(ba, synth.make_window(6, 200, layout="all", seed=24, root=None))


def test_structure_unsorted(ba):
    _structure_case(ba, synth.make_window(6, 300, layout="random", views=4, seed=25, shuffle_edges=True))


# ---------------------------------------------------------------- Hessian / Schur / solve of one trial
@pytest.mark.parametrize("kw", [dict(layout="all"), dict(layout="consecutive", views=5, mono_frac=0.3, fixed_point_frac=0.25)])
def test_reduced_system_matches_oracle(ba, kw):
    w = synth.make_window(9, 700, seed=31, **kw)
    lam = 3.7
    got, ref = ba.debug_trial(w, lam), O.reduced_system(w, lam)
    assert got["n"] == ref["n"]
    rel_close(got["chi2"], ref["chi2"], 1e-12, "chi2 at the input state")
    rel_close(got["S"], ref["S"], 1e-10, "reduced camera system")
    rel_close(got["b_s"], ref["b_s"], 1e-10, "reduced rhs")
    rel_close(got["x_pose"], ref["x"][: ref["n"]], 1e-7, "pose step")
    lam0 = ba.debug_trial(w, -1.0)["lambda_used"]
    rel_close(lam0, ref["lambda_init"], 1e-12, "initial damping")


# ---------------------------------------------------------------- full two-pass solves
def test_solve_small_stereo_window(ba):
    w = synth.make_window(6, 300, layout="all", seed=41)
    check_solution(ba.solve(w), O.solve(w), "6x300")


def test_solve_c1(ba):
    w = synth.config_c1()
    got, ref = ba.solve(w), O.solve(w)
    check_solution(got, ref, "C1")
    assert ref["n_outliers"] > 500 and ref["chi2_final"] < ref["chi2_initial"]


def test_solve_c1_with_fixed_points(ba):
    w = synth.config_c1(fixed_point_frac=0.3)
    check_solution(ba.solve(w), O.solve(w), "C1 30% fixed points")


def test_solve_c2_mixed_mono_stereo(ba):
    w = synth.config_c2()
    check_solution(ba.solve(w), O.solve(w), "C2")


def test_solve_with_rejected_steps(ba):
    # a badly initialised window: LM has to reject steps and raise the damping
    cases = [(102, (0.3, np.deg2rad(6.0)), 20), (110, (0.3, np.deg2rad(6.0)), 20), (130, (0.3, np.deg2rad(6.0)), 20),
             (139, (0.3, np.deg2rad(6.0)), 20)]
    for seed, pn, iters in cases:
        w = rejecting_window(seed, pn, iters)
        ref = O.solve(w)
        assert sum(ref["trials_run"]) > sum(ref["iterations_run"]), "window does not exercise the reject path"
        check_solution(ba.solve(w), ref, f"rejections seed {seed}")


def test_solve_chaotic_window_same_decisions(ba):
    # 20 degree / 1 m initial error: the first pass needs a 4-trial iteration and ends far from convergence; rounding
    # differences are amplified to ~1e-5 in chi2 there, so only the discrete decisions and a loose chi2 are compared
    w = rejecting_window(100, (1.0, np.deg2rad(20.0)), 10)
    got, ref = ba.solve(w), O.solve(w)
    for k in ("status", "iterations_run", "trials_run", "stop_reason", "n_free_poses"):
        assert got[k] == ref[k], k
    assert abs(got["chi2_pass1"] - ref["chi2_pass1"]) <= 1e-3 * ref["chi2_pass1"]


def test_solve_everything_culled_second_pass_empty(ba):
    w = rejecting_window(101, (1.0, np.deg2rad(20.0)))
    ref = O.solve(w)
    assert ref["stop_reason"][1] == 3 and ref["n_outliers"] == w["n_edges"]
    check_solution(ba.solve(w), ref, "all edges culled")


def test_solve_gauss_newton(ba):
    # undamped steps send ill-observed landmarks to 1e13 m in any implementation (g2o included); the
    # parity window therefore has no gross outliers and a good initial structure
    w = synth.make_window(6, 300, layout="all", seed=44, trust_region=1, outlier_frac=0.0, point_noise=0.02)
    check_solution(ba.solve(w), O.solve(w), "gauss-newton")


def test_solve_pcg(ba):
    w = synth.make_window(8, 400, layout="consecutive", views=5, seed=45, solver=2)
    check_solution(ba.solve(w), O.solve(w), "pcg")


def test_solve_no_robust_kernel_single_pass(ba):
    w = synth.make_window(6, 300, layout="all", seed=46, huber_delta=0.0, outlier_frac=0.0)
    got, ref = ba.solve(w), O.solve(w)
    check_solution(got, ref, "no kernel")
    assert got["stop_reason"][1] == 0 and got["n_outliers"] == 0


def test_solve_unsorted_edges(ba):
    w = synth.make_window(6, 300, layout="random", views=4, seed=47, shuffle_edges=True)
    check_solution(ba.solve(w), O.solve(w), "unsorted")


def test_solve_odd_iterations_and_one_iteration(ba):
    for it in (7, 1):
        w = synth.make_window(5, 200, layout="all", seed=48, iterations=it)
        check_solution(ba.solve(w), O.solve(w), f"iterations={it}")


def test_solve_no_fixed_pose_gauge_free(ba):
    w = synth.make_window(6, 300, layout="all", seed=49, root=None)
    check_solution(ba.solve(w), O.solve(w), "no fixed pose")


def test_solve_ragged_inputs(ba):
    # landmarks without edges, a pose without edges, two-pose window
    w = synth.make_window(5, 120, layout="random", views=2, seed=50)
    keep = (w["edge_point"] % 7 != 0) & (w["edge_pose"] != 1)
    for k in ("edge_obs", "edge_pose", "edge_point", "edge_kind"):
        w[k] = np.ascontiguousarray(w[k][keep])
    w["n_edges"] = int(keep.sum())
    check_solution(ba.solve(w), O.solve(w), "ragged")
    w2 = synth.make_window(2, 60, layout="all", seed=51)
    check_solution(ba.solve(w2), O.solve(w2), "two poses")


def test_solve_numeric_failure_reports_pass1(ba):
    w = synth.make_window(4, 50, layout="all", seed=52)
    w["point_xyz"][3] = np.nan
    got, ref = ba.solve(w), O.solve(w)
    assert got["status"] == ref["status"] == 3


def test_solve_batch_matches_individual_solves(ba):
    ws = [synth.make_window(6, 200 + 37 * k, layout="all" if k % 2 == 0 else "consecutive", views=4, seed=60 + k,
                            fixed_point_frac=0.1 * (k % 3)) for k in range(7)]
    ws.append(rejecting_window(102, (0.3, np.deg2rad(6.0))))
    got = ba.solve_batch(ws)
    for k, (g, w) in enumerate(zip(got, ws)):
        check_solution(g, O.solve(w), f"batch window {k}")


def test_resident_rerun_is_deterministic(ba):
    w = synth.config_c1()
    ba.upload([w])
    ba.run_resident()
    a = ba.download()[0]
    a = {k: (v.copy() if hasattr(v, "copy") else v) for k, v in a.items()}
    ba.run_resident()
    b = ba.download()[0]
    assert np.array_equal(a["pose_tq"], b["pose_tq"]) and np.array_equal(a["point_xyz"], b["point_xyz"])
    assert a["chi2_final"] == b["chi2_final"]
    t = ba.timing()
    assert t["lm_iterations"] == sum(b["iterations_run"]) and t["total_ms"] > 0


# ---------------------------------------------------------------- full-size properties
def test_c3_batch_properties(ba):
    ws = synth.config_c3_windows(16)
    got = ba.solve_batch(ws)
    ref0 = O.solve(ws[0])
    check_solution(got[0], ref0, "C3 window 0")
    for g in got:
        assert g["status"] == 0 and g["chi2_final"] < g["chi2_pass1"] < g["chi2_initial"]
        assert np.isfinite(g["pose_tq"]).all() and np.isfinite(g["point_xyz"]).all()
        assert np.allclose(np.linalg.norm(g["pose_tq"][:, 3:7], axis=1), 1.0, atol=1e-12)
    # idempotence of the fixed gauge: the fixed pose never moves
    for g, w in zip(got, ws):
        fixed = w["pose_fixed"].astype(bool)
        # (quaternion to 2 ulp: like the CameraPose constructor the library re-normalises every input quaternion)
        assert np.array_equal(g["pose_tq"][fixed, :3], w["pose_tq"][fixed, :3])
        assert np.allclose(g["pose_tq"][fixed, 3:], w["pose_tq"][fixed, 3:], rtol=0, atol=5e-16)


def test_pipelined_batch_groups_match_oracle(ba):
    # >= 32 windows: visfs_ba_solve_batch cuts the batch into groups with their own streams and host threads
    ws = [synth.make_window(5 + k % 3, 150 + 11 * k, layout="all" if k % 2 == 0 else "consecutive", views=4, seed=300 + k)
          for k in range(70)]
    ws[33] = rejecting_window(102, (0.3, np.deg2rad(6.0)))
    got = ba.solve_batch(ws)
    for k in (0, 1, 17, 33, 34, 52, 69):
        check_solution(got[k], O.solve(ws[k]), f"pipelined window {k}")
    t = ba.timing()
    assert t["lm_iterations"] == sum(sum(g["iterations_run"]) for g in got)
    assert t["h2d_bytes"] > 0 and t["d2h_bytes"] > 0


def test_page_locked_caller_arrays_take_the_direct_dma_path(ba):
    # arrays in visfs_ba_host_alloc memory are DMA'd without staging: bit-identical results to the pageable (staged) route,
    # for one window, a small batch and a pipelined batch; window sizes straddle the 32 KB direct-copy threshold
    ws = [synth.make_window(5 + k % 3, 120 + 60 * k, layout="all" if k % 2 == 0 else "consecutive", views=4, seed=700 + k,
                            mono_frac=0.2 if k % 3 == 0 else 0.0, fixed_point_frac=0.1) for k in range(40)]
    for sub in (ws[:1], ws[5:12], ws):
        staged = ba.solve_batch(sub)
        packed = ba.prepare_batch(sub, pinned=True)
        ba.solve_packed(packed)
        direct = ba.packed_results(packed)
        for a, b in zip(staged, direct):
            assert a["status"] == b["status"] and a["trials_run"] == b["trials_run"] and a["chi2_final"] == b["chi2_final"]
            assert np.array_equal(a["pose_tq"], b["pose_tq"]) and np.array_equal(a["point_xyz"], b["point_xyz"])
            assert np.array_equal(a["edge_level"], b["edge_level"])
    check_solution(direct[39], O.solve(ws[39]), "page-locked window 39")
    # the resident API: upload returns only after the caller's arrays have been read
    packed = ba.prepare_batch(ws[5:12], pinned=True)
    n, probs, res, outs, keep = packed
    ba._check(ba.lib.visfs_ba_upload(ba.h, n, probs))
    for k in range(n):                               # scribbling over the inputs now must not matter
        np.ctypeslib.as_array(probs[k].edge_obs, shape=(probs[k].n_edges * 3,))[:] = 0.0
    ba.run_resident()
    ba._check(ba.lib.visfs_ba_download(ba.h, n, res))
    for a, b in zip(ba.solve_batch(ws[5:12]), ba.packed_results(packed)):
        assert a["chi2_final"] == b["chi2_final"] and np.array_equal(a["pose_tq"], b["pose_tq"])


def test_float_observations_travel_as_floats_and_change_nothing(ba):
    # edge_obs_f32: 12 instead of 24 bytes per edge over PCIe; bit-identical results for one window, a pipelined batch and a
    # batch that mixes both forms (the float windows are then widened while packing)
    ws = [synth.make_window(5 + k % 3, 150 + 20 * k, layout="all" if k % 2 == 0 else "consecutive", views=4, seed=800 + k,
                            mono_frac=0.2 if k % 3 == 0 else 0.0) for k in range(40)]
    fw = [capi.with_float_observations(w) for w in ws]
    ref = ba.solve_batch(ws)
    bytes_double = ba.timing()["h2d_bytes"]
    got = ba.solve_batch(fw)
    bytes_float = ba.timing()["h2d_bytes"]
    n_edges = sum(w["n_edges"] for w in ws)
    assert bytes_double - bytes_float >= 12 * n_edges - 4096
    mixed = ba.solve_batch([fw[k] if k % 2 else ws[k] for k in range(40)])
    pairs = [(g, ref[k]) for k in range(40) for g in (got[k], mixed[k])]
    pairs.append((ba.solve(fw[7]), ba.solve(ws[7])))       # (a single window is cut into chunks differently from a batch)
    for g, r in pairs:
        assert g["trials_run"] == r["trials_run"] and g["chi2_final"] == r["chi2_final"]
        assert np.array_equal(g["pose_tq"], r["pose_tq"]) and np.array_equal(g["point_xyz"], r["point_xyz"])
        assert np.array_equal(g["edge_level"], r["edge_level"])
    packed = ba.prepare_batch(ws, pinned=True, float_obs=True)       # page-locked floats: the direct DMA route
    ba.solve_packed(packed)
    for a, b in zip(ba.packed_results(packed), ref):
        assert a["chi2_final"] == b["chi2_final"] and np.array_equal(a["pose_tq"], b["pose_tq"])


def test_malformed_problems_are_rejected_with_a_message_and_the_handle_survives(ba):
    # the host-side validation is load-bearing (an out-of-range index would be an illegal address on the device): every
    # malformed input comes back as a status + message, nothing is launched, and the handle solves the next window
    good = synth.make_window(5, 120, layout="all", seed=910)

    def bad(match, **changes):
        w = dict(good)
        for k, v in changes.items():
            w[k] = v(good[k].copy()) if callable(v) else v
        with pytest.raises(capi.BAError, match=match):
            ba.solve(w)

    def poke(i, val):
        def f(a):
            a[i] = val
            return a
        return f

    bad("edge index out of range", edge_pose=poke(7, 5))
    bad("edge index out of range", edge_pose=poke(0, -1))
    bad("edge index out of range", edge_point=poke(3, 120))
    bad("pixel_variance", pixel_variance=0.0)
    # one observation per (point, pose): a repeated edge in a sorted list, and in an unsorted one (found after
    # the device sort), is refused instead of racing in the per-pose accumulators
    bad("duplicate", edge_pose=poke(1, int(good["edge_pose"][0])))
    perm = np.random.default_rng(3).permutation(int(good["n_edges"]))
    dup = {k: good[k][perm].copy() for k in ("edge_obs", "edge_pose", "edge_point", "edge_kind")}
    dup["edge_pose"][5], dup["edge_point"][5] = dup["edge_pose"][400], dup["edge_point"][400]
    bad("duplicate", **dup)
    bad("pose_id not strictly ascending", pose_id=poke(2, int(good["pose_id"][1])))
    bad("point_id not strictly ascending", point_id=poke(10, int(good["point_id"][9])))
    linked = synth.make_window(5, 120, layout="all", seed=910, links="chain")
    w = dict(linked); w["link_to"] = linked["link_to"].copy(); w["link_to"][1] = 9
    with pytest.raises(capi.BAError, match="odometry link index out of range"):
        ba.solve(w)
    w = dict(linked); w["odometry_variance"] = 0.0
    with pytest.raises(capi.BAError, match="odometry_variance"):
        ba.solve(w)
    # a batch with one bad window is refused as a whole, naming the window
    ws = [synth.make_window(5, 100 + k, layout="all", seed=920 + k) for k in range(40)]
    ws[37] = dict(ws[37]); ws[37]["edge_point"] = ws[37]["edge_point"].copy(); ws[37]["edge_point"][5] = 10**6
    with pytest.raises(capi.BAError, match="problem 37: edge index out of range"):
        ba.solve_batch(ws)
    check_solution(ba.solve(good), O.solve(good), "after the rejected inputs")


# ---------------------------------------------------------------- odometry links (EdgePoseConstraint)
def test_links_structure_pattern(ba):
    w = synth.make_window(9, 300, layout="random", views=2, seed=401, links="chain")
    _structure_case(ba, w)
    ref = O.structure(w)
    chain = {(i, i + 1) for i in range(ref["n_free_poses"] - 1)}
    assert chain <= set(zip(ref["schur_rows"].tolist(), ref["schur_cols"].tolist())), "links must add their pose-pose blocks"


@pytest.mark.parametrize("kw", [dict(layout="all"), dict(layout="consecutive", views=4, mono_frac=0.3, fixed_point_frac=0.2)])
def test_links_reduced_system_matches_oracle(ba, kw):
    w = synth.make_window(8, 500, seed=402, links="chain", **kw)
    lam = 1.3
    got, ref = ba.debug_trial(w, lam), O.reduced_system(w, lam)
    assert got["n"] == ref["n"]
    rel_close(got["chi2"], ref["chi2"], 1e-12, "chi2 at the input state (visual + odometry)")
    rel_close(got["S"], ref["S"], 1e-10, "reduced camera system with pose-pose blocks")
    rel_close(got["b_s"], ref["b_s"], 1e-10, "reduced rhs")
    rel_close(got["x_pose"], ref["x"][: ref["n"]], 1e-7, "pose step")
    rel_close(ba.debug_trial(w, -1.0)["lambda_used"], ref["lambda_init"], 1e-12, "initial damping")
    bare = {k: v for k, v in w.items() if not k.startswith("link") and k != "n_links"}
    assert abs(O.reduced_system(bare, lam)["chi2"] - ref["chi2"]) > 1e-3 * ref["chi2"], "the links do not contribute: test is void"


def test_solve_with_odometry_links(ba):
    w = synth.make_window(7, 400, layout="all", seed=403, links="chain")
    check_solution(ba.solve(w), O.solve(w), "chain of links")
    # a pose that only the links hold (it loses all its visual edges), and no fixed pose at all
    w = synth.make_window(6, 250, layout="all", seed=404, links="chain", root=None)
    keep = w["edge_pose"] != 2
    for k in ("edge_obs", "edge_pose", "edge_point", "edge_kind"):
        w[k] = np.ascontiguousarray(w[k][keep])
    w["n_edges"] = int(keep.sum())
    check_solution(ba.solve(w), O.solve(w), "pose held by links only")


def test_solve_links_with_pcg_and_in_a_batch(ba):
    a = synth.make_window(6, 300, layout="all", seed=405, links="chain", solver=2)
    check_solution(ba.solve(a), O.solve(a), "links + PCG")
    ws = [synth.make_window(5 + k % 3, 160 + 10 * k, layout="all", seed=410 + k, links=("chain" if k % 2 == 0 else None)) for k in range(6)]
    got = ba.solve_batch(ws)
    for k, (g, w) in enumerate(zip(got, ws)):
        check_solution(g, O.solve(w), f"batch window {k}")


# ---------------------------------------------------------------- empty inputs
def test_empty_and_degenerate_windows(ba):
    # no edges at all: nothing is active, both passes stop as "empty", the state comes back untouched
    w = synth.make_window(4, 30, layout="all", seed=501)
    for k in ("edge_obs", "edge_pose", "edge_point", "edge_kind"):
        w[k] = np.ascontiguousarray(w[k][:0])
    w["n_edges"] = 0
    got, ref = ba.solve(w), O.solve(w)
    check_solution(got, ref, "no edges")
    assert np.array_equal(got["pose_tq"], w["pose_tq"]) and np.array_equal(got["point_xyz"], w["point_xyz"])
    # every vertex fixed: edges exist but none is active
    w = synth.make_window(3, 20, layout="all", seed=502, fixed_point_frac=1.0)
    w["pose_fixed"][:] = 1
    check_solution(ba.solve(w), O.solve(w), "everything fixed")
    # a batch that mixes an empty window with ordinary ones
    ws = [synth.make_window(5, 100, layout="all", seed=503), dict(w), synth.make_window(4, 80, layout="all", seed=504)]
    for g, x in zip(ba.solve_batch(ws), ws):
        check_solution(g, O.solve(x), "mixed batch")


# ---------------------------------------------------------------- seeded sweep over window shapes and options
def _sweep_windows():
    rng = np.random.default_rng(20261018)
    ws = []
    for k in range(36):
        P = int(rng.integers(2, 33))                      # crosses the warp-specialised (<= 20 poses) / fallback build boundary
        layout = ("all", "consecutive", "random")[k % 3] if P <= 20 else ("consecutive", "random")[k % 2]
        views = int(rng.integers(2, min(P, 12) + 1))
        L = int(rng.integers(40, 400))
        kw = dict(layout=layout, views=views, seed=5000 + k, mono_frac=float(rng.choice([0.0, 0.0, 0.25, 1.0])),
                  fixed_point_frac=float(rng.choice([0.0, 0.0, 0.15])), outlier_frac=float(rng.choice([0.0, 0.05, 0.15])),
                  root=rng.choice(["second_newest", "first", None]), shuffle_edges=bool(rng.integers(0, 4) == 0),
                  iterations=int(rng.choice([10, 10, 6, 7])), solver=int(rng.choice([0, 0, 0, 2])),
                  trust_region=int(rng.choice([0, 0, 0, 1])), huber_delta=float(rng.choice([8.0, 8.0, 3.0])),
                  links="chain" if k % 4 == 1 else None)
        if kw["trust_region"] == 1:                        # undamped Gauss-Newton: start inside its basin
            kw.update(outlier_frac=0.0, pose_noise=(0.004, np.deg2rad(0.1)), point_noise=0.01)
        if kw["root"] is None or kw["mono_frac"] == 1.0:   # gauge / scale only held by damping: keep LM
            kw.update(trust_region=0)
        ws.append(synth.make_window(P, L, **kw))
    return ws


def test_seeded_sweep_of_window_shapes_and_options(ba):
    ws = _sweep_windows()
    small = [w for w in ws if w["n_poses"] <= 32]
    got = ba.solve_batch(small)                            # one heterogeneous batch (pipelined: >= 32 windows)
    bad = []
    for k, (w, g) in enumerate(zip(small, got)):
        r = O.solve(w)
        try:
            check_solution(g, r, f"sweep window {k}")
        except AssertionError as e:
            bad.append(str(e)[:300])
    assert not bad, "\n".join(bad)


# ---------------------------------------------------------------- the opt-in tensor-pipe build kernel (k_build_ds)
def test_tensor_pipe_build_variant_matches_oracle(ba, monkeypatch):
    """VISFS_BA_USE_DS=1 routes windows of <= 10 poses through k_build_ds (Schur products as DMMA); same gates."""
    monkeypatch.setenv("VISFS_BA_USE_DS", "1")
    cases = {
        "C1": synth.config_c1(),
        "C1 30% fixed points": synth.config_c1(fixed_point_frac=0.3),
        "consecutive views": synth.make_window(8, 500, layout="consecutive", views=3, seed=77),
        "rejected steps": rejecting_window(102, (0.3, np.deg2rad(6.0))),
        "chain of links": synth.make_window(7, 400, layout="all", seed=403, links="chain"),
        "no fixed pose": synth.make_window(5, 200, layout="all", seed=12, root=None),
    }
    for what, w in cases.items():
        check_solution(ba.solve(w), O.solve(w), "ds " + what)
    ws = synth.config_c3_windows(24)
    got = ba.solve_batch(ws)
    monkeypatch.delenv("VISFS_BA_USE_DS")
    ref = ba.solve_batch(ws)
    for g, r in zip(got, ref):
        check_solution(g, r, "ds batch vs default batch")

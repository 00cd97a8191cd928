"""GPU parity tests of the large-window path (block-skyline reduced system, ba_large.cuh) and of the landmark
partition that the multi-GPU global BA uses: CUDA through the C ABI against the CPU oracle, same gates as
tests/test_gpu_parity.py."""
import os
import subprocess
import sys

import numpy as np
import pytest

from tests import oracle_api as O
from tests.test_gpu_parity import check_solution, rel_close
from visfs_b200 import capi, partition, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def banded(seed=71, P=40, L=1500, **kw):
    return synth.make_window(P, L, views=6, layout="consecutive", trajectory="line", seed=seed, **kw)


def loop(seed=72, P=48, L=2000, **kw):
    return synth.make_window(P, L, views=7, layout="consecutive", trajectory="loop", seed=seed, **kw)


def dense(seed=73, P=36, L=1200, **kw):
    return synth.make_window(P, L, views=8, layout="random", trajectory="orbit", seed=seed, **kw)


def _structure_case(ba, w, level=None):
    cap = w["n_poses"] * (w["n_poses"] + 1) // 2
    got, ref = ba.structure(w, level, cap), O.structure(w, level, cap)
    for k in ("pose_hidx", "point_hidx", "edge_active", "hpl_row", "hpl_col", "schur_rows", "schur_cols"):
        assert np.array_equal(got[k], ref[k]), f"{k} differs"
    for k in ("n_schur_blocks", "n_free_poses", "n_free_points", "n_active_edges", "n_hpl_blocks"):
        assert got[k] == ref[k], (k, got[k], ref[k])


@pytest.mark.parametrize("make", [banded, loop, dense])
def test_large_structure_bit_exact(ba, make):
    _structure_case(ba, make(fixed_point_frac=0.2))


def test_large_structure_with_culled_edges(ba):
    w = loop()
    rng = np.random.default_rng(9)
    level = (rng.random(w["n_edges"]) < 0.3).astype(np.uint8)
    level[w["edge_pose"] == 5] = 1
    _structure_case(ba, w, level)


@pytest.mark.parametrize("make", [banded, loop, dense])
def test_large_reduced_system_matches_oracle(ba, make):
    w = make(mono_frac=0.2, fixed_point_frac=0.1)
    lam = 2.5
    got, ref = ba.debug_trial(w, lam), O.reduced_system(w, lam)
    assert got["n"] == ref["n"]
    rel_close(got["chi2"], ref["chi2"], 1e-12, "chi2 at the input state")
    rel_close(got["S"], ref["S"], 1e-10, "reduced camera system")
    rel_close(got["b_s"], ref["b_s"], 1e-10, "reduced rhs")
    rel_close(got["x_pose"], ref["x"][: ref["n"]], 1e-7, "pose step")
    lam0 = ba.debug_trial(w, -1.0)["lambda_used"]
    rel_close(lam0, ref["lambda_init"], 1e-12, "initial damping")


@pytest.mark.parametrize("make", [banded, loop, dense])
def test_large_solve_matches_oracle(ba, make):
    w = make()
    check_solution(ba.solve(w), O.solve(w), make.__name__)


def test_large_solve_with_fixed_points_mono_and_no_fixed_pose(ba):
    w = loop(seed=75, mono_frac=0.3, fixed_point_frac=0.25, root=None)
    check_solution(ba.solve(w), O.solve(w), "loop, gauge free")


def test_large_gauss_newton_and_single_pass(ba):
    # Gauss-Newton has no damping: use a start inside its convergence basin, otherwise rounding decides the outcome
    w = banded(seed=77, trust_region=1, outlier_frac=0.0, pose_noise=(0.005, np.deg2rad(0.1)), point_noise=0.01)
    check_solution(ba.solve(w), O.solve(w), "GN")
    w = banded(seed=83, huber_delta=0.0)
    check_solution(ba.solve(w), O.solve(w), "no kernel")


def test_large_solve_pcg(ba):
    # Optimizer/Solver = 2 on the block skyline: g2o's block-Jacobi PCG, including its tolerance quirk
    w = loop(seed=85, P=40, L=1200, solver=2, iterations=6)
    check_solution(ba.solve(w), O.solve(w), "PCG, loop")
    w = dense(seed=86, P=34, L=600, solver=2, iterations=4)
    check_solution(ba.solve(w), O.solve(w), "PCG, dense")


def test_large_with_odometry_links(ba):
    w = loop(seed=87, P=40, L=900, links="chain")
    _structure_case(ba, w)
    lam = 0.9
    got, ref = ba.debug_trial(w, lam), O.reduced_system(w, lam)
    rel_close(got["chi2"], ref["chi2"], 1e-12, "chi2 with links")
    rel_close(got["S"], ref["S"], 1e-10, "reduced camera system with pose-pose blocks")
    rel_close(got["b_s"], ref["b_s"], 1e-10, "reduced rhs")
    rel_close(ba.debug_trial(w, -1.0)["lambda_used"], ref["lambda_init"], 1e-12, "initial damping")
    check_solution(ba.solve(w), O.solve(w), "large window with links")
    part = partition.partition_window(w, 1, 0)
    check_solution(partition.merge_results(w, [part], [ba.solve(part)]), O.solve(w), "1-rank partition with links")


def test_multifrontal_band_solver_matches_oracle(ba, monkeypatch):
    # ba_mf.cuh: nested dissection of the key-frame chain, dense fronts, DMMA corner updates.  It is chosen from 256 free poses
    # up; VISFS_BA_MF_MIN (read at the start of every pass) brings it down to windows the oracle checks in a second.
    monkeypatch.setenv("VISFS_BA_MF_MIN", "8")
    cases = [("loop with closure rows", loop(seed=171, P=96, L=2500)),
             ("plain band", banded(seed=172, P=150, L=3000)),
             ("narrow band, gauge free, mono + fixed points", synth.make_window(70, 1500, views=3, layout="consecutive", seed=173, mono_frac=0.3,
                                                                               fixed_point_frac=0.2, root=None)),
             ("10 views like C4", synth.make_window(260, 6000, views=10, layout="consecutive", trajectory="loop", seed=174)),
             ("band with odometry links", loop(seed=175, P=64, L=1500, links="chain"))]
    for name, w in cases:
        lam = 0.7
        got, ref = ba.debug_trial(w, lam), O.reduced_system(w, lam)
        rel_close(got["x_pose"], ref["x"][: ref["n"]], 1e-7, name + ": pose step of one damped trial")
        check_solution(ba.solve(w), O.solve(w), name)
    # the same windows through the one-CTA frontal solver give the same answer: both are exercised
    monkeypatch.setenv("VISFS_BA_NO_MF", "1")
    check_solution(ba.solve(cases[0][1]), O.solve(cases[0][1]), "frontal solver")


def test_band_chunks_build_matches_oracle(ba, monkeypatch):
    # ba_band.cuh: the large build through ws::k_build_band + k_band_gather (no atomics).  Chosen for maps of >= 4096
    # landmarks with band structure; VISFS_BA_BAND_FORCE brings it to windows the oracle checks in a second, including maps where
    # only some chunks qualify (loop closure, random views) and lg::k_build_large_run takes the rest.
    monkeypatch.setenv("VISFS_BA_BAND_FORCE", "1")
    cases = [("plain band", banded(seed=181, P=80, L=2500, mono_frac=0.2, fixed_point_frac=0.1)),
             ("loop: closure landmarks outside the chunks", loop(seed=182, P=64, L=2500)),
             ("random views: most chunks do not qualify", dense(seed=183)),
             ("10 views like C4", synth.make_window(260, 6000, views=10, layout="consecutive", trajectory="loop", seed=184)),
             ("gauge free, mono + fixed points", synth.make_window(70, 1500, views=3, layout="consecutive", seed=185, mono_frac=0.3,
                                                                   fixed_point_frac=0.2, root=None)),
             ("band with odometry links", loop(seed=186, P=48, L=1200, links="chain"))]
    for name, w in cases:
        lam = 1.3
        got, ref = ba.debug_trial(w, lam), O.reduced_system(w, lam)
        rel_close(got["chi2"], ref["chi2"], 1e-12, name + ": chi2")
        rel_close(got["S"], ref["S"], 1e-10, name + ": reduced camera system")
        rel_close(got["b_s"], ref["b_s"], 1e-10, name + ": reduced rhs")
        check_solution(ba.solve(w), O.solve(w), name)
    part = partition.partition_window(cases[3][1], 1, 0)
    check_solution(partition.merge_results(cases[3][1], [part], [ba.solve(part)]), O.solve(cases[3][1]), "1-rank partition, band chunks")


def test_large_rejected_steps(ba):
    w = synth.make_window(34, 400, views=6, layout="consecutive", seed=78, pose_noise=(0.3, np.deg2rad(6.0)), point_noise=0.5,
                          iterations=20, depth_range=(1.0, 6.0))
    ref = O.solve_timed(w)
    assert max(ref["trials_per_iteration"]) > 1, "window does not reject any step: pick another seed"
    check_solution(ba.solve(w), ref, "rejecting large window")


def test_large_degree_limit_is_reported(ba):
    w = synth.make_window(40, 30, layout="all", seed=79)    # every landmark seen by 40 poses > 32
    with pytest.raises(capi.BAError):
        ba.solve(w)


def test_small_window_through_large_path(ba, monkeypatch):
    # the same window through both code paths (VISFS_BA_FORCE_LARGE is read at upload)
    w = synth.make_window(8, 500, layout="all", seed=80)
    small = ba.solve(w)
    monkeypatch.setenv("VISFS_BA_FORCE_LARGE", "1")
    big = ba.solve(w)
    # ... and through the band chunks of the large path: every landmark starts at the same frame, so the whole map is ONE chunk
    # and every skyline block has exactly one source in the gather list
    monkeypatch.setenv("VISFS_BA_BAND_FORCE", "1")
    band = ba.solve(w)
    got_sys, ref_sys = ba.debug_trial(w, 0.8), O.reduced_system(w, 0.8)
    monkeypatch.delenv("VISFS_BA_BAND_FORCE")
    monkeypatch.delenv("VISFS_BA_FORCE_LARGE")
    ref = O.solve(w)
    check_solution(small, ref, "small path")
    check_solution(big, ref, "large path")
    check_solution(band, ref, "large path, one band chunk")
    rel_close(got_sys["S"], ref_sys["S"], 1e-10, "one band chunk: reduced camera system")
    rel_close(got_sys["b_s"], ref_sys["b_s"], 1e-10, "one band chunk: reduced rhs")


def test_partitioned_single_rank_equals_unpartitioned(ba):
    w = loop(seed=81)
    part = partition.partition_window(w, 1, 0)
    got = partition.merge_results(w, [part], [ba.solve(part)])
    check_solution(got, O.solve(w), "1-rank partition")


def test_c5_reduced_size_properties(ba):
    w = synth.config_c5(n_poses=120, n_points=20000)
    g = ba.solve(w)
    assert g["status"] == 0 and g["chi2_final"] < g["chi2_pass1"] < g["chi2_initial"]
    assert np.isfinite(g["pose_tq"]).all() and np.isfinite(g["point_xyz"]).all()
    fixed = w["pose_fixed"].astype(bool)
    assert np.array_equal(g["pose_tq"][fixed, :3], w["pose_tq"][fixed, :3])       # the gauge pose never moves
    assert np.allclose(g["pose_tq"][fixed, 3:], w["pose_tq"][fixed, 3:], rtol=0, atol=5e-16)   # (re-normalised like CameraPose(q, t): 2 ulp)


@pytest.mark.skipif("torch" not in sys.modules and False, reason="")
def test_global_ba_two_ranks_nccl():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "global_ba_ranks.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "GLOBAL_BA_OK" in r.stdout


# ---------------------------------------------------------------- seeded sweep over large-window shapes and options
def _large_sweep():
    rng = np.random.default_rng(20261019)
    ws = []
    for k in range(14):
        P = int(rng.integers(33, 140))
        shape = k % 3
        kw = dict(seed=6000 + k, views=int(rng.integers(3, 12)), mono_frac=float(rng.choice([0.0, 0.0, 0.3])),
                  fixed_point_frac=float(rng.choice([0.0, 0.1])), outlier_frac=float(rng.choice([0.0, 0.05])),
                  root=rng.choice(["second_newest", "first"]), shuffle_edges=bool(rng.integers(0, 4) == 0),
                  iterations=int(rng.choice([10, 6])), solver=int(rng.choice([0, 0, 2])), links="chain" if k % 4 == 2 else None)
        kw.update([dict(layout="consecutive", trajectory="line"), dict(layout="consecutive", trajectory="loop"),
                   dict(layout="random", trajectory="orbit")][shape])
        if shape == 1 and P > 90:
            P = 90      # (longer loops leave the origin far enough for the reference's Jacobian mismatch to make LM chaotic)
        ws.append(synth.make_window(P, int(rng.integers(300, 2500)), **kw))
    return ws


def test_seeded_sweep_of_large_windows(ba):
    bad = []
    for k, w in enumerate(_large_sweep()):
        g, r = ba.solve(w), O.solve(w)
        try:
            if w["solver"] == 2:
                # g2o's PCG stops its first solve at a relative residual of 1e-6, so two correct implementations agree to
                # about that in the step: same decisions and chi2, state to 1e-5
                for key in ("status", "iterations_run", "trials_run", "stop_reason", "n_outliers"):
                    assert g[key] == r[key], (key, g[key], r[key])
                assert np.array_equal(g["edge_level"], r["edge_level"])
                assert abs(g["chi2_final"] - r["chi2_final"]) <= 1e-6 * r["chi2_final"]
                np.testing.assert_allclose(g["pose_tq"], r["pose_tq"], rtol=1e-5, atol=1e-5)
                # (a weakly observed landmark amplifies the PCG residual: observed up to 9e-5 on 2 of 4 500 coordinates, run to run,
                # because the large build adds with floating-point atomics and the CG iterates feel the summation order)
                np.testing.assert_allclose(g["point_xyz"], r["point_xyz"], rtol=5e-4, atol=5e-4)
            else:
                check_solution(g, r, f"large sweep window {k} (P={w['n_poses']})")
        except AssertionError as e:
            bad.append(f"window {k} (P={w['n_poses']}, solver {w['solver']}): " + str(e)[:300])
    assert not bad, "\n".join(bad)


# ---------------------------------------------------------------- BASELINE configs C4 / C5 at full size
# The multi-threaded oracle finishes these in well under a minute, so the full sizes are checked against it directly
# (decisions exact, state to rounding) on top of the size-independent properties.
#
# C4 note: the reference pairs an exponential-map Jacobian (OptimizeTypeDefine.h:157-176, rotation columns built from the
# camera-frame point) with a decoupled update (CameraPose::update, OptimizeTypeDefine.cpp:7-14: t += dt, q = dq * q).  The
# two differ by [t_cw]x, so far from the world origin (the 48 m loop of C4) the LM gain ratio turns negative, lambda climbs
# and pass 1 hardly moves; most edges then sit above chi2 > delta and are culled.  That is the reference's behaviour, the
# oracle restates it and the CUDA path reproduces it decision for decision — hence no "few outliers" assertion here.
def _check_properties(w, g):
    assert g["status"] == 0 and g["chi2_final"] < g["chi2_initial"]
    assert np.isfinite(g["pose_tq"]).all() and np.isfinite(g["point_xyz"]).all()
    assert np.allclose(np.linalg.norm(g["pose_tq"][:, 3:7], axis=1), 1.0, atol=1e-12)
    fixed = w["pose_fixed"].astype(bool)
    assert np.allclose(g["pose_tq"][fixed], w["pose_tq"][fixed], rtol=0, atol=5e-16 * max(1.0, np.abs(w["pose_tq"]).max()))   # the gauge pose never moves (re-normalised like CameraPose(q, t): 2 ulp)
    assert g["n_outliers"] == int(g["edge_level"].sum())
    assert g["n_outliers"] >= 0.04 * w["n_edges"]                            # the planted +-20 px outliers are among the culled


def _check_against_oracle(w, g):
    o = O.solve(w, threads=O.threads(), omp=True)
    assert g["status"] == o["status"] == 0
    for k in (0, 1):
        # once lambda has run away (C4 note above) the steps are below 1e-12 of the state and the sign of the gain ratio is
        # rounding noise: whether such a pass ends on its 9th or 10th null trial is not a decision worth comparing, the
        # accepted state below is
        if max(g["lambda_final"][k], o["lambda_final"][k]) > 1e12:
            continue
        assert g["iterations_run"][k] == o["iterations_run"][k] and g["trials_run"][k] == o["trials_run"][k]
        assert g["stop_reason"][k] == o["stop_reason"][k]
    assert g["n_outliers"] == o["n_outliers"] and np.array_equal(g["edge_level"], o["edge_level"])
    for k in ("chi2_initial", "chi2_pass1", "chi2_final"):
        assert abs(g[k] - o[k]) <= 1e-6 * abs(o[k]), k
    np.testing.assert_allclose(g["pose_tq"], o["pose_tq"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(g["point_xyz"], o["point_xyz"], rtol=1e-6, atol=1e-6)


def test_c4_full_size_oracle_parity_and_partition_invariance(ba):
    w = synth.config_c4()                      # 2 000 key frames on a loop / 500 000 landmarks / 5 M edges
    g = ba.solve(w)
    _check_properties(w, g)
    _check_against_oracle(w, g)
    # the same problem handed over as a 1-rank landmark partition takes the same code path: same decisions, same answer
    part = partition.partition_window(w, 1, 0)
    p = partition.merge_results(w, [part], [ba.solve(part)])
    assert p["iterations_run"] == g["iterations_run"] and p["trials_run"] == g["trials_run"]
    assert np.array_equal(p["edge_level"], g["edge_level"])
    assert np.allclose(p["pose_tq"], g["pose_tq"], rtol=1e-7, atol=1e-9)     # (red.global.add: summation order is not fixed)
    assert abs(p["chi2_final"] - g["chi2_final"]) <= 1e-7 * g["chi2_final"]


def test_c5_full_size_oracle_parity(ba):
    w = synth.config_c5()                      # 200 key frames / 200 000 landmarks / 2 M edges, dense reduced system
    g = ba.solve(w)
    _check_properties(w, g)
    _check_against_oracle(w, g)

"""Stage-by-stage GPU-vs-oracle comparison with verbose output (run on the GPU box when a parity test
fails: `python -m tests.gpu_debug`)."""
import sys
import numpy as np
from tests import oracle_api as O
from visfs_b200 import capi, synth


def d(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b))) / max(float(np.max(np.abs(b))), 1e-300) if b.size else 0.0


def main():
    ba = capi.BundleAdjuster(0)
    print("fp64 probe TFLOP/s:", ba.probe_fp64())
    w = synth.make_window(6, 300, layout="all", seed=41)
    g, r = ba.linearize(w), O.linearize(w)
    print("linearize", {k: d(g[k], r[k]) for k in g})
    gs, rs = ba.structure(w), O.structure(w)
    print("structure", {k: bool(np.array_equal(gs[k], rs[k])) if hasattr(gs[k], "shape") else (gs[k], rs[k]) for k in gs})
    for lam in (3.7, -1.0):
        gt, rt = ba.debug_trial(w, lam), O.reduced_system(w, lam if lam >= 0 else 0.0)
        print(f"trial lam={lam}: n", gt["n"], rt["n"], "chi2", gt["chi2"], rt["chi2"], "lambda", gt["lambda_used"], rt["lambda_init"])
        if lam >= 0:
            print("  S", d(gt["S"], rt["S"]), "bs", d(gt["b_s"], rt["b_s"]), "xp", d(gt["x_pose"], rt["x"][: rt["n"]]),
                  "trial chi2", gt["trial_chi2"])
            if d(gt["S"], rt["S"]) > 1e-9:
                np.set_printoptions(linewidth=200, precision=4)
                print("  S gpu [0:12,0:12]\n", gt["S"][:12, :12], "\n  S ref\n", rt["S"][:12, :12])
                print("  bs gpu", gt["b_s"][:12], "\n  bs ref", rt["b_s"][:12])
            pts = w["point_xyz"] + rt["x"][rt["n"]:].reshape(-1, 3)[: w["n_points"]]
            print("  trial points", d(gt["trial_points"], pts))
    got, ref = ba.solve(w), O.solve(w)
    for k in ("status", "n_outliers", "iterations_run", "trials_run", "stop_reason", "n_free_poses", "n_free_points",
              "chi2_initial", "chi2_pass1", "chi2_final", "chi2_last_trial", "lambda_final"):
        print(f"  {k:16s} gpu {got[k]}   ref {ref[k]}")
    print("  poses", d(got["pose_tq"], ref["pose_tq"]), "points", d(got["point_xyz"], ref["point_xyz"]),
          "levels equal", bool(np.array_equal(got["edge_level"], ref["edge_level"])))
    import os
    ba2 = capi.BundleAdjuster(0, profile_kernels=True)
    keys = ("total_ms", "build_ms", "solve_ms", "update_ms", "other_ms", "lm_trials", "kernel_launches", "solve_clocks")
    for mode in ("ws", "ws-nocluster"):
        os.environ.pop("VISFS_BA_NO_WS", None); os.environ.pop("VISFS_BA_NO_CLUSTER", None)
        if mode == "v1":
            os.environ["VISFS_BA_NO_WS"] = "1"
        if mode == "ws-nocluster":
            os.environ["VISFS_BA_NO_CLUSTER"] = "1"
        for name, ws in (("C1", [synth.config_c1()]), ("C2", [synth.config_c2()]), ("C3x64", synth.config_c3_windows(64)),
                         ("C3x256", synth.config_c3_windows(256))):
            ba2.upload(ws); ba2.run_resident(); ba2.run_resident(); ba2.run_resident()
            t = ba2.timing()
            print(f"{mode:13s} {name:7s}", {k: (round(t[k], 3) if isinstance(t[k], float) else t[k]) for k in keys}, flush=True)
    os.environ.pop("VISFS_BA_NO_WS", None); os.environ.pop("VISFS_BA_NO_CLUSTER", None)


if __name__ == "__main__":
    sys.exit(main())

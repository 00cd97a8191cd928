"""Timing of the large-window configurations (C4 global BA, C5 dense window) on one GPU: python tools/time_large.py [scale]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from visfs_b200 import capi, synth  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
ba = capi.BundleAdjuster(0, profile_kernels=True)
keys = ("solve_clocks", "total_ms", "build_ms", "solve_ms", "update_ms", "other_ms", "lm_iterations", "lm_trials", "kernel_launches")
for name, make in (("C5", lambda: synth.config_c5(n_poses=200, n_points=int(200000 * scale))),
                   ("C4", lambda: synth.config_c4(n_poses=int(2000 * scale), n_points=int(500000 * scale)))):
    t0 = time.time()
    w = make()
    t1 = time.time()
    ba.upload([w])
    for _ in range(3):
        ba.run_resident()
    t = ba.timing()
    r = ba.download()[0]
    print(name, f"gen {t1 - t0:.1f}s  P {w['n_poses']} L {w['n_points']} E {w['n_edges']}",
          {k: (round(t[k], 3) if isinstance(t[k], float) else t[k]) for k in keys},
          "status", r["status"], "chi2", r["chi2_initial"], r["chi2_pass1"], r["chi2_final"], "iters", r["iterations_run"], r["trials_run"],
          flush=True)

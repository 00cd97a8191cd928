"""Small windows through every hot kernel, for compute-sanitizer (tools/sanitize.sh -> profiles/r2_sanitizer_*.txt):
k_build_ws with clusters (single window) and without (batch), k_update, k_solve, the block-skyline path with
k_solve_front (loop trajectory) and with the dense solver (random views), odometry links."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from visfs_b200 import capi, synth  # noqa: E402

ba = capi.BundleAdjuster(0)
cases = [
    ("single 6x300 (k_build_ws in clusters of 4, k_solve, k_update)", lambda: ba.solve(synth.make_window(6, 300, layout="all", seed=7))),
    ("batch 40 x 5x120 (k_build_ws without clusters)", lambda: ba.solve_batch([synth.make_window(5, 120, layout="all", seed=20 + k) for k in range(40)])),
    ("single 8x200 with odometry links", lambda: ba.solve(synth.make_window(8, 200, layout="all", seed=8, links="chain"))),
    ("loop 40x600 (block skyline, k_solve_front)", lambda: ba.solve(synth.make_window(40, 600, views=6, layout="consecutive", trajectory="loop", seed=9))),
    ("dense 40x800 random views (block skyline, dense solver)", lambda: ba.solve(synth.make_window(40, 800, views=8, layout="random", trajectory="orbit", seed=10))),
]
for name, fn in cases:
    r = fn()
    r = r if isinstance(r, dict) else r[0]
    print(f"{name}: status {r['status']} chi2 {r['chi2_initial']:.3f} -> {r['chi2_final']:.3f}", flush=True)
ba.close()

"""Resident C4 (global BA) for ncu launch lists of the large-path kernels: python tools/profile_c4.py [scale]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from visfs_b200 import capi, synth  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
ba = capi.BundleAdjuster(0, profile_kernels=True)
ba.upload([synth.config_c4(n_poses=int(2000 * scale), n_points=int(500000 * scale))])
ba.run_resident()
ba.run_resident()
t = ba.timing()
print({k: t[k] for k in ("total_ms", "build_ms", "solve_ms", "update_ms", "other_ms", "lm_trials", "kernel_launches")})

"""Resident C5 window (dense reduced system) for ncu / phase timing of the dense DMMA Cholesky: python tools/profile_c5.py [points]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from visfs_b200 import capi, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
ba = capi.BundleAdjuster(0, profile_kernels=True)
ba.upload([synth.config_c5(n_poses=200, n_points=n)])
ba.run_resident()
ba.run_resident()
t = ba.timing()
print({k: t[k] for k in ("total_ms", "build_ms", "solve_ms", "update_ms", "other_ms", "lm_trials", "kernel_launches")})

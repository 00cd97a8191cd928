"""One resident C1 window (10 key frames / 2 000 landmarks / 20 000 edges) for an ncu launch list: python tools/profile_c1.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from visfs_b200 import capi, synth  # noqa: E402

ba = capi.BundleAdjuster(0, profile_kernels=False)
ba.upload([synth.config_c1()])
for _ in range(3):
    ba.run_resident()
t = ba.timing()
print({k: t[k] for k in ("total_ms", "lm_trials", "kernel_launches", "solve_clocks")})

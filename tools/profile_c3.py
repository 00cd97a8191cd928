"""Resident C3 batch for an ncu capture of the small-window kernels (k_build_ws / k_update / k_solve)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from visfs_b200 import capi, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
ba = capi.BundleAdjuster(0)
ba.upload(synth.config_c3_windows(n))
ba.run_resident()
ba.run_resident()
print(ba.timing()["total_ms"])

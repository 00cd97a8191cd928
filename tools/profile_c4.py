"""One C4 solve (global BA, one rank) for ncu: python tools/profile_c4.py [scale]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from visfs_b200 import capi, synth  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
ba = capi.BundleAdjuster(0)
w = synth.config_c4(n_poses=int(2000 * scale), n_points=int(500000 * scale))
ba.upload([w])
ba.run_resident()
print(ba.timing()["total_ms"])

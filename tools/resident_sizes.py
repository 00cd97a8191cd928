"""bench.py's resident-window legs alone (c0 and c2 sized maps): python tools/resident_sizes.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from visfs_b200 import capi  # noqa: E402

ba = capi.BundleAdjuster(0)
for size in ("c0", "c2"):
    print(json.dumps(bench.resident_window_numbers(ba, False, size=size))[:700])

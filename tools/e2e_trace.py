"""Per-group timeline of one pipelined visfs_ba_solve_batch call (VISFS_BA_TRACE=1 prints it to stderr)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from visfs_b200 import capi, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
pinned = len(sys.argv) > 2 and sys.argv[2] == "pinned"
ws = synth.config_c3_windows(n)
ba = capi.BundleAdjuster(0)
packed = ba.prepare_batch(ws, pinned=pinned)
for _ in range(3):
    ba.solve_packed(packed)
os.environ["VISFS_BA_TRACE"] = "1"
ba.solve_packed(packed)

"""Pipelined batch with page-locked inputs: staging route vs direct DMA inside the groups (VISFS_BA_DIRECT_GROUPS), for a
host with few cores per GPU (run under `taskset -c 0,1` with LOCAL_WORLD_SIZE=8 to mimic 8 ranks on 16 cores)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from visfs_b200 import capi, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
ws = synth.config_c3_windows(n)
ba = capi.BundleAdjuster(0)
packed = ba.prepare_batch(ws, pinned=True, float_obs=True)
for direct in (False, True):
    if direct:
        os.environ["VISFS_BA_DIRECT_GROUPS"] = "1"
    else:
        os.environ.pop("VISFS_BA_DIRECT_GROUPS", None)
    for g in [int(x) for x in os.environ.get("GROUP_LIST", "2,4,8").split(",")]:
        os.environ["VISFS_BA_GROUPS"] = str(g)
        for _ in range(2):
            ba.solve_packed(packed)
        t0 = time.perf_counter()
        for _ in range(3):
            ba.solve_packed(packed)
        dt = (time.perf_counter() - t0) / 3
        print(f"direct {direct!s:5} groups {g:2d}: {1e3 * dt:7.2f} ms per batch of {n} windows", flush=True)

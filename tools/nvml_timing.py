"""How fast can NVML be polled from a Python thread while the library runs resident solves (bench.py's ClockSampler)?"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from visfs_b200 import capi, synth  # noqa: E402

ba = capi.BundleAdjuster(0, profile_kernels=len(sys.argv) > 2)
ba.upload(synth.config_c3_windows(int(sys.argv[1]) if len(sys.argv) > 1 else 256))
ba.run_resident()
c = bench.ClockSampler(0)
with c:
    t0 = time.perf_counter()
    for _ in range(3):
        ba.run_resident()
        t = ba.timing()
    dt = time.perf_counter() - t0
print("3 resident solves %.1f ms" % (1e3 * dt), c.summary())
c = bench.ClockSampler(0)
with c:
    time.sleep(0.1)
print("100 ms idle", c.summary())

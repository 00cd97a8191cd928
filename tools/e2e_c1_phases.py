"""Where the end-to-end time of ONE C1 window goes: upload / resident LM run / download, host wall clock.  python tools/e2e_c1_phases.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from visfs_b200 import capi, synth  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "c1"
w = {"c1": synth.config_c1, "c2": synth.config_c2}[which]()
ba = capi.BundleAdjuster(0)
packed = ba.prepare_batch([w], pinned=True, float_obs=True)
for _ in range(20):
    ba.solve_packed(packed)
t0 = time.perf_counter()
N = 200
for _ in range(N):
    ba.solve_packed(packed)
t_all = (time.perf_counter() - t0) / N
up = run = down = 0.0
for _ in range(N):
    a = time.perf_counter(); ba.upload([w]); b = time.perf_counter(); ba.run_resident(); c = time.perf_counter(); ba.download(); d = time.perf_counter()
    up += b - a; run += c - b; down += d - c
print({"solve_packed_ms": t_all * 1e3, "upload_ms (python marshalling included)": up / N * 1e3, "run_resident_ms": run / N * 1e3, "download_ms": down / N * 1e3,
       "device_ms": ba.timing()["total_ms"]})

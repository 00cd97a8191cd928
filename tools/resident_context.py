import json, os, sys, time
sys.path.insert(0, "/root/repo")
import bench
from visfs_b200 import capi, synth
ba = capi.BundleAdjuster(0)
ws = synth.config_c3_windows(512)
packed = ba.prepare_batch(ws, pinned=True, float_obs=True)
ba.solve_packed(packed); ba.solve_packed(packed)
r = bench.resident_window_numbers(ba, False, size="c2")
print("same handle after a pipelined batch:", round(r["per_frame_ms_resident"], 3), round(r["per_frame_ms_full_call"], 3))
ba2 = capi.BundleAdjuster(0)
r = bench.resident_window_numbers(ba2, False, size="c2")
print("fresh handle, same process:", round(r["per_frame_ms_resident"], 3), round(r["per_frame_ms_full_call"], 3))
del packed
r = bench.resident_window_numbers(ba2, False, size="c2")
print("fresh handle, batch buffers released:", round(r["per_frame_ms_resident"], 3), round(r["per_frame_ms_full_call"], 3))

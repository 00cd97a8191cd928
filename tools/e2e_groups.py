"""End-to-end time of visfs_ba_solve_batch (host buffers) for different pipeline group counts and size ramps: VISFS_BA_GROUPS / VISFS_BA_RAMP are read per call."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from visfs_b200 import capi, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
ws = synth.config_c3_windows(n)
ba = capi.BundleAdjuster(0)
packed = ba.prepare_batch(ws, pinned=True, float_obs=True)
for rep in range(2):
    sweep = ((1.22, 99, 16), (1.22, 12, 16), (1.22, 10, 20), (1.22, 10, 24), (1.3, 8, 20), (1.15, 99, 16), (1.22, 99, 20))
    if len(sys.argv) > 2 and sys.argv[2] == "big":   # large batches: more, flatter groups
        sweep = ((1.22, 10, 20), (1.22, 10, 16), (1.22, 12, 24), (1.3, 10, 20), (1.22, 8, 20), (1.22, 10, 28), (1.18, 12, 20))
    if len(sys.argv) > 3:   # explicit group counts: default ramp, plus 0 = the library's own choice
        sweep = tuple((1.22, max(4, int(x) // 2), int(x)) for x in sys.argv[3].split(','))
    for ramp, flat, g in sweep:
        os.environ["VISFS_BA_RAMP"] = str(ramp)
        os.environ["VISFS_BA_RAMP_FLAT"] = str(flat)
        if g > 0:
            os.environ["VISFS_BA_GROUPS"] = str(g)
            os.environ["VISFS_BA_RAMP_FLAT"] = str(flat)
        else:
            os.environ.pop("VISFS_BA_GROUPS", None); os.environ.pop("VISFS_BA_RAMP_FLAT", None)
        for _ in range(2):
            ba.solve_packed(packed)
        t0 = time.perf_counter()
        for _ in range(6):
            ba.solve_packed(packed)
        dt = (time.perf_counter() - t0) / 6
        print(f"ramp {ramp:4.2f} flat {flat:2d} groups {g:2d}: {1e3 * dt:7.2f} ms per batch of {n} windows", flush=True)

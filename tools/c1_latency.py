"""Single-window (C1 / C2) latency for different chunk / cluster policies: env VISFS_BA_CHUNKS, VISFS_BA_CLUSTER are read at upload."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from visfs_b200 import capi, synth  # noqa: E402

ba = capi.BundleAdjuster(0, profile_kernels=True)
for name, w in (("C1", synth.config_c1()), ("C2", synth.config_c2())):
    for chunks, cl in ((None, 4), (128, 4), (120, 8), (128, 8), (96, 4), (64, 4), (144, 4), (112, 8), (64, 8)):
        if chunks is None:
            os.environ.pop("VISFS_BA_CHUNKS", None)
        else:
            os.environ["VISFS_BA_CHUNKS"] = str(chunks)
        os.environ["VISFS_BA_CLUSTER"] = str(cl)
        ba.upload([w])
        ts = []
        for _ in range(12):
            ba.run_resident()
            ts.append(ba.timing())
        ts = ts[2:]
        med = lambda k: float(np.median([t[k] for t in ts]))
        print(f"{name} chunks {chunks} cluster {cl}: total {med('total_ms'):.3f} build {med('build_ms'):.3f} solve {med('solve_ms'):.3f} "
              f"update {med('update_ms'):.3f} other {med('other_ms'):.3f}", flush=True)

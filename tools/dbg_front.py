import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from visfs_b200 import capi, synth
P, L = int(sys.argv[1]), int(sys.argv[2])
ba = capi.BundleAdjuster(0, profile_kernels=True)
w = synth.config_c4(n_poses=P, n_points=L)
try:
    r = ba.solve(w)
    print("P", P, "L", L, "ok status", r["status"], r["iterations_run"], r["trials_run"], r["chi2_final"], flush=True)
except Exception as e:
    print("P", P, "L", L, "FAIL", e, flush=True)

"""visfs_b200 — B200-native (sm_100a) local stereo/RGBD bundle adjustment behind VISFS's
`Optimizer::localOptimize` interface.  See DESIGN.md.

  csrc/   CUDA kernels + the C ABI of include/visfs_ba.h  (libvisfs_ba.so)
  host/   C++17 mirror of VISFS::Optimizer::Optimizer (the reference-facing plugin)
  capi.py ctypes binding used by tests/ and bench.py
  synth.py seeded synthetic windows
"""
__all__ = ["capi", "synth"]

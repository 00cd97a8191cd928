#!/usr/bin/env python
"""Benchmark of the B200 bundle adjustment (BASELINE.json metric: BA LM iterations/s and edges/s).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle on the host cores

Workload (config.workload): C3 of BASELINE.json — a batch of 4096 independent 10-key-frame / 2 000-landmark /
20 000-stereo-edge windows (C1 windows with different seeds; `--windows` = the TOTAL over all GPUs, sharded
4096 / N per rank: strong scaling, no collective), each solved with the reference's defaults (Iterations=10 ->
5+5, Huber 8, PixelVariance 1.5, LM, direct solver).  One step = one two-pass local BA of every window.

  value  LM iterations/s over all windows and GPUs, inputs resident in HBM (visfs_ba_run_resident),
         timed with CUDA events on the library's stream, max over ranks
  e2e    the same through visfs_ba_solve_batch with HOST buffers: H2D + LM + D2H inside the timed region
  roofline / fp64   build kernel (linearise + Hessian + Schur) against the measured HBM peak and an
         FP64 FMA probe measured in the same process
  single_window     C1 as ONE window (the latency case the >= 50x target is quoted on) and C2
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from visfs_b200 import synth  # noqa: E402

METRIC = "BA LM iterations/sec"
UNIT = "LM iterations/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def shard(total, world, rank):
    """Windows [lo, hi) of the global batch that rank `rank` of `world` solves."""
    return (total * rank) // world, (total * (rank + 1)) // world


def make_windows(total, world, rank):
    """This rank's shard of the global C3 batch: window k of the batch has the same seed whatever the rank count."""
    lo, hi = shard(total, world, rank)
    return [synth.make_window(10, 2000, 10, seed=synth.BASE_SEED + 3 + 1000 * (k + 1), layout="all") for k in range(lo, hi)]


def flops_per_trial(w):
    """634 E + L (50 + 144 d + 108 d (d + 1)) with d = E / L."""
    E, L = int(w["n_edges"]), max(int(w["n_points"]), 1)
    d = E / L
    return 634.0 * E + L * (50.0 + 144.0 * d + 108.0 * d * (d + 1.0))


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock and clock-event reasons of one GPU, sampled every few ms DURING the timed region by an in-process NVML
    thread (the timed calls are ctypes calls that release the GIL); falls back to an `nvidia-smi -lms` child process."""
    NAMES = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, device):
        self.device, self.sm, self.mx, self.reasons = device, [], [], set()
        self.proc = self.thread = self.nv = self.handle = None
        self.stop = threading.Event()
        try:   # NVML is initialised here, outside the timed region
            self.handle = self._nvml_handle()
            self.mx.append(float(self.nv.nvmlDeviceGetMaxClockInfo(self.handle, self.nv.NVML_CLOCK_SM)))
            self.nv.nvmlDeviceGetClockInfo(self.handle, self.nv.NVML_CLOCK_SM)
        except Exception:
            self.handle = None

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        self.nv = pynvml
        try:
            import torch
            uuid = str(torch.cuda.get_device_properties(self.device).uuid)
            return pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        except Exception:
            return pynvml.nvmlDeviceGetHandleByIndex(self.device)

    def _sample(self):
        nv, h = self.nv, self.handle
        reasons_fn = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        try:
            self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
            bits = int(reasons_fn(h))
            for bit, name in self.NAMES.items():
                if bits & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def _poll(self):
        while True:
            self._sample()
            if self.stop.wait(0.004):
                break

    def _read_smi(self):
        for line in self.proc.stdout:
            r = [c.strip() for c in line.split(",")]
            try:
                self.sm.append(float(r[0])); self.mx.append(float(r[1]))
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[3:7]):
                    if v.lower().startswith("active"):
                        self.reasons.add(name)
            except Exception:
                pass

    def __enter__(self):
        self.t_enter = time.perf_counter()
        if self.handle is not None:
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return self
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device),
                 "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                 "clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read_smi, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *exc):
        if self.handle is not None:
            self._sample()          # (the GPU is still at its loaded clocks right after the last synchronise; an NVML call
            self.stop.set()         #  of the polling thread occasionally blocks for tens of ms under load)
            self.thread.join(timeout=1.0)
            if os.environ.get("VISFS_BENCH_DEBUG"):
                print(f"[clocks] {len(self.sm)} samples in {time.perf_counter() - self.t_enter:.3f} s", file=sys.stderr)
        elif self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm),
                "source": "nvml, 4 ms period, during the timed region" if self.handle is not None else "nvidia-smi -lms 100"}


# ---------------------------------------------------------------------------------------------- CPU arm
def cpu_solve_windows(windows, threads):
    """The oracle on the host cores: one window per thread (liboracle.so releases the GIL under ctypes)."""
    from tests import oracle_api as O
    O.lib(False)
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        res = list(ex.map(lambda w: O.solve_timed(w, threads=1), windows))
    dt = time.perf_counter() - t0
    iters = sum(sum(r["iterations_run"]) for r in res)
    trials = sum(sum(r["trials_run"]) for r in res)
    return dt, iters, trials, res


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    n_sample = max(cores, min(16 * cores, args.windows))
    windows = make_windows(args.windows, 1, 0)[:n_sample] if n_sample >= args.windows else \
        make_windows(n_sample, 1, 0)
    for _ in range(min(args.warmup, 1)):
        cpu_solve_windows(windows[:cores], cores)
    tot_t, tot_it, tot_tr = 0.0, 0, 0
    for _ in range(args.steps):
        dt, it, tr, _ = cpu_solve_windows(windows, cores)
        tot_t += dt; tot_it += it; tot_tr += tr
    value = tot_it / tot_t
    sample = (f"the first {n_sample} of the {args.windows} C1 windows per step, one window per host thread, "
              f"oracle/ba_oracle.cpp (g2o-equivalent CPU port, -O3; the reference's g2o path cannot be built here)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args),
            "edges_per_s": tot_tr * 20000 / tot_t,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


def workload_config(args):
    world = max(int(os.environ.get("WORLD_SIZE", args.gpus)), 1)
    per_gpu = args.windows / world
    return {"workload": f"C3 (BASELINE.json configs[2]): batch of {args.windows} independent C1 windows (10 key frames / 2000 landmarks / "
                        f"20000 stereo edges each) sharded over {world} GPU(s), two-pass LM 5+5 iterations, Huber 8, outlier culling",
            "windows": args.windows, "windows_per_gpu": per_gpu, "poses": 10, "landmarks": 2000, "edges": 20000, "iterations": 10,
            "cache": "inputs larger than L2 (%.0f MB of edge/point/pose records per GPU per trial)" % (per_gpu * 1.43),
            "parallelism": f"windows sharded over {world} GPU(s) ({args.windows} / {world} each), no collective"}


# ---------------------------------------------------------------------------------------------- GPU arm
def single_window_numbers(ba, O, cores, quick, ba_e2e=None):
    """C1 and C2 as single windows: device LM time, end-to-end call time, CPU oracle time.  ba: handle with per-phase CUDA
    events (kernel_ms breakdown, device_ms); ba_e2e: plain handle for the end-to-end calls (the events cost ~0.2 ms per call)."""
    out = {}
    ba_e2e = ba_e2e or ba
    # c0: the window the reference actually runs (corelib/include/Parameters.h:148,161: LocalMap/MapSize 5 + 1 frames,
    # Tracker/MaxFeature 300 -> <= 1 800 edges); c1 / c2: BASELINE.json configs[0] / configs[1]
    for name, w in (("c0", synth.make_window(6, 300, layout="all", seed=synth.BASE_SEED + 10)), ("c1", synth.config_c1()), ("c2", synth.config_c2())):
        packed = ba_e2e.prepare_batch([w], pinned=True, float_obs=True)
        for _ in range(3):
            ba_e2e.solve_packed(packed)
        reps = 5 if quick else 50
        t0 = time.perf_counter()
        for _ in range(reps):
            ba_e2e.solve_packed(packed)
        e2e_ms = 1e3 * (time.perf_counter() - t0) / reps
        ba_e2e.upload([w])
        dev = []
        for _ in range(reps + 3):
            ba_e2e.run_resident()
            dev.append(ba_e2e.timing()["total_ms"])
        dev_ms = float(np.median(dev[3:]))       # device time of the LM passes, handle without per-phase events
        ba.upload([w])
        for _ in range(3):
            ba.run_resident()
        t = ba.timing()                          # per-phase breakdown (its events add ~0.2 ms to the total)
        iters = t["lm_iterations"]
        entry = {"lm_iterations": iters, "lm_trials": t["lm_trials"], "device_ms": dev_ms, "e2e_ms": e2e_ms,
                 "device_iters_per_s": iters / (dev_ms * 1e-3), "e2e_iters_per_s": iters / (e2e_ms * 1e-3),
                 "kernel_ms": {k: t[k] for k in ("build_ms", "solve_ms", "update_ms", "other_ms")},
                 "launches": t["kernel_launches"]}
        if O is not None:
            r1 = min((O.solve_timed(w, threads=1) for _ in range(2)), key=lambda r: r["seconds"])
            rn = min((O.solve_timed(w, threads=cores, omp=True) for _ in range(2)), key=lambda r: r["seconds"])
            cit = sum(r1["iterations_run"])
            assert cit == iters, "CPU and GPU ran different iteration counts"
            entry["cpu_1thread_iters_per_s"] = cit / r1["seconds"]
            entry["cpu_allcores_iters_per_s"] = cit / rn["seconds"]
            entry["speedup_e2e_vs_cpu_allcores"] = entry["e2e_iters_per_s"] / max(entry["cpu_allcores_iters_per_s"], 1e-30)
            entry["speedup_e2e_vs_cpu_1thread"] = entry["e2e_iters_per_s"] / entry["cpu_1thread_iters_per_s"]
        try:   # the same window through the reference-facing C++ class (std::map in, std::map out): VISFS::Optimizer::Optimizer
            import tempfile
            from tests import host_io
            if os.path.exists(host_io.EXE):
                with tempfile.TemporaryDirectory() as td:
                    fin = os.path.join(td, "w.bin")
                    host_io.write_window(fin, w)
                    th = host_io.run_time(fin, os.path.join(td, "o.bin"))
                entry["local_optimize_ms"] = th["local_optimize_ms_mean"]
                entry["local_optimize_marshal_ms"] = th["marshal_ms"]
        except Exception as e:   # the C++ mirror is optional for the bench line
            entry["local_optimize_error"] = str(e)[:200]
        out[name] = entry
    return out


def resident_window_numbers(ba, quick, size="c0", _warm=True):
    """per-frame cost of the reference's real window (6 frames, ~300 features per frame) when the local
    map stays in HBM and only LocalMap's deltas travel (visfs_ba_window_*), against handing the whole window over every frame
    (visfs_ba_solve with page-locked float buffers).  Same sequence, same results (tests/test_gpu_window.py)."""
    from visfs_b200 import capi
    # size "c0": the reference's real window (LocalMap/MapSize 5 + 1 frames, ~300 features per frame); "c2": a C2-sized map
    # (10 frames, ~10 000 features per frame, ~100 000 edges per solve), where the host-side map walk dominates a full call
    if _warm:   # one untimed replay first: the handle's device buffers grow to the size of this map once, as in a long session
        resident_window_numbers(ba, quick, size, _warm=False)
    n_frames, views, per_frame_new = ((16 if quick else 40), 6, 50) if size == "c0" else ((14 if quick else 24), 10, 1000)
    seq = synth.make_window(n_frames, per_frame_new * n_frames, views=views, layout="consecutive", seed=synth.BASE_SEED + 11)
    max_pts, max_obs = (4096, 32768) if size == "c0" else (32768, 262144)
    first_seen = np.full(seq["n_points"], 10**9)
    np.minimum.at(first_seen, seq["edge_point"], seq["edge_pose"])
    win = capi.ResidentWindow(ba, views + 1, max_pts, max_obs, fx=seq["fx"], fy=seq["fy"], cx=seq["cx"], cy=seq["cy"], bf=seq["bf"],
                              pixel_variance=seq["pixel_variance"], huber_delta=seq["huber_delta"], iterations=seq["iterations"])
    # the full-call arm has a handle of its own, like the resident map has (Optimizer and ResidentLocalMap each create theirs):
    # the two arms do not take turns on one set of device buffers
    ba_full = capi.BundleAdjuster(device=ba.device, profile_kernels=False)
    # every array a frame hands over is converted BEFORE the clock starts (the harness' numpy work is not the product's)
    import ctypes as C
    lib, W = ba.lib, win.w
    I64, F64, F32, U8 = capi._i64p, capi._dp, C.POINTER(C.c_float), capi._u8p
    per_frame = []
    for f in range(n_frames):
        sel = np.nonzero(seq["edge_pose"] == f)[0]
        new = np.unique(seq["edge_point"][sel][first_seen[seq["edge_point"][sel]] == f])
        per_frame.append(dict(
            fid=int(seq["pose_id"][f]), tq=np.ascontiguousarray(seq["pose_tq"][f]),
            new_id=np.ascontiguousarray(seq["point_id"][new], dtype=np.int64), new_xyz=np.ascontiguousarray(seq["point_xyz"][new]),
            pid=np.ascontiguousarray(seq["point_id"][seq["edge_point"][sel]], dtype=np.int64),
            ob=np.ascontiguousarray(seq["edge_obs"][sel], dtype=np.float32), kind=np.ascontiguousarray(seq["edge_kind"][sel], dtype=np.uint8)))
    cap = max_obs
    r_fid, r_tq = np.zeros(views + 1, dtype=np.int64), np.zeros((views + 1, 7))
    r_op, r_of = np.zeros(cap, dtype=np.int64), np.zeros(cap, dtype=np.int64)
    res = capi.WindowResult(frame_id=capi._ptr(r_fid, I64), pose_tq=capi._ptr(r_tq, F64), outlier_point_id=capi._ptr(r_op, I64),
                            outlier_frame_id=capi._ptr(r_of, I64), outlier_capacity=cap)
    frames = []
    t_res, t_full, h2d_res, h2d_full, edges, solves = 0.0, 0.0, 0, 0, 0, 0
    for f in range(n_frames):
        d = per_frame[f]
        fid = d["fid"]
        before = win.h2d_bytes_total()
        t0 = time.perf_counter()
        if len(d["new_id"]):
            ba._check(lib.visfs_ba_window_set_points(W, len(d["new_id"]), capi._ptr(d["new_id"], I64), capi._ptr(d["new_xyz"], F64), None))
        ba._check(lib.visfs_ba_window_insert_frame(W, fid, capi._ptr(d["tq"], F64), len(d["pid"]), capi._ptr(d["pid"], I64),
                                                   capi._ptr(d["ob"], F32), capi._ptr(d["kind"], U8)))
        frames.append(f)
        if len(frames) > views:
            old = frames.pop(0)
            ba._check(lib.visfs_ba_window_remove_frame(W, int(seq["pose_id"][old])))
        r = None
        if len(frames) >= 2:
            ba._check(lib.visfs_ba_window_solve(W, fid - 1, C.byref(res)), allow=(0, 3, 4))
            r = {"n_edges": res.n_edges}
            k = min(res.n_outliers, cap)
            if k:
                ba._check(lib.visfs_ba_window_remove_observations(W, k, capi._ptr(r_op, I64), capi._ptr(r_of, I64)))
        dt = time.perf_counter() - t0
        if r is None or len(frames) < views:
            continue
        t_res += dt; h2d_res += win.h2d_bytes_total() - before; edges += r["n_edges"]; solves += 1
        # the same window handed over whole (the plain C-ABI call the C++ shim makes), marshalled before the clock starts
        keep = np.isin(seq["edge_pose"], frames)
        lm = np.unique(seq["edge_point"][keep])
        remap = -np.ones(seq["n_points"], dtype=np.int64); remap[lm] = np.arange(len(lm))
        fmap = -np.ones(seq["n_poses"], dtype=np.int64); fmap[frames] = np.arange(len(frames))
        w = dict(seq)
        w.update(n_poses=len(frames), n_points=len(lm), n_edges=int(keep.sum()), pose_tq=seq["pose_tq"][frames], pose_id=seq["pose_id"][frames],
                 pose_fixed=(seq["pose_id"][frames] == fid - 1).astype(np.uint8), point_xyz=seq["point_xyz"][lm], point_id=seq["point_id"][lm],
                 point_fixed=seq["point_fixed"][lm], edge_obs=seq["edge_obs"][keep], edge_pose=fmap[seq["edge_pose"][keep]].astype(np.int32),
                 edge_point=remap[seq["edge_point"][keep]].astype(np.int32), edge_kind=seq["edge_kind"][keep])
        for k in ("link_from", "link_to", "link_tq", "n_links"):
            w.pop(k, None)
        packed = ba_full.prepare_batch([w], pinned=True, float_obs=True)
        ba_full.solve_packed(packed)
        t1 = time.perf_counter()
        ba_full.solve_packed(packed)
        t_full += time.perf_counter() - t1
        h2d_full += int(ba_full.timing()["h2d_bytes"])
    win.close()
    ba_full.close()
    n = max(solves, 1)
    return {"workload": f"{n_frames}-frame sequence through a {views}-frame local map, ~{edges // n} edges per solve, root = newest - 1",
            "solves": solves, "per_frame_ms_resident": 1e3 * t_res / n, "per_frame_ms_full_call": 1e3 * t_full / n,
            "h2d_bytes_per_frame_resident": h2d_res // n, "h2d_bytes_per_frame_full_call": h2d_full // n,
            "note": "resident: set_points + insert_frame + remove_frame + solve + remove_observations per frame; full call: one visfs_ba_solve "
                    "on a pre-marshalled page-locked window.  Neither clock contains host marshalling: the reference's std::map walk that the "
                    "resident map makes unnecessary is single_window.*.local_optimize_marshal_ms (c0: 0.02 ms, c2: 1.3 ms per call).  At the "
                    "c0 size both are bound by the ~58 launches of the two-pass LM and the resident path adds ~10 short launches that build "
                    "the window on the device; it pays off from C2-sized maps up, where the map walk is more than the whole GPU call"}


def global_ba_numbers(ba, dist, world, rank, local, barrier, reduce_max, reduce_min, quick, use_oracle):
    """BASELINE config C4: one global BA (2 000 key frames on a loop / 500 000 landmarks / 5 M edges), landmarks partitioned
    over the ranks, reduced camera system summed with one ncclAllReduce per LM trial (strong scaling: total work fixed).

    Parity on the path the driver runs: at N > 1 EVERY rank also solves the unpartitioned problem on its own GPU (the 1-rank
    path) and compares decisions, chi2, poses and the outlier flags of its partition; at N = 1 the result is compared with
    the multi-threaded CPU oracle (skipped with --no-cpu / --quick)."""
    from visfs_b200 import capi, partition
    scale = 0.25 if quick else 1.0
    w = synth.config_c4(n_poses=int(2000 * scale), n_points=int(500000 * scale))
    part = partition.partition_window(w, world, rank)
    if world > 1:
        ba.comm_init_torch(dist)
    ba.upload([part])
    ba.run_resident()
    barrier()
    reps = 2
    dev_ms, t = 0.0, None
    t0 = time.perf_counter()
    for _ in range(reps):
        ba.run_resident()
        t = ba.timing()
        dev_ms += t["total_ms"]
    barrier()
    wall = time.perf_counter() - t0
    dev_s = reduce_max(dev_ms * 1e-3)
    r = ba.download()[0]
    if world > 1:       # all ranks leave the communicator together, before rank 0 goes on alone
        barrier()
        ba.comm_destroy()
    out = {"workload": f"C4 global BA: {w['n_poses']} key frames on a loop, {w['n_points']} landmarks x 10 views = {w['n_edges']} "
                       f"stereo edges, landmarks partitioned over {world} GPU(s), NCCL all-reduce of the block-skyline reduced system",
           "scaling": "strong", "lm_iterations": int(t["lm_iterations"]), "lm_trials": int(t["lm_trials"]),
           "ms_per_solve": 1e3 * dev_s / reps, "lm_iterations_per_s": reps * t["lm_iterations"] / dev_s,
           "kernel_ms_this_rank": {k: t[k] for k in ("build_ms", "solve_ms", "update_ms", "other_ms")},
           "wall_ms_per_solve": 1e3 * wall / reps, "status": int(r["status"]),
           "chi2": [r["chi2_initial"], r["chi2_pass1"], r["chi2_final"]]}
    # HBM roofline of the Jacobian + Schur pass: algorithmic bytes of this rank's build launches / their time
    if t["build_ms"] > 0:
        out["build_GBps_this_rank"] = t["alg_bytes_build"] / (t["build_ms"] * 1e-3) / 1e9
    out["_edge_trials_this_rank"] = float(t["edge_trials"])
    out["_dev_s"] = dev_s

    # ---- parity of this very run
    ok, what, detail = 1.0, None, {}
    try:
        if world > 1:
            solo = capi.BundleAdjuster(device=local)
            ref = solo.solve(w)
            solo.close()
            what = "1-rank solve of the same C4 on every rank's own GPU"
            idx = part["part_edge_index"]
            l0, l1 = part["part_range"]
            same = same_decisions(r, ref, r["edge_level"], ref["edge_level"][idx])
            chi_rel = max(abs(r[k] - ref[k]) / max(abs(ref[k]), 1e-300) for k in ("chi2_initial", "chi2_pass1", "chi2_final"))
            pose_abs = float(np.abs(r["pose_tq"] - ref["pose_tq"]).max())
            point_abs = float(np.abs(r["point_xyz"] - ref["point_xyz"][l0:l1]).max())
            ok = 1.0 if (same and chi_rel <= 1e-6 and pose_abs <= 1e-6 and point_abs <= 1e-5) else 0.0
            detail = {"chi2_max_rel": chi_rel, "pose_max_abs": pose_abs, "point_max_abs_this_rank": point_abs, "same_decisions_and_outliers": bool(same)}
        elif use_oracle:
            from tests import oracle_api as O
            cores = os.cpu_count() or 1
            t_cpu = time.perf_counter()
            ref = O.solve(w, threads=cores, omp=True)
            out["cpu_oracle_s"] = time.perf_counter() - t_cpu
            out["cpu_oracle_threads"] = cores
            what = f"CPU oracle (OpenMP, {cores} threads) solving the same C4"
            same = same_decisions(r, ref, r["edge_level"], ref["edge_level"])
            chi_rel = max(abs(r[k] - ref[k]) / max(abs(ref[k]), 1e-300) for k in ("chi2_initial", "chi2_pass1", "chi2_final"))
            pose_abs = float(np.abs(r["pose_tq"] - ref["pose_tq"]).max())
            ok = 1.0 if (same and chi_rel <= 1e-6 and pose_abs <= 1e-6) else 0.0
            detail = {"chi2_max_rel": chi_rel, "pose_max_abs": pose_abs, "same_decisions_and_outliers": bool(same)}
    except Exception as exc:   # a failed check is a failed check, never a silent pass
        ok, detail = 0.0, {"error": repr(exc)[:300]}
    if what is not None:
        all_ok = reduce_min(ok)
        out["parity"] = "ok" if all_ok >= 1.0 else "FAILED"
        out["parity_vs"] = what
        out["parity_detail_rank0"] = detail
    else:
        out["parity"] = "not checked (--no-cpu / --quick at N = 1)"
    return out


def same_decisions(r, ref, lev_r, lev_ref):
    """Equal LM decisions and outlier sets.  A pass whose damping ran away (lambda > 1e12: steps below 1e-12 of the state, the
    sign of the gain ratio is rounding noise — C4's first pass, see tests/test_gpu_large.py) is compared by its accepted state
    only, like the full-size parity test does."""
    for k in (0, 1):
        if max(r["lambda_final"][k], ref["lambda_final"][k]) > 1e12:
            continue
        if r["iterations_run"][k] != ref["iterations_run"][k] or r["trials_run"][k] != ref["trials_run"][k]:
            return False
    return bool(np.array_equal(lev_r, lev_ref))


def dense_window_numbers(ba, quick):
    """BASELINE config C5: 200 key frames orbiting one scene, 200 000 landmarks x 10 random views (dense reduced system)."""
    w = synth.config_c5(n_poses=200, n_points=50000 if quick else 200000)
    ba.upload([w])
    runs = []
    for _ in range(4):          # one warm-up, then the median of three (the build adds with L2 atomics: 10-20 % run-to-run spread
        ba.run_resident()       # was observed right after the GPU had idled through the CPU oracle of C4)
        runs.append(ba.timing())
    t = sorted(runs[1:], key=lambda x: x["total_ms"])[1]
    r = ba.download()[0]
    return {"workload": f"C5 dense window: {w['n_poses']} key frames, {w['n_points']} landmarks, {w['n_edges']} edges, one GPU",
            "lm_iterations": int(t["lm_iterations"]), "lm_trials": int(t["lm_trials"]), "ms_per_solve": t["total_ms"],
            "lm_iterations_per_s": t["lm_iterations"] / (t["total_ms"] * 1e-3),
            "edges_per_s": t["edge_trials"] / (t["total_ms"] * 1e-3),
            "edges_per_s_note": "active edge-trials (culled edges of pass 2 not counted) / device time",
            "kernel_ms": {k: t[k] for k in ("build_ms", "solve_ms", "update_ms", "other_ms")}, "status": int(r["status"]),
            "chi2": [r["chi2_initial"], r["chi2_pass1"], r["chi2_final"]]}


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from visfs_b200 import capi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the bundle adjustment has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce(v, op):
        if world == 1:
            return v
        t = torch.tensor([float(v)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())

    MAX = dist.ReduceOp.MAX if world > 1 else None
    SUM = dist.ReduceOp.SUM if world > 1 else None
    MIN = dist.ReduceOp.MIN if world > 1 else None

    t_gen = time.perf_counter()
    windows = make_windows(args.windows, world, rank)
    log(f"[rank {rank}] generated {len(windows)} windows in {time.perf_counter() - t_gen:.1f}s")
    ba = capi.BundleAdjuster(device=local, profile_kernels=True)
    if args.profile_run:   # short deterministic launch sequence for ncu: upload + two device-resident solves
        ba.upload(windows)
        ba.run_resident()
        ba.run_resident()
        print(json.dumps({"profile_run": True, "timing": ba.timing()}), flush=True)
        return 0
    fp64_peak = ba.probe_fp64()
    # the end-to-end calls go through a handle WITHOUT the per-phase CUDA events that `ba` records for the kernel breakdown
    ba_e2e = capi.BundleAdjuster(device=local, profile_kernels=False)
    packed = ba_e2e.prepare_batch(windows, pinned=True, float_obs=True)

    # ---- end to end through the C ABI with host buffers (H2D + LM + D2H in the timed region)
    for _ in range(args.warmup):
        ba_e2e.solve_packed(packed)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ba_e2e.solve_packed(packed)
    barrier()
    e2e_s = reduce(time.perf_counter() - t0, MAX)
    t_e2e = ba_e2e.timing()
    res = packed[2]
    iters_step = sum(r.iterations_run[0] + r.iterations_run[1] for r in res)
    bad = sum(1 for r in res if r.status != 0)

    # ---- device-resident: inputs already in HBM
    ba.upload(windows)
    clocks = ClockSampler(local)
    for _ in range(args.warmup):
        ba.run_resident()
    barrier()
    dev_ms, tims = 0.0, []
    with clocks:
        for _ in range(args.steps):
            ba.run_resident()
            t = ba.timing()
            tims.append(t)
            dev_ms += t["total_ms"]
    barrier()
    dev_s = reduce(dev_ms * 1e-3, MAX)
    tot_iters = reduce(sum(t["lm_iterations"] for t in tims), SUM)
    tot_trials = reduce(sum(t["lm_trials"] for t in tims), SUM)
    tot_edge_trials = reduce(sum(t["edge_trials"] for t in tims), SUM)
    e2e_iters = reduce(iters_step * args.steps, SUM)
    n_bad = reduce(bad, SUM)

    # ---- global BA (C4) over all ranks: every rank takes part, rank 0 reports
    gba = None
    if not args.no_global:
        try:
            gba = global_ba_numbers(ba, dist if world > 1 else None, world, rank, local, barrier, lambda v: reduce(v, MAX),
                                    lambda v: reduce(v, MIN), args.quick, use_oracle=not (args.no_cpu or args.quick))
            gba["edges_per_s"] = reduce(gba.pop("_edge_trials_this_rank"), SUM) * 2 / gba.pop("_dev_s")
            gba["edges_per_s_note"] = "active edge-trials of all ranks (culled edges of pass 2 not counted) / device time"
        except Exception as exc:  # reported, never silently dropped
            gba = {"error": repr(exc)}
            try:
                ba.comm_destroy()
            except Exception:
                pass

    shard_parity = None
    if world > 1 and not args.no_cpu:
        # N > 1: every rank checks a sample of ITS shard against the oracle (the host cores are shared by the ranks, so not all
        # windows as at N = 1); the flags are combined over the ranks
        cores = max(1, (os.cpu_count() or 1) // world)
        n_sample = min(len(windows), 16 * cores)
        _, _, _, cres = cpu_solve_windows(windows[:n_sample], cores)
        gres = ba.packed_results(packed)
        worst, bad = 0.0, 0
        for k in range(n_sample):
            c, g = cres[k], gres[k]
            same = (c["iterations_run"] == g["iterations_run"] and c["trials_run"] == g["trials_run"] and c["status"] == g["status"]
                    and np.array_equal(c["edge_level"], g["edge_level"]))
            if not same:
                bad += 1
                continue
            worst = max(worst, abs(c["chi2_final"] - g["chi2_final"]) / max(abs(c["chi2_final"]), 1e-300),
                        float(np.abs(c["pose_tq"] - g["pose_tq"]).max()))
        v = torch.tensor([float(bad), float(n_sample)], dtype=torch.float64, device="cuda")
        m = torch.tensor([worst], dtype=torch.float64, device="cuda")
        dist.all_reduce(v, op=dist.ReduceOp.SUM)
        dist.all_reduce(m, op=dist.ReduceOp.MAX)
        parity_ok = v[0].item() == 0.0 and m.item() <= 1e-6
        shard_parity = {"status": "ok" if parity_ok else "FAILED", "windows_checked": int(v[1].item()), "of": args.windows,
                          "against": "oracle/ba_oracle.cpp on the host cores, a sample of every rank's shard of the timed batch",
                          "max_rel_chi2_or_abs_pose": m.item(), "windows_with_other_decisions": int(v[0].item()),
                          "gate": "same iterations / trials / outlier set, chi2 and poses 1e-6"}
        assert parity_ok, f"GPU and CPU results differ: {shard_parity}"

    if rank != 0:
        ba.close()
        if world > 1:
            dist.destroy_process_group()
        return 0

    build_ms = sum(t["build_ms"] for t in tims)
    build_n = sum(t["build_launches"] for t in tims)
    alg_build = sum(t["alg_bytes_build"] for t in tims)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = alg_build / (build_ms * 1e-3) / 1e9 if build_ms > 0 else 0.0
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "build_kernel_traffic.json")))
        # (one ncu --set full capture on the 4096-window batch, stored per window: a launch of this run holds len(windows))
        traffic = tj["dram_bytes_per_window"] * len(windows) if "dram_bytes_per_window" in tj else tj.get("dram_bytes_per_launch")
    except Exception:
        pass
    # FP64 view of the same kernel (the binding roofline per SURVEY.md §8d): Schur + Hessian flops of one trial
    # (active edges only: pass 2 runs without the edges culled after pass 1; the mean landmark degree of the whole run is
    #  used in the quadratic Schur term, which slightly under-counts)
    trials_rank0 = sum(t["lm_trials"] for t in tims)
    et_rank0 = sum(t["edge_trials"] for t in tims)
    L_win = int(windows[0]["n_points"])
    d_mean = et_rank0 / max(trials_rank0 * L_win, 1)
    flops_build = 594.0 * et_rank0 + trials_rank0 * L_win * (50.0 + 144.0 * d_mean + 108.0 * d_mean * (d_mean + 1.0))
    fp64_achieved = flops_build / (build_ms * 1e-3) / 1e12 if build_ms > 0 else 0.0

    line = {"metric": METRIC, "value": tot_iters / dev_s, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dev_s / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(args),
            "edges_per_s": tot_edge_trials / dev_s, "lm_trials_per_s": tot_trials / dev_s,
            "e2e": {"value": e2e_iters / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(t_e2e["h2d_bytes"]),
                    "d2h_bytes_per_step": int(t_e2e["d2h_bytes"]), "ms_per_step": 1e3 * e2e_s / args.steps},
            "gpu_launches": int(sum(t["kernel_launches"] for t in tims)),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": traffic, "kernel": "ws::k_build_ws (linearise + Hessian blocks + Schur partials; k_build<MODE_BUILD> is its fallback for 20-32 poses)",
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                         "alg_bytes_per_launch": alg_build / max(build_n, 1), "avg_launch_ms": build_ms / max(build_n, 1),
                         "note": "the kernel sits right of the FP64 ridge (28 flop/B): the binding roofline is the FP64 pipe, see the fp64 object"},
            "fp64": {"kernel": "ws::k_build_ws", "achieved_tflops": fp64_achieved, "peak_tflops": fp64_peak,
                     "frac": fp64_achieved / fp64_peak if fp64_peak else None,
                     "peak_source": "visfs_ba_probe_fp64: dependent-free DFMA loop, this process, this GPU (MEASURED_PEAKS.json has no FP64 "
                                    "entry; NVIDIA's nominal B200 FP64 figure is 37-40 TFLOP/s)", "nominal_tflops": 40.0,
                     "frac_of_nominal": fp64_achieved / 40.0},
            "kernel_ms_per_step": {k: sum(t[k] for t in tims) / args.steps for k in ("build_ms", "solve_ms", "update_ms", "other_ms")},
            "clocks": clocks.summary(), "windows_failed": int(n_bad)}
    if gba is not None:
        line["global_ba"] = gba
    if world == 1 and not args.no_global:
        try:
            line["dense_window"] = dense_window_numbers(ba, args.quick)
        except Exception as exc:  # reported, never silently dropped
            line["dense_window"] = {"error": repr(exc)}

    # ---- CPU baseline (rank 0, N = 1 only) and the single-window latency cases
    if world == 1 and not args.no_cpu:
        from tests import oracle_api as O
        cores = os.cpu_count() or 1
        n_sample = len(windows) if not args.quick else min(len(windows), 4 * cores)
        n_sample = max(min(n_sample, 256 * cores), min(cores, len(windows)))     # ~20 s of CPU work: all 4096 windows from 16 cores up
        cpu_solve_windows(windows[:cores], cores)
        dt, it, tr, cres = cpu_solve_windows(windows[:n_sample], cores)
        # every CPU-solved window against the GPU result of the timed end-to-end batch: equal LM decisions (otherwise the
        # comparison of rates is void), same outlier set, chi2 and poses inside the north_star gate (1e-6 relative)
        gres = ba.packed_results(packed)
        worst_chi, worst_pose, mismatched = 0.0, 0.0, []
        for k in range(n_sample):
            c, g = cres[k], gres[k]
            same = (c["iterations_run"] == g["iterations_run"] and c["trials_run"] == g["trials_run"] and c["status"] == g["status"]
                    and np.array_equal(c["edge_level"], g["edge_level"]))
            if not same:
                mismatched.append(k)
                continue
            worst_chi = max(worst_chi, abs(c["chi2_final"] - g["chi2_final"]) / max(abs(c["chi2_final"]), 1e-300))
            worst_pose = max(worst_pose, float(np.abs(c["pose_tq"] - g["pose_tq"]).max()))
        parity_ok = not mismatched and worst_chi <= 1e-6 and worst_pose <= 1e-6
        line["parity"] = {"status": "ok" if parity_ok else "FAILED", "windows_checked": n_sample, "of": len(windows),
                          "against": "oracle/ba_oracle.cpp on the host cores, same windows as the timed batch",
                          "chi2_final_max_rel": worst_chi, "pose_max_abs": worst_pose, "windows_with_other_decisions": mismatched[:8],
                          "gate": "same iterations / trials / outlier set, chi2 and poses 1e-6"}
        assert parity_ok, f"GPU and CPU results differ: {line['parity']}"
        line["cpu_baseline"] = {"value": it / dt, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{n_sample} of the {args.windows} windows of one step, one window per host thread, "
                                          f"oracle/ba_oracle.cpp (g2o-equivalent CPU port; the reference's g2o path cannot be built here)"}
        line["single_window"] = single_window_numbers(ba, O, cores, args.quick, ba_e2e)
    elif world == 1:
        line["single_window"] = single_window_numbers(ba, None, 1, args.quick, ba_e2e)
    if shard_parity is not None:
        line["parity"] = shard_parity
    if world == 1:
        try:
            line["resident_window"] = resident_window_numbers(ba_e2e, args.quick)
            line["resident_window_c2"] = resident_window_numbers(ba_e2e, args.quick, size="c2")
        except Exception as exc:  # reported, never silently dropped
            line["resident_window"] = {"error": repr(exc)[:300]}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--windows", type=int, default=4096, help="windows of the whole batch (BASELINE.json: 4096), sharded over the GPUs")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--no-global", action="store_true", help="skip the C4 global-BA and C5 dense-window legs")
    ap.add_argument("--profile-run", action="store_true", help="upload + two resident solves only (for ncu)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())

"""World-size-2 test of the multi-GPU host logic on CPU (gloo): the landmark partition used by the global BA.

Each rank takes its landmark range of one window (visfs_b200.partition), forms its partial reduced camera system with the
CPU oracle, the ranks sum them with an all-reduce — the exchange step the CUDA library performs with ncclAllReduce on
the block-skyline buffer — and the sum has to equal the reduced system of the unpartitioned window.  The GPU side of the
same path is covered by tests/test_gpu_large.py::test_global_ba_two_ranks_nccl.
"""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

LAMBDA = 0.75


def _window():
    from visfs_b200 import synth
    return synth.make_window(6, 240, layout="all", seed=1234, mono_frac=0.2, fixed_point_frac=0.1)


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tests import oracle_api as O
    from visfs_b200 import partition
    w = _window()
    part = partition.partition_window(w, world, rank)
    assert part["n_poses"] == w["n_poses"] and (part["flags"] & 1)
    r = O.reduced_system(part, LAMBDA)
    n = r["n"]
    buf = torch.from_numpy(np.concatenate([r["S"].reshape(-1), r["b_s"], [r["chi2"]]]))
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    S = buf[: n * n].numpy().reshape(n, n) - (world - 1) * LAMBDA * np.eye(n)   # every rank damped its own diagonal
    counts = torch.tensor([part["n_points"], part["n_edges"]], dtype=torch.int64)
    dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    if rank == 0:
        np.savez(os.path.join(out_dir, "sum.npz"), S=S, b=buf[n * n: n * n + n].numpy(), chi2=buf[-1].item(), counts=counts.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_partitioned_reduced_systems_sum_to_the_full_one(tmp_path):
    from tests import oracle_api as O
    O.lib(False)   # build the oracle once, before the workers race for it
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / "sum.npz")
    w = _window()
    ref = O.reduced_system(w, LAMBDA)
    assert list(got["counts"]) == [w["n_points"], w["n_edges"]]
    scale = np.abs(ref["S"]).max()
    assert np.abs(got["S"] - ref["S"]).max() <= 1e-12 * scale
    assert np.abs(got["b"] - ref["b_s"]).max() <= 1e-12 * np.abs(ref["b_s"]).max()
    assert abs(got["chi2"] - ref["chi2"]) <= 1e-12 * ref["chi2"]


def test_partition_and_merge_round_trip():
    from visfs_b200 import partition
    w = _window()
    parts = [partition.partition_window(w, 3, r) for r in range(3)]
    assert sum(p["n_points"] for p in parts) == w["n_points"] and sum(p["n_edges"] for p in parts) == w["n_edges"]
    fake = [dict(pose_tq=w["pose_tq"], point_xyz=p["point_xyz"], edge_level=(p["edge_pose"] % 2).astype(np.uint8), n_outliers=int((p["edge_pose"] % 2).sum()),
                 status=0) for p in parts]
    m = partition.merge_results(w, parts, fake)
    assert np.array_equal(m["point_xyz"], w["point_xyz"])
    assert np.array_equal(m["edge_level"], (w["edge_pose"] % 2).astype(np.uint8))
    assert m["n_outliers"] == int((w["edge_pose"] % 2).sum())
    # ragged: more ranks than landmarks leaves empty partitions
    tiny = {k: v for k, v in w.items()}
    empty = partition.partition_window(tiny, 1000, 1)
    assert empty["n_points"] in (0, 1) and empty["n_edges"] == empty["n_points"] * 6

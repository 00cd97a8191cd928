"""Run under torchrun (one rank per GPU): landmark-partitioned global BA with the NCCL all-reduce of the reduced camera
system, checked on rank 0 against the CPU oracle solving the unpartitioned problem.  Used by
tests/test_gpu_large.py::test_global_ba_two_ranks_nccl and by hand:
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/global_ba_ranks.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from visfs_b200 import capi, partition, synth
    from tests import oracle_api as O
    from tests.test_gpu_parity import check_solution

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ba = capi.BundleAdjuster(device=local)
    ba.comm_init_torch(dist)

    w = synth.make_window(48, 3000, views=7, layout="consecutive", trajectory="loop", seed=91, mono_frac=0.2, fixed_point_frac=0.1,
                          links="chain")
    part = partition.partition_window(w, world, rank)
    res = ba.solve(part)
    # gather the per-rank results on rank 0
    gathered = [None] * world
    dist.all_gather_object(gathered, {k: v for k, v in res.items()})
    if rank == 0:
        parts = [partition.partition_window(w, world, r) for r in range(world)]
        merged = partition.merge_results(w, parts, gathered)
        ref = O.solve(w)
        check_solution(merged, ref, f"global BA on {world} ranks")
        for r in range(1, world):
            assert np.array_equal(gathered[r]["pose_tq"], gathered[0]["pose_tq"]), "ranks disagree on the poses"
        print("GLOBAL_BA_OK", world, "ranks, chi2", merged["chi2_initial"], "->", merged["chi2_final"], flush=True)
    dist.barrier()
    ba.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

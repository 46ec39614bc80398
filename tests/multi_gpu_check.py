"""torchrun entry: every rank runs its partition on its own GPU over NCCL; rank 0 checks the union of results against
the oracle.  python -m torch.distributed.run --nproc-per-node N tests/multi_gpu_check.py"""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def main():
    import oracle_lib as O
    from sst_core_b200 import models
    from sst_core_b200.run import simulate_distributed

    lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    rank, world = dist.get_rank(), dist.get_world_size()
    cases = [
        ("mesh", models.mesh_model(96, 64, stop_at="60ns", stats_enabled=True), {}),
        ("phold", models.phold_model(64, 48, stop_at="80ns", stats_enabled=True), {"calendar_slots": 16}),
        ("clockpop", models.clockpop_model(40, 40, stop_at="300ns", comm_freq=5, stats_enabled=True), {}),
    ]
    for name, m, opts in cases:
        res, st, pm = simulate_distributed(m, **opts)
        gathered = [None] * world
        dist.all_gather_object(gathered, (res, st.events_delivered, st.clock_ticks, st.end_time))
        if rank == 0:
            ref = O.run_model(m)
            allres = np.concatenate([g[0] for g in gathered], axis=0)
            got = np.zeros_like(allres)
            got[pm.orig_id] = allres
            assert np.array_equal(got, ref.results), f"{name}: results differ from the oracle"
            assert sum(g[1] for g in gathered) == ref.events, name
            assert sum(g[2] for g in gathered) == ref.ticks, name
            assert all(g[3] == ref.end_time for g in gathered), name
            print(f"{name}: {ref.events} events over {world} GPUs bit-exact", flush=True)
        dist.barrier()
    if rank == 0:
        print("MULTI_GPU_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

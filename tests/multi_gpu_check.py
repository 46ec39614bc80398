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
    # both exchange modes: peer stores over NVLink + mailbox (default), and one NCCL all-to-all per window
    for mode in ("peer", "nccl"):
      os.environ.pop("SSTB200_NO_PEER", None)
      if mode == "nccl":
          os.environ["SSTB200_NO_PEER"] = "1"
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
            print(f"{name} [{mode}]: {ref.events} events over {world} GPUs bit-exact", flush=True)
        dist.barrier()
    # a mesh of bench-like size against the committed digest of the CPU golden model (tests/golden/bench_digests.json)
    import json
    from sst_core_b200.digest import records_digest
    os.environ.pop("SSTB200_NO_PEER", None)
    fix = json.loads((ROOT / "tests" / "golden" / "bench_digests.json").read_text())["digests"]["mesh_1024x1024_w13_stats1"]
    m = models.mesh_model(1024, 1024, stop_at="13ns", verbose=0, stats_enabled=True)
    res, st, pm = simulate_distributed(m, calendar_slots=2, bin_capacity=2048)
    c0 = int(np.searchsorted(pm.rank, rank))
    part = torch.tensor([np.array([records_digest(pm.orig_id[c0:c0 + res.shape[0]], res[:, :6])], dtype=np.uint64)
                         .view(np.int64)[0], st.events_delivered], dtype=torch.int64, device="cuda")
    dist.all_reduce(part)
    if rank == 0:
        dig = int(np.array([int(part[0].item())], dtype=np.int64).view(np.uint64)[0])
        assert f"{dig:016x}" == fix["digest"] and int(part[1].item()) == fix["events"], "mesh 1024^2 digest differs"
        print(f"mesh 1024x1024 x 13 windows over {world} GPUs: digest {dig:016x} == CPU golden model", flush=True)
    dist.barrier()
    if rank == 0:
        print("MULTI_GPU_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

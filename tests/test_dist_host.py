"""The N>1 host path on CPU: two processes over gloo run partitions of a model through DistributedRunner with the
oracle as the per-rank engine; the union of their results must equal the serial oracle (the reference's own contract:
a serial run is the reference for the parallel run, tests/testsuite_default_partitioner.py:33-63)."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _worker(rank, world, port, kind, out_dir):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    import torch.distributed as dist

    from oracle_rank_engine import OracleRankEngine
    from sst_core_b200 import models
    from sst_core_b200.engine import DistributedRunner, partition_model

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m = _model(kind, models)
    pm = partition_model(m, world)
    e = OracleRankEngine(pm, rank, world)
    e.setup()
    DistributedRunner(e).run(check_every=8)
    st = e.status()
    np.save(os.path.join(out_dir, f"res_{rank}.npy"), e.results())
    np.save(os.path.join(out_dir, f"meta_{rank}.npy"),
            np.array([e.c0, e.c1, st.events_delivered, st.clock_ticks, st.end_time], dtype=np.int64))
    np.save(os.path.join(out_dir, f"orig_{rank}.npy"), pm.orig_id)
    dist.destroy_process_group()


def _model(kind, models):
    if kind == "mesh":
        return models.mesh_model(6, 6, stop_at="60ns", stats_enabled=True)
    if kind == "phold":
        m = models.phold_model(5, 4, stop_at="90ns", init_events=6, delay_mod=2500, stats_enabled=True)
        m.nocut_pairs = np.array([[0, 19], [3, 4]], dtype=np.uint32)  # forces a non-trivial renumbering
        return m
    return models.clockpop_model(4, 4, stop_at="700ns", comm_freq=7, stats_enabled=True)


@pytest.mark.parametrize("kind", ["mesh", "phold", "clockpop"])
def test_two_rank_gloo_run_equals_serial_oracle(tmp_path, kind):
    import oracle_lib as O
    from sst_core_b200 import models

    world, port = 2, 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, kind, str(tmp_path)), nprocs=world, join=True)
    ref = O.run_model(_model(kind, models))
    orig = np.load(tmp_path / "orig_0.npy")
    got = np.zeros_like(ref.results)
    ev = ticks = 0
    for r in range(world):
        c0, c1, e, t, end = np.load(tmp_path / f"meta_{r}.npy")
        got[orig[c0:c1]] = np.load(tmp_path / f"res_{r}.npy")
        ev, ticks = ev + e, ticks + t
        assert end == ref.end_time
    assert np.array_equal(got, ref.results)
    assert ev == ref.events and ticks == ref.ticks


def _report_worker(rank, world, port, kind, rate, out_dir):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    import torch.distributed as dist

    from oracle_rank_engine import OracleRankEngine
    from sst_core_b200 import models
    from sst_core_b200.engine import DistributedRunner, partition_model

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pm = partition_model(_model(kind, models), world)
    e = OracleRankEngine(pm, rank, world)
    e.setup()
    runner, t, snaps = DistributedRunner(e), rate, []
    while True:
        rec, fin = runner.run_to_report(t)
        if rec is None:
            break
        snaps.append(rec)
        t += rate
        if fin:
            break
    runner.run()
    np.save(os.path.join(out_dir, f"snaps_{rank}.npy"), np.array(snaps))
    np.save(os.path.join(out_dir, f"res_{rank}.npy"), e.results())
    np.save(os.path.join(out_dir, f"meta_{rank}.npy"), np.array([e.c0, e.c1], dtype=np.int64))
    np.save(os.path.join(out_dir, f"orig_{rank}.npy"), pm.orig_id)
    dist.destroy_process_group()


@pytest.mark.parametrize("kind,rate", [("mesh", 7000), ("phold", 13300), ("clockpop", 99999)])
def test_two_rank_periodic_reports_equal_the_serial_oracle(tmp_path, kind, rate):
    """DistributedRunner.run_to_report over gloo: every rank brackets the window that contains the report time; the
    union of the ranks' snapshots equals the serial oracle stepped to the same times (report times inside windows)"""
    from sst_core_b200 import models
    from test_periodic_stats import oracle_periodic

    world, port = 2, 31500 + (os.getpid() % 2000)
    mp.spawn(_report_worker, args=(world, port, kind, rate, str(tmp_path)), nprocs=world, join=True)
    m = _model(kind, models)
    m.stat_rate = rate
    want, final = oracle_periodic(m)
    orig = np.load(tmp_path / "orig_0.npy")
    snaps = [np.load(tmp_path / f"snaps_{r}.npy") for r in range(world)]
    assert all(len(s) == len(want) for s in snaps) and len(want) >= 3
    for i, (_, w) in enumerate(want):
        got = np.zeros_like(w)
        for r in range(world):
            c0, c1 = np.load(tmp_path / f"meta_{r}.npy")
            got[orig[c0:c1]] = snaps[r][i]
        assert np.array_equal(got, w), (kind, i)
    got = np.zeros_like(final)
    for r in range(world):
        c0, c1 = np.load(tmp_path / f"meta_{r}.npy")
        got[orig[c0:c1]] = np.load(tmp_path / f"res_{r}.npy")
    assert np.array_equal(got, final)

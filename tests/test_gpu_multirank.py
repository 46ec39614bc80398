"""The multi-rank device path: partitions of a model stepped through outbox emission / ingest / gathered control words.
On one GPU all ranks are stepped by one process (LocalMultiRankRunner); with >= 2 GPUs the same check runs as real
NCCL ranks under torchrun (tests/multi_gpu_check.py)."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

import oracle_lib as O
from sst_core_b200 import models

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _check_local(m, n_ranks, **opts):
    from sst_core_b200.engine import LocalMultiRankRunner
    ref = O.run_model(m)
    r = LocalMultiRankRunner(m, n_ranks, **opts)
    try:
        r.setup()
        assert r.run()
        res = r.results()
        sts = [e.status() for e in r.engines]
    finally:
        r.close()
    got = np.zeros_like(res)
    got[r.pm.orig_id] = res
    bad = np.nonzero((got != ref.results).any(axis=1))[0]
    assert bad.size == 0, f"{bad.size} components differ, first {bad[:5]}"
    assert sum(s.events_delivered for s in sts) == ref.events
    assert sum(s.clock_ticks for s in sts) == ref.ticks
    assert all(s.end_time == ref.end_time for s in sts)


# peer=True: boundary events stored straight into the owner's calendar + mailbox (the NVLink peer mode, here inside one
# process); peer=False: outboxes, exchanged by copies, ingest kernel (the path a collective-based driver uses)
@pytest.mark.parametrize("peer", [True, False])
@pytest.mark.parametrize("n_ranks", [2, 3, 8])
def test_mesh_partitions_on_one_gpu(n_ranks, peer):
    _check_local(models.mesh_model(32, 24, stop_at="80ns", stats_enabled=True), n_ranks, peer=peer)


def test_mesh_partitions_larger_blocks_cross_ranks_event_parallel_form():
    # 96 x 64 = 24 blocks of 256 components over 3 ranks: the event-parallel kernel with remote destinations in both modes
    for peer in (True, False):
        _check_local(models.mesh_model(96, 64, stop_at="120ns", stats_enabled=True), 3, peer=peer)


@pytest.mark.parametrize("peer", [True, False])
@pytest.mark.parametrize("n_ranks", [2, 4])
def test_phold_partitions_on_one_gpu(n_ranks, peer):
    m = models.phold_model(20, 20, stop_at="100ns", stats_enabled=True)
    m.nocut_pairs = np.array([[0, 399], [5, 200]], dtype=np.uint32)  # non-contiguous no-cut groups: renumbering
    _check_local(m, n_ranks, calendar_slots=16, peer=peer)


def test_clockpop_partitions_on_one_gpu():
    _check_local(models.clockpop_model(16, 16, stop_at="500ns", comm_freq=3, stats_enabled=True), 4)


def test_real_multi_gpu_nccl_ranks():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (covered on one GPU by the LocalMultiRankRunner tests)")
    n = 2 if n < 4 else 4
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29617", str(ROOT / "tests" / "multi_gpu_check.py")],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MULTI_GPU_OK" in r.stdout


def test_cli_under_torchrun_prints_what_the_reference_prints(tmp_path):
    """python -m torch.distributed.run -m sst_core_b200 model.py: one process per GPU, console output and
    StatisticOutput.csv against the reference's golden file / fixtures"""
    import json
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    n = 2 if n < 4 else 4
    gold = json.loads((ROOT / "tests" / "golden" / "reference_golden.json").read_text())
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", PYTHONPATH=str(ROOT))

    def cli(*args):
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                            "--master-addr", "127.0.0.1", "--master-port", "29619", "-m", "sst_core_b200", *args],
                           capture_output=True, text=True, timeout=600, env=env, cwd=tmp_path)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
        return r.stdout

    out = cli("--model-options", "6 6", str(ROOT / "tests/models/mesh.py"))
    want = [f"{i} received {c} messages" for i, c in gold["mesh_6x6_10us"]["counts"].items()]
    want.append("Simulation is complete, simulated time: 10 us")
    assert sorted(l for l in out.strip().splitlines() if "received" in l or "Simulation is" in l) == sorted(want)
    cli("--stop-at", "100ns", "--model-options", "3 3 1 1 2 1", str(ROOT / "tests/models/mesh.py"))
    rows = (tmp_path / "StatisticOutput.csv").read_text().strip().splitlines()
    ref_rows = gold["mesh_3x3_100ns_csv"]["rows"]
    # the Rank column holds the GPU that owns the component here (the fixture was written by a one-rank run)
    strip_rank = lambda rr: [", ".join(x.split(", ")[:5] + x.split(", ")[6:]) for x in rr]
    assert strip_rank(rows) == strip_rank(ref_rows)


def test_cli_under_torchrun_profiling_checkpoints_and_restart_on_other_rank_counts(tmp_path):
    import json
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    gold = json.loads((ROOT / "tests" / "golden" / "reference_golden.json").read_text())
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", PYTHONPATH=str(ROOT))

    def cli(n, *args):
        head = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr",
                "127.0.0.1", "--master-port", "29621", "-m", "sst_core_b200"] if n > 1 else [sys.executable, "-m", "sst_core_b200"]
        r = subprocess.run([*head, *args], capture_output=True, text=True, timeout=600, env=env, cwd=tmp_path)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
        return r.stdout

    keep = lambda text: sorted(l for l in text.splitlines() if "received" in l or "Simulation is" in l)
    mesh = str(ROOT / "tests/models/mesh.py")
    # handler counts gathered from both ranks: the reference's profiling golden (component level)
    cli(2, "--model-options", "4 4", "--profiling-output", "p.txt", "--enable-profiling",
        "events:sst.profile.handler.event.count(level=component)[event]", mesh)
    strip = lambda t: sorted(l for l in t.splitlines() if not l.startswith("  Simulation Input File"))
    assert strip((tmp_path / "p.txt").read_text()) == strip(gold["profiling_4x4_event_component"]["text"])
    # checkpoints written by 2 ranks; restart on 1 GPU (2 -> 1) and under torchrun (2 -> 2)
    want = sorted([f"{i} received {c} messages" for i, c in gold["mesh_6x6_10us"]["counts"].items()]
                  + ["Simulation is complete, simulated time: 10 us"])
    out = cli(2, "--model-options", "6 6", "--checkpoint-sim-period", "4us", "--checkpoint-prefix", "two", mesh)
    assert keep(out) == want
    d = tmp_path / "two" / "two_1_4000000"
    assert sorted(p.name for p in d.iterdir()) == ["two_1_4000000.sstcpt", "two_1_4000000_0_0.bin", "two_1_4000000_1_0.bin"]
    assert keep(cli(1, "--load-checkpoint", str(d / "two_1_4000000.sstcpt"))) == want
    assert keep(cli(2, "--load-checkpoint", str(d / "two_1_4000000.sstcpt"))) == want
    # a single-GPU checkpoint restarted on 2 GPUs (1 -> 2)
    cli(1, "--model-options", "6 6", "--checkpoint-sim-period", "7us", "--checkpoint-prefix", "one", mesh)
    assert keep(cli(2, "--load-checkpoint", str(tmp_path / "one" / "one_1_7000000" / "one_1_7000000.sstcpt"))) == want


def test_cli_under_torchrun_periodic_statistics(tmp_path):
    import json
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    gold = json.loads((ROOT / "tests" / "golden" / "reference_golden.json").read_text())
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", PYTHONPATH=str(ROOT))
    key = "rate_phold_3x3_100ns"
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                        "127.0.0.1", "--master-port", "29623", "-m", "sst_core_b200", "--calendar-slots", "16",
                        "--model-options", gold[key]["opts"], str(ROOT / "tests/models/phold.py")],
                       capture_output=True, text=True, timeout=600, env=env, cwd=tmp_path)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    rows = (tmp_path / "StatisticOutput.csv").read_text().strip().splitlines()
    strip_rank = lambda rr: [", ".join(x.split(", ")[:5] + x.split(", ")[6:]) for x in rr]
    assert strip_rank(rows) == strip_rank(gold[key]["rows"])

"""Checkpoint / restart of the device state (reference: Simulation::checkpoint / restart,
simulation.cc:2015-2200, testsuite_default_Checkpoint.py): a run that is checkpointed, torn down and restarted ends
bit-identically to an uninterrupted run -- and to the oracle."""
import json
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

import oracle_lib as O
from sst_core_b200 import models
from sst_core_b200.engine import Engine, EngineError, LocalMultiRankRunner
from sst_core_b200.run import result_words

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
GOLD = json.loads((ROOT / "tests" / "golden" / "reference_golden.json").read_text())


def _finish(e):
    e.run()
    st = e.status()
    e.raise_on_error(st)
    return e.results(), st


CASES = [
    ("mesh", lambda: models.mesh_model(24, 24, stop_at="200ns", stats_enabled=True), dict(), [1, 7, 60, 150]),
    ("phold", lambda: models.phold_model(16, 16, stop_at="300ns", stats_enabled=True),
     dict(calendar_slots=16), [1, 5, 33, 250]),
    ("clockpop", lambda: models.clockpop_model(8, 8, stop_at="1us", comm_freq=10, stats_enabled=True),
     dict(calendar_slots=4), [3, 400]),
]


@pytest.mark.parametrize("name,build,opts,cuts", CASES, ids=[c[0] for c in CASES])
def test_restart_continues_bit_identically(name, build, opts, cuts):
    m = build()
    ref = O.run_model(m)
    for k in cuts:
        a = Engine(m, handler_counts=True, **opts)
        a.setup()
        a.run(max_windows=k, check_every=k)
        blob = a.checkpoint().copy()
        a.close()
        b = Engine(m, handler_counts=True, **opts)
        b.restore(blob)
        assert b.status().windows == k
        res, st = _finish(b)
        recv, calls = b.handler_counts()
        b.close()
        assert np.array_equal(res, ref.results), (name, k)
        assert (st.events_delivered, st.clock_ticks, st.end_time) == (ref.events, ref.ticks, ref.end_time)
        assert int(recv.sum()) == ref.events and int(calls.sum()) == ref.ticks  # counters travel with the state


def test_checkpoint_of_a_checkpointed_run_and_smaller_staging():
    """restart twice; the staging capacity (a launch parameter, not state) may change across a restart"""
    m = models.phold_model(12, 12, stop_at="200ns", stats_enabled=True)
    ref = O.run_model(m)
    e = Engine(m, calendar_slots=16, smem_events=2048)
    e.setup()
    for k in (10, 50):
        e.run(max_windows=k, check_every=k)
        blob = e.checkpoint().copy()
        e.close()
        e = Engine(m, calendar_slots=16, smem_events=1024)
        e.restore(blob)
    res, st = _finish(e)
    e.close()
    assert np.array_equal(res, ref.results) and st.events_delivered == ref.events and st.windows == 200


def test_multi_rank_restart():
    m = models.phold_model(12, 10, stop_at="150ns", stats_enabled=True)
    ref = O.run_model(m)
    r = LocalMultiRankRunner(m, 3, calendar_slots=16)
    r.setup()
    r.run(max_windows=40, check_every=40)
    blobs = [e.checkpoint().copy() for e in r.engines]
    r.close()
    r2 = LocalMultiRankRunner(m, 3, calendar_slots=16)
    for e, b in zip(r2.engines, blobs):
        e.restore(b)
    r2.run()
    res = r2.results()
    back = np.empty_like(res)
    back[r2.pm.orig_id] = res
    r2.close()
    assert np.array_equal(back, ref.results)


def test_mismatched_or_foreign_blobs_are_refused():
    m = models.mesh_model(8, 8, stop_at="50ns")
    a = Engine(m)
    with pytest.raises(EngineError, match="before engine_setup"):
        a.checkpoint()
    a.setup()
    a.run(max_windows=8, check_every=8)
    blob = a.checkpoint().copy()
    a.close()
    b = Engine(m, bin_capacity=4096)
    with pytest.raises(EngineError, match="different model slice or calendar geometry"):
        b.restore(blob)
    with pytest.raises(EngineError, match="not a checkpoint"):
        b.restore(np.zeros(4096, dtype=np.uint8))
    b.close()
    c = Engine(models.mesh_model(8, 9, stop_at="50ns"))
    with pytest.raises(EngineError, match="different model slice"):
        c.restore(blob)
    c.close()
    t = Engine(m, trace_capacity=100000)
    t.setup()
    with pytest.raises(EngineError, match="tracing and checkpoints"):
        t.checkpoint()
    t.close()


def test_restart_at_full_width_matches_the_uninterrupted_run():
    """size-independent property at 1024 x 1024 (1 M components, 2.9 GB of state): checkpoint -> new engine -> same end"""
    m = models.mesh_model(1024, 1024, stop_at="14ns", verbose=0, stats_enabled=True)
    w = result_words(m)
    a = Engine(m, bin_capacity=2048)
    a.setup()
    a.run(max_windows=6, check_every=6)
    blob = a.checkpoint()
    a.run()
    res_a, st_a = a.results(w, compact=True), a.status()
    a.close()
    b = Engine(m, bin_capacity=2048)
    b.restore(blob)
    del blob
    b.run()
    res_b, st_b = b.results(w, compact=True), b.status()
    b.close()
    assert np.array_equal(res_a, res_b) and st_a.events_delivered == st_b.events_delivered
    assert st_b.end_time == 14000 and int(res_b[:, 0].astype(np.int64).sum()) == st_b.events_delivered


def _cli(args, cwd):
    r = subprocess.run([sys.executable, "-m", "sst_core_b200", *args], capture_output=True, text=True, cwd=cwd,
                       env={**__import__("os").environ, "PYTHONPATH": str(ROOT)}, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def test_cli_checkpoint_directories_and_restart(tmp_path):
    """--checkpoint-sim-period / --checkpoint-prefix / --load-checkpoint as in the reference (config.cc:195-250):
    the restarted run prints what the reference's golden file holds for the uninterrupted one"""
    out = _cli(["--model-options", "6 6", "--checkpoint-sim-period", "4us", "--checkpoint-prefix", "ck",
                str(ROOT / "tests/models/mesh.py")], tmp_path)
    want = [f"{i} received {c} messages" for i, c in GOLD["mesh_6x6_10us"]["counts"].items()]
    want.append("Simulation is complete, simulated time: 10 us")
    lines = out.strip().splitlines()
    assert [l.split("(")[0] for l in lines if l.startswith("#")] == [
        "# Simulation Checkpoint: Simulated Time 4 us ", "# Simulation Checkpoint: Simulated Time 8 us "]
    assert sorted(l for l in lines if not l.startswith("#")) == sorted(want)
    d = tmp_path / "ck" / "ck_2_8000000"
    assert sorted(p.name for p in d.iterdir()) == ["ck_2_8000000.sstcpt", "ck_2_8000000_0_0.bin"]
    out2 = _cli(["--load-checkpoint", str(tmp_path / "ck" / "ck_1_4000000" / "ck_1_4000000.sstcpt")], tmp_path)
    assert sorted(out2.strip().splitlines()) == sorted(want)


@pytest.mark.parametrize("n_old,n_new", [(3, 1), (3, 2), (1, 4), (2, 5)])
def test_restart_on_another_number_of_ranks(n_old, n_new):
    """testsuite_default_Checkpoint.py:175-193 restarts runs on other rank counts: the old ranks' blobs are re-cut by
    checkpoint.repartition_checkpoints; the restarted run ends bit-identically to the oracle (ranks share one GPU here)"""
    from sst_core_b200 import checkpoint
    from sst_core_b200.engine import partition_model
    for m, opts, k in [(models.phold_model(24, 23, stop_at="120ns", stats_enabled=True), dict(calendar_slots=16), 37),
                       (models.mesh_model(28, 21, stop_at="90ns", stats_enabled=True), dict(handler_counts=True), 40),
                       (models.clockpop_model(25, 22, stop_at="300ns", comm_freq=5, stats_enabled=True), dict(calendar_slots=4), 111)]:
        ref = O.run_model(m)
        if n_old > 1:
            a = LocalMultiRankRunner(m, n_old, **opts)
            a.setup()
            a.run(max_windows=k, check_every=k)
            assert np.array_equal(a.pm.orig_id, np.arange(m.ncomp))  # sst.linear kept the numbering
            blobs = [e.checkpoint().copy() for e in a.engines]
            a.close()
        else:
            e = Engine(m, **opts)
            e.setup()
            e.run(max_windows=k, check_every=k)
            blobs = [e.checkpoint().copy()]
            e.close()
        pm = partition_model(m, n_new)
        new = checkpoint.repartition_checkpoints(blobs, m, pm.rank, n_new)
        if n_new > 1:
            b = LocalMultiRankRunner(m, n_new, **opts)
            for e, blob in zip(b.engines, new):
                e.restore(blob)
            b.run()
            res = b.results()
            if opts.get("handler_counts"):
                recv = np.concatenate([e.handler_counts()[0] for e in b.engines])
                assert int(recv.sum()) == ref.events
            b.close()
        else:
            e = Engine(m, **opts)
            e.restore(new[0])
            e.run()
            st = e.status()
            e.raise_on_error(st)
            res = e.results()
            assert (st.events_delivered, st.clock_ticks, st.end_time) == (ref.events, ref.ticks, ref.end_time)
            e.close()
        assert np.array_equal(res, ref.results), (n_old, n_new, int(m.comp_type[0]))

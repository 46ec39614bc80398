"""The `sst`-like command line (python -m sst_core_b200) end to end on the GPU: console output and StatisticOutput.csv
against the reference's golden file / fixtures, compared the way the reference's testsuites do (sorted diff)."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
GOLD = json.loads((ROOT / "tests" / "golden" / "reference_golden.json").read_text())


def _run(args, cwd):
    r = subprocess.run([sys.executable, "-m", "sst_core_b200", *args], capture_output=True, text=True, cwd=cwd,
                       env={**__import__("os").environ, "PYTHONPATH": str(ROOT)}, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def test_cli_mesh_6x6_matches_reference_golden_file(tmp_path):
    out = _run(["--model-options", "6 6", str(ROOT / "tests/models/mesh.py")], tmp_path)
    want = [f"{i} received {c} messages" for i, c in GOLD["mesh_6x6_10us"]["counts"].items()]
    want.append("Simulation is complete, simulated time: 10 us")
    assert sorted(out.strip().splitlines()) == sorted(want)


def test_cli_stop_at_override_and_csv(tmp_path):
    out = _run(["--stop-at", "100ns", "--model-options", "3 3 1 1 2 1", str(ROOT / "tests/models/mesh.py")], tmp_path)
    assert out.strip().splitlines()[-1] == "Simulation is complete, simulated time: 100 ns"
    rows = (tmp_path / "StatisticOutput.csv").read_text().strip().splitlines()
    assert rows == GOLD["mesh_3x3_100ns_csv"]["rows"]


def test_cli_unknown_element_is_fatal(tmp_path):
    bad = tmp_path / "bad.py"
    bad.write_text('import sst\nsst.Component("c0", "nolib.nothing")\n')
    r = subprocess.run([sys.executable, "-m", "sst_core_b200", str(bad)], capture_output=True, text=True,
                       env={**__import__("os").environ, "PYTHONPATH": str(ROOT)})
    assert r.returncode == 1 and "can't find requested component" in r.stderr


def test_cli_runs_the_reference_json_interchange_format(tmp_path):
    out = _run([str(ROOT / "tests/golden/mesh_3x2_stats.json")], tmp_path)
    counts = GOLD["mesh_3x2_json_run"]["counts"]
    want = [f"{i} received {c} messages" for i, c in counts.items()] + ["Simulation is complete, simulated time: 40 ns"]
    assert sorted(out.strip().splitlines()) == sorted(want)
    assert (tmp_path / "StatisticOutput.csv").read_text().strip().splitlines() == GOLD["mesh_3x2_json_run"]["rows"]


def test_auto_tuning_retries_after_a_reported_overflow():
    from sst_core_b200 import models
    from sst_core_b200.run import simulate_model
    import oracle_lib as O
    m = models.phold_model(16, 16, stop_at="60ns", init_events=24, stats_enabled=True)
    res, st = simulate_model(m, calendar_slots=2, bin_capacity=1024, smem_events=1024)  # overflows at first
    ref = O.run_model(m)
    assert __import__("numpy").array_equal(res, ref.results) and st.events_delivered == ref.events


def _comparable(text):
    return sorted(l for l in text.splitlines() if not l.startswith("  Simulation Input File"))


@pytest.mark.parametrize("level", ["global", "type", "component", "subcomponent"])
def test_cli_event_handler_counts_match_the_reference_profiling_goldens(tmp_path, level):
    """testsuite_default_profiling.py:36-80 pointed at the device engine: counters kept by dispatch_kernel."""
    _run(["--model-options", "4 4", "--profiling-output", f"prof_{level}.txt", "--enable-profiling",
          f"events:sst.profile.handler.event.count(level={level})[event]", str(ROOT / "tests/models/mesh.py")], tmp_path)
    got = (tmp_path / f"prof_{level}.txt").read_text()
    assert _comparable(got) == _comparable(GOLD[f"profiling_4x4_event_{level}"]["text"])


@pytest.mark.parametrize("key,script,opts,extra", [
    ("profiling_clockpop_3x2_200ns", "clockpop.py", "3 2 200ns 10", []),
    ("profiling_mesh_3x2_50ns_ports", "mesh.py", "3 2", ["--stop-at", "50ns"]),
    ("profiling_phold_4x3_60ns_type", "phold.py", "4 3 60ns", []),
])
def test_cli_handler_counts_match_reference_binary_runs(tmp_path, key, script, opts, extra):
    _run(["--model-options", opts, "--profiling-output", "p.txt", "--enable-profiling", GOLD[key]["spec"], *extra,
          str(ROOT / "tests/models" / script)], tmp_path)
    assert _comparable((tmp_path / "p.txt").read_text()) == _comparable(GOLD[key]["text"])


def test_cli_handler_time_and_sync_tools(tmp_path):
    """sst.profile.handler.{event,clock}.time.* and sst.profile.sync.* on the device: the counts inside the time records
    equal the count tools' (pinned to the reference above), every handler that ran has a non-zero time, and the layout
    is the reference's (keys, alignment)"""
    import re
    spec = ("ev:sst.profile.handler.event.time.steady(level=component)[event];"
            "ck:sst.profile.handler.clock.time.high_resolution(level=component)[clock];"
            "cc:sst.profile.handler.clock.count(level=component)[clock];"
            "ec:sst.profile.handler.event.count(level=component)[event];sy:sst.profile.sync.count[sync]")
    _run(["--model-options", "3 2 200ns 10", "--profiling-output", "p.txt", "--enable-profiling", spec,
          str(ROOT / "tests/models/clockpop.py")], tmp_path)
    text = (tmp_path / "p.txt").read_text()
    rec = {name: body for name, body in re.findall(r"\n\n(\w+):\n(.*?)(?=\n\n\w|\Z)", text, re.S)}
    counts = dict(re.findall(r"    (node\d+):\n      count: (\d+)", rec["cc"]))
    timed = dict(re.findall(r"    (node\d+):\n      avg_handler_time: [0-9.]+ [munp]?s\n      count: +(\d+)\n      handler_time: +[0-9.]+ [munp]?s", rec["ck"]))
    assert counts == timed and len(counts) == 6
    ecounts = dict(re.findall(r"  (node\d+):\n    recv_count: (\d+)", rec["ec"]))
    etimed = dict(re.findall(r"  (node\d+):\n(?:    avg_recv_time: [0-9.]+ [munp]?s\n)?    recv_count: +(\d+)\n    recv_time: +[0-9.]+ [munp]?s", rec["ev"]))
    assert ecounts == etimed and len(ecounts) == 6
    assert "  rank0:\n    ranksync_total_count:   0\n" in rec["sy"]  # one rank: no rank syncs, like the reference


def test_handler_counts_sum_to_the_delivered_events_on_a_large_mesh():
    """size-independent property at a size the oracle does not reach quickly: per-port counts add up to the number of
    deliveries, per component they equal the element's own receive counter (event-parallel and sequential forms)"""
    import numpy as np
    from sst_core_b200 import models
    from sst_core_b200.engine import Engine
    m = models.mesh_model(512, 512, stop_at="30ns", verbose=0)
    for par_max in (0, 2):
        e = Engine(m, handler_counts=True, parallel_max_events=par_max)
        e.setup()
        e.run()
        st = e.status()
        e.raise_on_error(st)
        recv, calls = e.handler_counts()
        res = e.results(1, compact=True)
        e.close()
        assert int(recv.sum()) == st.events_delivered and int(calls.sum()) == 0
        assert np.array_equal(recv.reshape(-1, 4).sum(axis=1), res[:, 0])

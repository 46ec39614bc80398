"""Periodic statistic output (the "rate" statistic parameter; StatisticProcessingEngine::performStatisticOutput at a
statistic clock, statengine.cc:346-376, priority 85: after the clock handlers and events of its time).
Fixtures: StatisticOutput.csv written by the reference binary for the reference's own MessageMesh option ("dump at
rate") and for our probe elements, whose event times do not line up with the lookahead windows.
CPU tier: the oracle stepped to each report time + the CSV writer; GPU tier: the engine through the command line."""
import ctypes as C
import json
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

import oracle_lib as O
from sst_core_b200 import config, output, wireup
from sst_core_b200.elements import NRESULT

ROOT = Path(__file__).resolve().parent.parent
GOLD = json.loads((ROOT / "tests" / "golden" / "reference_golden.json").read_text())
CASES = [("rate_mesh_2x2_600ns", "mesh.py", "600ns"), ("rate_phold_3x3_100ns", "phold.py", None),
         ("rate_clockpop_3x3_1us", "clockpop.py", None)]


def oracle_periodic(m):
    L = O.lib()
    L.oracle_set_ownership.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
    L.oracle_setup_owned.argtypes = [C.c_void_p]
    L.oracle_run_until.argtypes = [C.c_void_p, C.c_uint64]
    L.oracle_run_until.restype = C.c_uint64
    arrs = [np.ascontiguousarray(a) for a in (m.comp_type, m.comp_param, m.port_off, m.peer_port, m.latency, m.clock_period)]
    h = L.oracle_create(m.ncomp, *[a.ctypes.data_as(C.c_void_p) for a in arrs], 0, int(bool(m.stats_enabled)), 0)
    try:
        L.oracle_set_ownership(h, 0, m.ncomp)
        L.oracle_setup_owned(h)

        def records():
            r = np.zeros((m.ncomp, NRESULT), dtype=np.uint64)
            L.oracle_results(h, r.ctypes.data_as(C.c_void_p))
            return r

        out, t = [], m.stat_rate
        while t < m.stop_at:  # an output AT the stop time never happens: StopAction has priority 30
            L.oracle_run_until(h, t + 1)  # everything with time <= t
            out.append((t, records()))
            t += m.stat_rate
        L.oracle_run_until(h, m.stop_at)
        return out, records()
    finally:
        L.oracle_destroy(h)


@pytest.mark.parametrize("key,script,stop", CASES, ids=[c[0] for c in CASES])
def test_oracle_periodic_rows_match_the_reference_csv(key, script, stop):
    m = wireup.wire(config.build_graph(str(ROOT / "tests/models" / script), GOLD[key]["opts"]), stop_at_override=stop)
    assert m.stat_rate > 0
    periodic, final = oracle_periodic(m)
    assert output.stat_rows_csv(m, final, m.stop_at, periodic=periodic) == GOLD[key]["rows"]


def test_rate_handling_in_the_model_surface(tmp_path):
    src = ('import sst\n'
           'a = sst.Component("a", "pdes.phold"); a.addParams({"id": 0})\n'
           'b = sst.Component("b", "pdes.phold"); b.addParams({"id": 1})\n'
           'sst.Link("l").connect((a, "port0", "2ns"), (b, "port0", "2ns"))\n'
           'sst.setProgramOption("stop-at", "50ns")\n%s')
    def wire(tail):
        p = tmp_path / "m.py"
        p.write_text(src % tail)
        return wireup.wire(config.build_graph(str(p)))
    assert wire('sst.enableAllStatisticsForAllComponents({"rate": "7ns"})\n').stat_rate == 7000
    assert wire('sst.enableAllStatisticsForAllComponents()\n').stat_rate == 0
    assert wire('a.enableAllStatistics({"rate": "1GHz"})\n').stat_rate == 1000
    with pytest.raises(config.ModelError, match="different output rates"):
        wire('a.enableAllStatistics({"rate": "5ns"})\nb.enableAllStatistics({"rate": "6ns"})\n')
    with pytest.raises(TypeError, match="must be dict, not NoneType"):  # the reference: model construction fails
        wire('sst.enableAllStatisticsForAllComponents(None)\n')


@pytest.mark.gpu
@pytest.mark.parametrize("key,script,stop", CASES, ids=[c[0] for c in CASES])
def test_gpu_cli_periodic_csv_matches_the_reference(tmp_path, key, script, stop):
    args = ["--model-options", GOLD[key]["opts"]] + (["--stop-at", stop] if stop else [])
    r = subprocess.run([sys.executable, "-m", "sst_core_b200", *args, str(ROOT / "tests/models" / script)],
                       capture_output=True, text=True, cwd=tmp_path,
                       env={**__import__("os").environ, "PYTHONPATH": str(ROOT)}, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert (tmp_path / "StatisticOutput.csv").read_text().strip().splitlines() == GOLD[key]["rows"]


@pytest.mark.gpu
def test_gpu_report_records_match_the_oracle_mid_window_and_in_both_forms():
    """report times that fall inside windows, on window boundaries and one cycle before them; the mesh leaves its
    event-parallel form for the report window only; the run's end result is unchanged by the reports"""
    from sst_core_b200 import models
    from sst_core_b200.engine import Engine
    for m, opts, rate in [(models.phold_model(12, 9, stop_at="90ns", stats_enabled=True), dict(calendar_slots=16), 7300),
                          (models.phold_model(8, 8, stop_at="40ns", stats_enabled=True), dict(calendar_slots=16), 999),
                          (models.clockpop_model(9, 7, stop_at="400ns", comm_freq=7, stats_enabled=True), dict(calendar_slots=4), 33000),
                          (models.mesh_model(20, 20, stop_at="60ns", stats_enabled=True), dict(), 4000),
                          (models.mesh_model(20, 20, stop_at="60ns", stats_enabled=False), dict(), 2500)]:
        m.stat_rate = rate
        want, final = oracle_periodic(m)
        e = Engine(m, **opts)
        e.setup()
        got, t = [], rate
        while True:
            rec, fin = e.run_to_report(t)
            if rec is None:
                break
            got.append((t, rec))
            t += rate
        e.run()
        st = e.status()
        e.raise_on_error(st)
        res = e.results()
        e.close()
        assert [t for t, _ in got] == [t for t, _ in want]
        for (t, a), (_, b) in zip(got, want):
            assert np.array_equal(a, b), (rate, t)
        assert np.array_equal(res, final) and st.end_time == m.stop_at


@pytest.mark.gpu
def test_gpu_periodic_reports_over_three_partitions():
    """the multi-rank bracket (report_begin, dispatch, exchange, ingest, advance, report_end) with the CUDA engine:
    the union of the ranks' snapshots equals the serial oracle at report times inside windows"""
    from sst_core_b200 import models
    from sst_core_b200.engine import LocalMultiRankRunner
    for m, opts, rate in [(models.phold_model(13, 11, stop_at="70ns", stats_enabled=True), dict(calendar_slots=16), 6100),
                          (models.clockpop_model(10, 9, stop_at="300ns", comm_freq=5, stats_enabled=True), dict(calendar_slots=4), 41500),
                          (models.mesh_model(18, 14, stop_at="50ns", stats_enabled=True), dict(), 3000)]:
        m.stat_rate = rate
        want, final = oracle_periodic(m)
        r = LocalMultiRankRunner(m, 3, **opts)
        r.setup()
        t, got = rate, []
        while True:
            rec, fin = r.run_to_report(t)
            if rec is None:
                break
            back = np.empty_like(rec)
            back[r.pm.orig_id] = rec
            got.append(back)
            t += rate
        r.run()
        res = r.results()
        back = np.empty_like(res)
        back[r.pm.orig_id] = res
        r.close()
        assert len(got) == len(want)
        for a, (_, b) in zip(got, want):
            assert np.array_equal(a, b)
        assert np.array_equal(back, final)


def test_only_enabled_statistics_are_written_and_field_types_mix():
    """tests/models/partial_stats.py enables statistics by object, by name and by type on a mixed PHOLD / clocknode
    model: the reference writes only those, u64 and i32 field columns side by side"""
    m = wireup.wire(config.build_graph(str(ROOT / "tests/models/partial_stats.py")))
    assert [len(s) for s in m.stats] == [1, 0, 1, 0, 2, 0]
    run = O.run_model(m)
    assert output.stat_rows_csv(m, run.results, run.end_time) == GOLD["partial_stats_csv"]["rows"]


@pytest.mark.gpu
def test_gpu_cli_partial_statistics(tmp_path):
    r = subprocess.run([sys.executable, "-m", "sst_core_b200", "--calendar-slots", "8", str(ROOT / "tests/models/partial_stats.py")],
                       capture_output=True, text=True, cwd=tmp_path,
                       env={**__import__("os").environ, "PYTHONPATH": str(ROOT)}, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert (tmp_path / "StatisticOutput.csv").read_text().strip().splitlines() == GOLD["partial_stats_csv"]["rows"]

"""coreTestElement.coreTestComponent (testElements/coreTest_Component.cc) through the
graph of the reference's own model script tests/test_Component.py (tests/models/component.py builds it): console output against
tests/refFiles/test_Component.out, handler counts and statistics against runs of the reference binary.
CPU tier: the oracle; GPU tier: the engine through the command line."""
import json
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

import oracle_lib as O
from sst_core_b200 import config, output, profile, wireup

ROOT = Path(__file__).resolve().parent.parent
GOLD = json.loads((ROOT / "tests" / "golden" / "reference_golden.json").read_text())
MODEL = ROOT / "tests" / "models" / "component.py"
REF_SCRIPT = Path("/root/reference/tests/test_Component.py")


def comparable(text):
    return sorted(l for l in text.splitlines() if not l.startswith("  Simulation Input File"))


@pytest.mark.skipif(not REF_SCRIPT.exists(), reason="the reference tree is only present in the build container")
def test_model_script_builds_the_same_graph_as_the_reference_script():
    ours, ref = config.build_graph(str(MODEL)), config.build_graph(str(REF_SCRIPT))
    assert [(c.name, c.type, c.params) for c in ours.components] == [(c.name, c.type, c.params) for c in ref.components]
    ends = lambda g: sorted(tuple(sorted((o.name, p, lat) for o, p, lat in l.ends)) for l in g.links)
    assert ends(ours) == ends(ref) and ours.program_options == ref.program_options
    a, b = wireup.wire(ours), wireup.wire(ref)
    for f in ("comp_type", "comp_param", "port_off", "peer_port", "latency", "clock_period"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f


def test_oracle_console_and_handler_counts_match_the_reference():
    m = wireup.wire(config.build_graph(str(MODEL)))
    assert m.ncomp == 100 and m.stop_at == 25_000_000 and int(m.min_latency()) == 10_000
    run = O.run_model(m, trace=True)
    assert sorted(output.finish_lines(m, run.results, run.end_time)) == sorted(GOLD["component_golden"]["lines"])
    recv = np.zeros(m.nport, dtype=np.uint64)
    np.add.at(recv, m.port_off[run.trace["comp"]].astype(np.int64) + run.trace["port"].astype(np.int64), 1)
    calls = (run.end_time - 1) // m.clock_period
    text = profile.profiling_text(m, profile.parse_enable_profiling(GOLD["component_profile"]["spec"]), recv, calls,
                                  run.end_time, "x")
    assert comparable(text) == comparable(GOLD["component_profile"]["text"])
    assert run.ticks == 100 * 24999


def test_oracle_statistics_match_the_reference_csv():
    m = wireup.wire(config.build_graph(str(MODEL), "10 10 37 3us 1"))
    run = O.run_model(m)
    assert output.stat_rows_csv(m, run.results, run.end_time) == GOLD["component_stats_csv"]["rows"]


def test_constructor_errors_follow_the_reference():
    import textwrap
    def build(params, links=4):
        src = textwrap.dedent(f"""
            import sst
            a = sst.Component("a", "coreTestElement.coreTestComponent"); a.addParams({params!r})
            b = sst.Component("b", "coreTestElement.coreTestComponent"); b.addParams({{"workPerCycle": 1, "commFreq": 5}})
            for i, (p, q) in enumerate([("Nlink", "Slink"), ("Slink", "Nlink"), ("Elink", "Wlink"), ("Wlink", "Elink")][:{links}]):
                sst.Link(f"l{{i}}").connect((a, p, "1ns"), (b, q, "1ns"))
        """)
        path = Path(__import__("tempfile").mkdtemp()) / "m.py"
        path.write_text(src)
        return wireup.wire(config.build_graph(str(path)))
    with pytest.raises(config.ModelError, match="couldn't find work per cycle"):
        build({"commFreq": 5})
    with pytest.raises(config.ModelError, match="couldn't find communication frequency"):
        build({"workPerCycle": 1})
    with pytest.raises(config.ModelError, match="port Wlink is not connected"):
        build({"workPerCycle": 1, "commFreq": 5}, links=3)
    m = build({"workPerCycle": 1, "commFreq": 5, "clockFrequency": "2GHz"})
    assert list(m.clock_period) == [500, 1000]


@pytest.mark.gpu
def test_gpu_cli_runs_the_reference_component_script(tmp_path):
    env = {**__import__("os").environ, "PYTHONPATH": str(ROOT)}
    r = subprocess.run([sys.executable, "-m", "sst_core_b200", "--profiling-output", "p.txt", "--enable-profiling",
                        GOLD["component_profile"]["spec"], str(MODEL)], capture_output=True, text=True, cwd=tmp_path,
                       env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert sorted(r.stdout.strip().splitlines()) == sorted(GOLD["component_golden"]["lines"])
    assert comparable((tmp_path / "p.txt").read_text()) == comparable(GOLD["component_profile"]["text"])
    r = subprocess.run([sys.executable, "-m", "sst_core_b200", "--model-options", "10 10 37 3us 1", str(MODEL)],
                       capture_output=True, text=True, cwd=tmp_path, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert (tmp_path / "StatisticOutput.csv").read_text().strip().splitlines() == GOLD["component_stats_csv"]["rows"]


@pytest.mark.gpu
def test_gpu_engine_matches_oracle_on_a_large_component_torus():
    """same element on a 64 x 48 torus of 3 ns links, mixed clock frequencies, statistics on; 2 ranks as well"""
    from sst_core_b200.engine import LocalMultiRankRunner
    from sst_core_b200.run import simulate_model
    src = '''
import sst
X, Y = 64, 48
c = [[sst.Component(f"c{x}_{y}", "coreTestElement.coreTestComponent") for y in range(Y)] for x in range(X)]
for x in range(X):
    for y in range(Y):
        c[x][y].addParams({"workPerCycle": 0, "commFreq": 3 + (x + y) % 5, "clockFrequency": ["1GHz", "500MHz", "2GHz"][(x * 7 + y) % 3]})
for x in range(X):
    for y in range(Y):
        sst.Link(f"v{x}_{y}").connect((c[x][y], "Nlink", "3ns"), (c[x][(y + 1) % Y], "Slink", "3ns"))
        sst.Link(f"h{x}_{y}").connect((c[x][y], "Elink", "3ns"), (c[(x + 1) % X][y], "Wlink", "3ns"))
sst.setProgramOption("stop-at", "400ns")
sst.enableAllStatisticsForAllComponents()
'''
    path = Path(__import__("tempfile").mkdtemp()) / "big.py"
    path.write_text(src)
    m = wireup.wire(config.build_graph(str(path)))
    ref = O.run_model(m)
    res, st = simulate_model(m)
    assert np.array_equal(res, ref.results) and (st.events_delivered, st.clock_ticks) == (ref.events, ref.ticks)
    r = LocalMultiRankRunner(m, 2)
    r.setup()
    r.run()
    out = r.results()
    back = np.empty_like(out)
    back[r.pm.orig_id] = out
    r.close()
    assert np.array_equal(back, ref.results)

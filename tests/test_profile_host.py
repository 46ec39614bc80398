"""Handler-count profile tools, host half (sst_core_b200/profile.py): option parsing as
Simulation::initializeProfileTools (simulation.cc:1607-1760) and the --profiling-output text, checked on CPU against
the reference's own goldens (tests/refFiles/test_Profiling_event_*.out) and against runs of the reference binary, with
the per-port / per-component counts taken from the oracle's delivery trace (the GPU test feeds the engine's counters
through the same code)."""
import json
from pathlib import Path

import numpy as np
import pytest

import oracle_lib as O
from sst_core_b200 import config, profile, wireup

ROOT = Path(__file__).resolve().parent.parent
GOLD = json.loads((ROOT / "tests" / "golden" / "reference_golden.json").read_text())


def comparable(text):
    """What the reference's testsuite compares: sorted lines without the input-file line
    (testsuite_default_profiling.py:58-80)."""
    return sorted(l for l in text.splitlines() if not l.startswith("  Simulation Input File"))


def oracle_counts(m):
    run = O.run_model(m, trace=True)
    recv = np.zeros(m.nport, dtype=np.uint64)
    np.add.at(recv, m.port_off[run.trace["comp"]].astype(np.int64) + run.trace["port"].astype(np.int64), 1)
    calls = np.zeros(m.ncomp, dtype=np.uint64)
    clocked = m.clock_period != 0
    # ticks of a clocked component that never unregisters: one per period strictly before the end time
    calls[clocked] = (run.end_time - 1) // m.clock_period[clocked]
    return recv, calls, run.end_time


@pytest.mark.parametrize("level", ["global", "type", "component", "subcomponent"])
def test_event_count_levels_match_the_reference_golden_files(level):
    g = config.build_graph(str(ROOT / "tests/models/mesh.py"), "4 4")
    m = wireup.wire(g)
    recv, calls, end = oracle_counts(m)
    tools = profile.parse_enable_profiling(f"events:sst.profile.handler.event.count(level={level})[event]")
    text = profile.profiling_text(m, tools, recv, calls, end, "x.py")
    assert comparable(text) == comparable(GOLD[f"profiling_4x4_event_{level}"]["text"])


@pytest.mark.parametrize("key,script,opts,stop", [
    ("profiling_clockpop_3x2_200ns", "clockpop.py", "3 2 200ns 10", None),
    ("profiling_mesh_3x2_50ns_ports", "mesh.py", "3 2", "50ns"),
    ("profiling_phold_4x3_60ns_type", "phold.py", "4 3 60ns", None),
])
def test_profiling_text_matches_reference_binary_runs(key, script, opts, stop):
    g = config.build_graph(str(ROOT / "tests/models" / script), opts)
    m = wireup.wire(g, stop_at_override=stop)
    recv, calls, end = oracle_counts(m)
    tools = profile.parse_enable_profiling(GOLD[key]["spec"])
    text = profile.profiling_text(m, tools, recv, calls, end, "x.py")
    want = GOLD[key]["text"]
    assert comparable(text) == comparable(want)
    # and byte for byte, input-file line aside
    strip = lambda t: "\n".join(l for l in t.split("\n") if not l.startswith("  Simulation Input File"))
    assert strip(text) == strip(want)


def test_option_string_errors_follow_the_reference():
    P = profile.parse_enable_profiling
    with pytest.raises(config.ModelError, match="Invalid format for argument passed to --enable-profiling"):
        P("no_colon_here")
    with pytest.raises(config.ModelError, match=r"Invalid format for params \(level\)"):
        P("e:sst.profile.handler.event.count(level)[event]")
    with pytest.raises(config.ModelError, match="Cannot reuse tool name: e"):
        P("e:sst.profile.handler.event.count[event];e:sst.profile.handler.clock.count[clock]")
    with pytest.raises(config.ModelError, match="Invalid profile point specified: nothing"):
        P("e:sst.profile.handler.event.count[nothing]")
    with pytest.raises(config.ModelError, match="unsupported level specified for EventHandlerProfileTool: deep"):
        P("e:sst.profile.handler.event.count(level=deep)[event]")
    with pytest.raises(config.ModelError, match="not available on the device engine"):
        P("e:sst.profile.handler.event.time.cpu(level=global)[event]")
    assert P("e:sst.profile.handler.event.time.high_resolution(level=global)[event]")[0].params == {"level": "global"}
    t = P(" b : sst.profile.handler.clock.count [clock] ;a:sst.profile.handler.event.count( level=global )[ event ]")
    assert [x.name for x in t] == ["a", "b"] and t[0].params == {"level": "global"} and t[1].points == ["clock"]


def test_partitioned_models_keep_port_names():
    from sst_core_b200.engine import partition_model
    g = config.build_graph(str(ROOT / "tests/models/mesh.py"), "4 4")
    m = wireup.wire(g)
    pm = partition_model(m, 3)
    names = sorted(i[0] + ":" + i[2] for i in m.port_info)
    assert sorted(i[0] + ":" + i[2] for i in pm.port_info) == names
    for c in range(pm.ncomp):
        for p in range(int(pm.port_off[c]), int(pm.port_off[c + 1])):
            assert pm.port_info[p][0].split(":")[0] == pm.names[c]


def test_time_and_sync_tools_parse_and_format():
    """sst.profile.handler.{event,clock}.time.* and sst.profile.sync.*: keys and layout of the reference's records
    (eventHandlerProfileTool.cc:140-155, clockHandlerProfileTool.cc, syncProfileTool.cc:58-120); the times themselves are
    measurements"""
    import numpy as np
    from sst_core_b200 import config, profile, wireup
    m = wireup.wire(config.build_graph(str(ROOT / "tests" / "models" / "clockpop.py"), "2 2 200ns 10 1 0"))
    spec = ("ev:sst.profile.handler.event.time.steady(level=component)[event];"
            "ck:sst.profile.handler.clock.time.high_resolution(level=type)[clock];"
            "sy:sst.profile.sync.count[sync];st:sst.profile.sync.time.steady[sync]")
    tools = profile.parse_enable_profiling(spec)
    assert [t.name for t in tools] == ["ck", "ev", "st", "sy"] and profile.needs_times(tools)
    recv = np.arange(m.nport, dtype=np.uint64) % 5
    calls = np.full(m.ncomp, 100, dtype=np.uint64)
    times = (recv.astype(np.float64) * 311.0, np.full(m.ncomp, 399000.0))
    text = profile.profiling_text(m, tools, recv, calls, 200000, "x", times=times, sync=[(199, 40, 640, 1500000)])
    assert "ck:\n  0_0:\n    pdes.clocknode:\n      avg_handler_time: 3.99 us\n      count:            400\n      handler_time:     1.596 ms\n" in text
    assert "\nev:\n  node0:\n" in text and "    avg_recv_time: 311 ns\n" in text and "    recv_count:    " in text
    assert "sy:\n  rank0:\n    ranksync_total_count:   199\n    ranksync_total_data:    640 B\n    ranksync_total_events:  40\n" in text
    assert "    ranksync_avg_time:      7.537 us\n" in text and "    ranksync_total_time:    1.5 ms\n" in text
    with pytest.raises(config.ModelError, match="not available on the device engine"):
        profile.parse_enable_profiling("x:sst.profile.handler.event.bogus[event]")

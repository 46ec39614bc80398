"""Host-side logic on CPU: time parsing, the `sst` model surface, wire-up, partitioning, statistic output formats,
and that the C-ABI library loads and exports every symbol include/sstb200.h declares (no compute without a GPU)."""
import ctypes
import json
import re
from pathlib import Path

import numpy as np
import pytest

from sst_core_b200 import config, models, output, wireup
from sst_core_b200.timelord import TimeError, TimeLord

ROOT = Path(__file__).resolve().parent.parent
MODELS = ROOT / "tests" / "models"
GOLD = json.loads((ROOT / "tests" / "golden" / "reference_golden.json").read_text())


def test_timelord_cycles():
    t = TimeLord()
    assert t.cycles("1ns") == 1000 and t.cycles("10us") == 10_000_000 and t.cycles("1 ps") == 1
    assert [t.cycles(f) for f in ("1GHz", "500MHz", "250MHz", "2GHz")] == [1000, 2000, 4000, 500]
    assert t.cycles("3GHz") == 333 and t.cycles("10000s") == 10**16
    assert t.best_si(10_000_000) == "10 us" and t.best_si(100_000_000) == "100 us"
    with pytest.raises(TimeError):
        t.cycles("1fs")
    with pytest.raises(TimeError):
        t.cycles("12 parsecs")


def _mesh_script(tmp_path):
    # our own copy of the model (same API calls as the reference's tests/test_MessageMesh.py)
    p = tmp_path / "mesh.py"
    p.write_text('''
import sst, sys
sst.setProgramOption("stop-at", "10us")
x_size, y_size = int(sys.argv[1]), int(sys.argv[2])
mod = int(sys.argv[3]) if len(sys.argv) > 3 else 1
links = {}
def getLink(a, b):
    name = "link_%s_%s" % (a, b)
    if name not in links:
        links[name] = sst.Link(name)
    return links[name]
for i in range(x_size * y_size):
    mx, my = i % x_size, i // x_size
    comp = sst.Component("component%d" % i, "coreTestElement.message_mesh.enclosing_component")
    comp.addParam("id", i); comp.addParam("mod", mod); comp.addParam("verbose", True); comp.addParam("stats", 0)
    pxp = comp.setSubComponent("ports", "coreTestElement.message_mesh.message_port", 0)
    pxn = comp.setSubComponent("ports", "coreTestElement.message_mesh.message_port", 1)
    pyp = comp.setSubComponent("ports", "coreTestElement.message_mesh.port_slot", 2).setSubComponent(
        "port", "coreTestElement.message_mesh.message_port")
    pyn = comp.setSubComponent("ports", "coreTestElement.message_mesh.port_slot", 3).setSubComponent(
        "port", "coreTestElement.message_mesh.message_port")
    comp.setSubComponent("route", "coreTestElement.message_mesh.route_message")
    tx = (mx + 1) % x_size
    pxp.addLink(getLink("x%dy%d" % (mx, my), "x%dy%d" % (tx, my)), "port0", "1ns")
    if i % 2 == 0:
        getLink("x%dy%d" % (mx, my), "x%dy%d" % (tx, my)).setNoCut()
    tx = (mx - 1) % x_size
    pxn.addLink(getLink("x%dy%d" % (tx, my), "x%dy%d" % (mx, my)), "port0", "1ns")
    ty = (my + 1) % y_size
    pyp.addLink(getLink("x%dy%d" % (mx, my), "x%dy%d" % (mx, ty)), "port0", "1ns")
    ty = (my - 1) % y_size
    pyn.addLink(getLink("x%dy%d" % (mx, ty), "x%dy%d" % (mx, my)), "port0", "1ns")
''')
    return p


@pytest.mark.parametrize("dims", ["6 6", "2 2", "1 2", "5 3 3"])
def test_script_wireup_equals_vectorised_builder(tmp_path, dims):
    g = config.build_graph(str(_mesh_script(tmp_path)), dims)
    w = wireup.wire(g)
    a = [int(v) for v in dims.split()]
    m = models.mesh_model(a[0], a[1], mod=a[2] if len(a) > 2 else 1)
    for f in ("comp_type", "comp_param", "port_off", "peer_port", "latency", "clock_period"):
        assert np.array_equal(getattr(w, f), getattr(m, f)), f
    canon = lambda a: sorted(tuple(sorted(map(int, r))) for r in a)  # noqa: E731  (unordered pairs, any order)
    assert canon(w.nocut_pairs) == canon(m.nocut_pairs)
    assert w.stop_at == m.stop_at == 10_000_000
    assert w.names[1] == "component1"
    assert w.stats[0] == [] and not w.stats_enabled  # nothing enabled: the statistics are NullStatistics
    w1 = wireup.wire(config.build_graph(str(MODELS / "mesh.py"), dims.split()[0] + " " + dims.split()[1] + " 1 1 0 1"))
    assert w1.stats[0][0].owner_name == "component0:route" and w1.stats_enabled


def test_phold_and_clockpop_scripts_equal_builders():
    w = wireup.wire(config.build_graph(str(MODELS / "phold.py"), "5 4 300ns 7 3000"))
    m = models.phold_model(5, 4, stop_at="300ns", init_events=7, delay_mod=3000)
    for f in ("comp_type", "comp_param", "port_off", "peer_port", "latency", "clock_period"):
        assert np.array_equal(getattr(w, f), getattr(m, f)), f
    w = wireup.wire(config.build_graph(str(MODELS / "clockpop.py"), "6 5 2us 10"))
    m = models.clockpop_model(6, 5, stop_at="2us", comm_freq=10)
    for f in ("comp_type", "comp_param", "port_off", "peer_port", "latency", "clock_period"):
        assert np.array_equal(getattr(w, f), getattr(m, f)), f


def test_model_errors_mirror_reference(tmp_path):
    p = tmp_path / "bad.py"
    p.write_text('import sst\nc = sst.Component("c0", "coreTestElement.message_mesh.enclosing_component")\n')
    with pytest.raises(config.ModelError, match="Must specify param 'id'"):
        wireup.wire(config.build_graph(str(p)))
    p.write_text('import sst\nc = sst.Component("c0", "nolib.nothing")\n')
    with pytest.raises(config.ModelError, match="can't find requested component"):
        wireup.wire(config.build_graph(str(p)))
    p.write_text('import sst\nsst.Component("c0", "pdes.phold")\nsst.Component("c0", "pdes.phold")\n')
    with pytest.raises(config.ModelError, match="already exists"):
        config.build_graph(str(p))
    # a link used three times (the reference: "referenced more than two times")
    with pytest.raises(config.ModelError, match="two endpoints"):
        config.build_graph(str(_mesh_script(tmp_path)), "1 1")


def test_command_line_stop_at_overrides_script():
    g = config.build_graph(str(MODELS / "phold.py"), "2 2 1us")
    assert wireup.wire(g).stop_at == 1_000_000
    assert wireup.wire(g, stop_at_override="50ns").stop_at == 50_000


def test_linear_partition_matches_reference_rule():
    from sst_core_b200 import engine
    # 36 components, even ids glued to their x+ neighbour: 18 groups over 4 partitions -> 5,5,4,4 groups
    m = models.mesh_model(6, 6)
    r = engine.partition_linear(m.ncomp, m.nocut_pairs, 4)
    assert np.bincount(r).tolist() == [10, 10, 8, 8]
    assert all(r[i] == r[i + 1] for i in range(0, 36, 2))
    # fewer groups than partitions: each group its own partition
    assert engine.partition_linear(3, None, 8).tolist() == [0, 1, 2]
    # no no-cut links, 10 over 4: 3,3,2,2
    assert np.bincount(engine.partition_linear(10, None, 4)).tolist() == [3, 3, 2, 2]


def test_partition_16x16_n4_equals_reference_rank_assignment():
    # the reference run with -n 4 printed identical per-component counts (fixture mesh_16x16_200ns_n4); here check the
    # structure the linear rule gives: 128 groups of 2 -> 32 groups = 64 components = 4 rows per partition
    from sst_core_b200 import engine
    m = models.mesh_model(16, 16)
    r = engine.partition_linear(m.ncomp, m.nocut_pairs, 4)
    assert np.array_equal(r, np.repeat(np.arange(4), 64))


def test_c_abi_exports_every_declared_symbol():
    from sst_core_b200 import engine
    header = (ROOT / "include" / "sstb200.h").read_text()
    declared = sorted(set(re.findall(r"\b(sstb200_[a-z_0-9]+)\s*\(", header)))
    assert declared == sorted(engine.EXPORTS)
    L = ctypes.CDLL(str(engine.LIB_PATH))
    for name in declared:
        assert hasattr(L, name), name
    assert L.sstb200_abi_version() == 2
    from sst_core_b200 import checkpoint
    assert checkpoint._ABI == L.sstb200_abi_version()  # the repartitioning code parses blobs of exactly this ABI


def test_native_registry_matches_host_descriptors():
    from sst_core_b200 import elements, engine
    engine.check_registry()
    names = {n for n, *_ in engine.component_types()}
    assert names == set(elements.ELEMENTS)


def test_no_cpu_fallback_without_gpu():
    import torch
    from sst_core_b200 import engine
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(engine.EngineError, match="no CUDA device"):
        engine.Engine(models.mesh_model(2, 2))
    with pytest.raises(engine.EngineError, match="no CUDA device"):
        engine.rng_ticks(0, 1447, 0, 0, 1)


def test_statistic_csv_format_matches_reference():
    import oracle_lib as O
    for key, m in [
        ("phold_3x3_100ns_csv", models.phold_model(3, 3, stop_at="100ns", stats_enabled=True)),
        ("clockpop_3x3_1us_csv", models.clockpop_model(3, 3, stop_at="1us", comm_freq=10, stats_enabled=True)),
    ]:
        script = "phold.py" if "phold" in key else "clockpop.py"
        opts = "3 3 100ns 16 8000 1 1" if "phold" in key else "3 3 1us 10 1 1"
        w = wireup.wire(config.build_graph(str(MODELS / script), opts))
        assert w.stats_enabled
        r = O.run_model(w)
        assert output.stat_rows_csv(w, r.results, r.end_time) == GOLD[key]["rows"]


def test_mesh_csv_and_finish_lines_match_reference(tmp_path):
    import oracle_lib as O
    p = _mesh_script(tmp_path)
    txt = p.read_text().replace('comp.addParam("stats", 0)', 'comp.addParam("stats", 2)')
    txt += '\nsst.setStatisticOutput("sst.statOutputCSV")\nsst.enableAllStatisticsForAllComponents()\n'
    p.write_text(txt)
    w = wireup.wire(config.build_graph(str(p), "3 3"), stop_at_override="100ns")
    r = O.run_model(w)
    assert output.stat_rows_csv(w, r.results, r.end_time) == GOLD["mesh_3x3_100ns_csv"]["rows"]
    lines = output.finish_lines(w, r.results, r.end_time)
    assert lines[-1] == "Simulation is complete, simulated time: 100 ns"
    assert lines[0].startswith("0 received ")


def test_json_interchange_graph_equals_script_graph():
    # a graph written by the reference (--output-json) loads into the same wired model as the script that made it
    wj = wireup.wire(config.load_json(str(ROOT / "tests/golden/mesh_3x2_stats.json")))
    ws = wireup.wire(config.build_graph(str(MODELS / "mesh.py"), "3 2 1 1 1 1"), stop_at_override="40ns")
    for f in ("comp_type", "comp_param", "port_off", "peer_port", "latency", "clock_period"):
        assert np.array_equal(getattr(wj, f), getattr(ws, f)), f
    assert wj.stop_at == ws.stop_at == 40_000 and wj.stats_enabled and wj.names == ws.names
    assert wj.stat_output[0] == "sst.statOutputCSV"


def test_json_model_through_oracle_matches_reference_run_of_the_same_json():
    import oracle_lib as O
    w = wireup.wire(config.load_json(str(ROOT / "tests/golden/mesh_3x2_stats.json")))
    r = O.run_model(w)
    assert {str(i): int(r.results[i, 0]) for i in range(6)} == GOLD["mesh_3x2_json_run"]["counts"]
    assert output.stat_rows_csv(w, r.results, r.end_time) == GOLD["mesh_3x2_json_run"]["rows"]


def test_console_statistic_output_matches_reference():
    import oracle_lib as O
    w = wireup.wire(config.build_graph(str(MODELS / "mesh.py"), "2 2 1 1 1 1"), stop_at_override="30ns")
    r = O.run_model(w)
    assert output.stat_lines_console(w, r.results) == GOLD["mesh_2x2_console_stats"]["lines"]


def test_info_lists_every_registered_element(capsys):
    """sst-info analogue: the native element table joined with the host halves (needs no GPU)"""
    from sst_core_b200 import info
    from sst_core_b200.elements import ELEMENTS
    assert info.main([]) == 0
    out = capsys.readouterr().out
    for name in ELEMENTS:
        assert f"ELEMENT {name} " in out
    assert set(info.DOCS) == set(ELEMENTS)
    assert info.main(["pdes.phold"]) == 0
    assert capsys.readouterr().out.count("ELEMENT ") == 1


def test_sst_module_statistic_enabling_by_name_and_type(tmp_path):
    """the by-name / by-type variants of the reference's statistics API (pymodel.cc function table)"""
    from sst_core_b200 import config, wireup
    src = ('import sst\n'
           'a = sst.Component("a", "pdes.phold"); a.addParams({"id": 0})\n'
           'b = sst.Component("b", "pdes.phold"); b.addParams({"id": 1})\n'
           'c = sst.Component("c", "pdes.clocknode"); c.addParams({"id": 2})\n'
           'sst.Link("l0").connect((a, "port0", "2ns"), (b, "port0", "2ns"))\n'
           'sst.Link("l1").connect((b, "port1", "2ns"), (c, "port0", "2ns"))\n'
           'sst.setProgramOption("stop-at", "50ns")\n'
           'sst.enableStatisticForComponentName("a", "delay", {"rate": "10ns"})\n'
           'sst.enableStatisticsForComponentType("pdes.clocknode", ["N", "S"])\n'
           'sst.setStatisticLoadLevelForComponentName("b", 4)\n'
           'sst.setStatisticLoadLevelForComponentType("pdes.phold", 3)\n'
           'assert sst.getElapsedExecutionTime().endswith(" s") and sst.getLocalMemoryUsage().endswith(" KB")\n'
           'sst.setCallPythonFinalize(True)\n')
    p = tmp_path / "m.py"
    p.write_text(src)
    g = config.build_graph(str(p))
    by = {c.name: c for c in g.components}
    assert by["a"].stats_enabled == ["delay"] and by["a"].stat_params == {"rate": "10ns"}
    assert by["b"].stats_enabled is None and by["c"].stats_enabled == ["N", "S"]
    assert (by["a"].stat_load_level, by["b"].stat_load_level, by["c"].stat_load_level) == (3, 3, None)
    m = wireup.wire(g)
    assert m.stats_enabled and m.stat_rate == 10000
    with pytest.raises(config.ModelError, match="component name zz not found"):
        (tmp_path / "n.py").write_text(src + 'sst.enableStatisticForComponentName("zz", "delay")\n')
        config.build_graph(str(tmp_path / "n.py"))

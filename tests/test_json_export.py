"""Graph interchange towards stock SST: a model built through this package's `sst` surface is
written in the reference's JSON format (config.dump_json / --output-json) and run by the UNMODIFIED reference binary;
its output must be what the reference prints for the script itself.  And the round trip dump -> load gives the same
wired arrays.  CPU tier (needs oracle/_ref for the reference runs)."""
import json
from pathlib import Path

import numpy as np
import pytest

import oracle_lib as O
from sst_core_b200 import config, run, wireup

ROOT = Path(__file__).resolve().parent.parent
GOLD = json.loads((ROOT / "tests" / "golden" / "reference_golden.json").read_text())
MODELS = ROOT / "tests" / "models"


@pytest.mark.parametrize("script,opts", [("mesh.py", "3 2 1 1 1 1"), ("phold.py", "3 3 100ns 16 8000 1 1"),
                                          ("clockpop.py", "3 3 1us 10 1 1"), ("component.py", "4 3 37 2us 1")])
def test_dump_then_load_gives_the_same_wired_model(tmp_path, script, opts):
    g = config.build_graph(str(MODELS / script), opts)
    config.dump_json(g, str(tmp_path / "g.json"))
    a, b = wireup.wire(g), wireup.wire(config.load_json(str(tmp_path / "g.json")))
    for f in ("comp_type", "comp_param", "port_off", "peer_port", "latency", "clock_period", "nocut_pairs"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f
    assert (a.stop_at, a.stats_enabled, a.names, a.stat_output) == (b.stop_at, b.stats_enabled, b.names, b.stat_output)
    names = lambda m: [[(s.owner_name, s.name, s.subid) for s in sl] for sl in m.stats]
    assert names(a) == names(b)


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref is not built")
def test_the_reference_binary_runs_our_json(tmp_path):
    # MessageMesh: the fixture is the reference running the JSON it wrote itself for the same model
    assert run.main(["--run-mode", "init", "--stop-at", "40ns", "--output-json", "ours.json", "--output-directory",
                     str(tmp_path), "--model-options", "3 2 1 1 1 1", str(MODELS / "mesh.py")]) == 0
    out = O.run_ref(tmp_path / "ours.json", cwd=tmp_path)
    got = {int(l.split()[0]): int(l.split()[2]) for l in out.splitlines() if " received " in l}
    assert got == {int(k): v for k, v in GOLD["mesh_3x2_json_run"]["counts"].items()}
    assert (tmp_path / "StatisticOutput.csv").read_text().strip().splitlines() == GOLD["mesh_3x2_json_run"]["rows"]
    # PHOLD (our probe element, loaded by the reference from libpdes.so): statistics CSV of the script run
    g = config.build_graph(str(MODELS / "phold.py"), "3 3 100ns 16 8000 1 1")
    config.dump_json(g, str(tmp_path / "phold.json"))
    O.run_ref(tmp_path / "phold.json", cwd=tmp_path)
    assert (tmp_path / "StatisticOutput.csv").read_text().strip().splitlines() == GOLD["phold_3x3_100ns_csv"]["rows"]
    # coreTestComponent
    g = config.build_graph(str(MODELS / "component.py"), "10 10 37 3us 1")
    config.dump_json(g, str(tmp_path / "comp.json"))
    O.run_ref(tmp_path / "comp.json", cwd=tmp_path)
    assert (tmp_path / "StatisticOutput.csv").read_text().strip().splitlines() == GOLD["component_stats_csv"]["rows"]

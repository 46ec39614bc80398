"""coreTestElement.StatisticsComponent.{int,float} under the statistics engine:
output rates in time and in events, startat / stopat, resetOnOutput, a statistic registered at run time, shared
statistic objects, statistic groups with console / CSV / TXT outputs, load levels and every way of enabling statistics
from the model script (statengine.cc:107-204,346-505; baseComponent.h:756-803; configComponent.cc:285-330;
configGraph.cc:222-233).  Golden: tests/refFiles/test_StatisticsComponent_basic.out and the two group files of the
reference's own test (tests/golden/make_golden_stats.py); the reference's testsuite compares the console output sorted.
CPU tier: the oracle; GPU tier: the engine."""
import json
from pathlib import Path

import numpy as np
import pytest

import oracle_lib as O
from sst_core_b200 import config, output, wireup

ROOT = Path(__file__).resolve().parent.parent
GOLD = json.loads((ROOT / "tests" / "golden" / "stats_golden.json").read_text())
MODEL = ROOT / "tests" / "models" / "statistics_basic.py"
REF_SCRIPT = Path("/root/reference/tests/test_StatisticsComponent_basic.py")


def graph_signature(g):
    comps = [(c.name, c.type, c.params, {k: (v.name, v.shared, v.params) for k, v in c.stat_cfg.items()},
              c.all_stat_params, c.stat_load_level) for c in g.components]
    groups = [(x.name, [c.name for c in x.components], x.stat_map, x.frequency, (x.output.type, x.output.params))
              for x in g.stat_groups]
    links = sorted(tuple(sorted((o.name, p, lat) for o, p, lat in l.ends)) for l in g.links)
    return comps, groups, links, g.program_options, g.stat_output, g.stat_load_level


def oracle_records(run):
    o = run.output
    rec = np.zeros((len(o["time"]), 8), dtype=np.uint64)
    rec[:, 0], rec[:, 2], rec[:, 3], rec[:, 4:7] = o["time"], o["a"], o["b"], o["cde"]
    rec[:, 1] = o["comp"].astype(np.uint64) | (o["code"].astype(np.uint64) << np.uint64(32))
    return rec


def all_output(m, records, results, end_time):
    console, files = output.stat_engine_outputs(m, records, end_time)
    lines = [l for c in m.construct_console for l in c] + console + output.finish_lines(m, results, end_time)
    return lines, files


@pytest.mark.skipif(not REF_SCRIPT.exists(), reason="the reference tree is only present in the build container")
def test_model_script_builds_the_same_graph_as_the_reference_script():
    assert graph_signature(config.build_graph(str(MODEL))) == graph_signature(config.build_graph(str(REF_SCRIPT)))


def test_which_statistics_exist_follows_the_reference_rules():
    m = wireup.wire(config.build_graph(str(MODEL)))
    have = {n: [s.name if s else None for s in sl] for n, sl in zip(m.names, m.stat_slots)}
    assert have["StatGlobal0"] == ["stat1_U32", "stat2_U64", None, None, "stat5_dyn"]  # enable-all at load level 2
    assert have["StatGlobal1"] == ["stat1_U32", "stat2_U64", "stat3_I32", "stat4_I64", "stat5_dyn"]  # own level 4
    assert have["StatBasic2"] == ["stat1_U32", "stat2_U64", "stat3_I32", "stat4_I64", None]  # named ones, any level
    assert have["StatBasic10"] == ["stat1_U32", None, None, None, "stat5_dyn"]  # level 1 set by name
    assert have["StatType0"] == ["stat1_F32", None, None] and have["StatType2"] == ["stat1_F32", "stat2_F64", None]
    assert have["StatObjComp1"] == ["stat1_F32" if False else None, "statobj1", "statobj1"]
    assert have["StatGroupObj4"] == ["stat1_U32", "stat2_U64", None, None, None]
    assert int(m.min_latency()) == 1000 and m.xparam_off[-1] == m.xparam.size


def test_oracle_reproduces_the_reference_output_and_group_files():
    m = wireup.wire(config.build_graph(str(MODEL)))
    run = O.run_model(m)
    assert run.end_time == 101000
    lines, files = all_output(m, oracle_records(run), run.results, run.end_time)
    assert sorted(lines) == sorted(GOLD["basic_out"]["lines"])
    for ext in ("csv", "txt"):
        assert files[GOLD["basic_" + ext]["file"]] == GOLD["basic_" + ext]["lines"]  # same order, too


def test_end_of_simulation_records_equal_the_oracles():
    m = wireup.wire(config.build_graph(str(MODEL)))
    run = O.run_model(m)
    rec = oracle_records(run)
    ends = rec[(rec[:, 1] >> np.uint64(40)) & np.uint64(1) == 1]
    mine = output.stat_end_records(m, run.results, run.end_time)
    key = lambda a: sorted(map(tuple, a[:, :7].tolist()))
    assert key(mine) == key(ends) and len(mine) > 50


@pytest.mark.gpu
def test_gpu_engine_reproduces_the_reference_output_and_group_files():
    from sst_core_b200.run import simulate_model
    m = wireup.wire(config.build_graph(str(MODEL)))
    extras = {}
    res, st = simulate_model(m, extras=extras)
    run = O.run_model(m)
    assert st.end_time == run.end_time == 101000 and st.clock_ticks == run.ticks
    assert np.array_equal(res[:, :26], run.results[:, :26])  # every accumulator, float ones bit for bit
    rec = np.concatenate([extras["output"], output.stat_end_records(m, res, st.end_time)])
    lines, files = all_output(m, rec, res, st.end_time)
    assert sorted(lines) == sorted(GOLD["basic_out"]["lines"])
    for ext in ("csv", "txt"):
        assert sorted(files[GOLD["basic_" + ext]["file"]]) == sorted(GOLD["basic_" + ext]["lines"])
    # every statistic's own sequence of outputs equals the oracle's, record for record
    def per_stat(a):
        d = {}
        for r in a:
            d.setdefault((int(r[1]) & 0xFFFFFFFF, (int(r[1]) >> 32) & 0xFF), []).append(tuple(int(v) for v in r[[0, 2, 3, 4, 5, 6]]))
        return d
    orec = oracle_records(run)
    assert per_stat(rec) == per_stat(orec)

"""Clock semantics beyond one handler per component: several clocks and handlers
per component, Clock::Group ordered handlers, activate / deactivate / unregisterClock / reregisterClock and the
re-registration alignment rule (clock.cc:25-155,193-250,305-420), self links (baseComponent.cc:374-378), through
coreTestElement.coreTestClockerComponent and the graph of the reference's tests/test_Clocks.py
(tests/models/clocks.py builds it).  Golden: tests/refFiles/test_Clocks_basic.out and runs of the reference binary on
other chain lengths / link latencies (tests/golden/make_golden_clocks.py).  The reference's own testsuite compares this
output sorted (testsuite_default_Clocks.py:51-56), because the interleaving of components is not part of the contract;
the lines of ONE component are also compared in order here, oracle against engine.
CPU tier: the oracle; GPU tier: the engine."""
import json
from collections import defaultdict
from pathlib import Path

import numpy as np
import pytest

import oracle_lib as O
from sst_core_b200 import config, output, wireup

ROOT = Path(__file__).resolve().parent.parent
GOLD = json.loads((ROOT / "tests" / "golden" / "clocks_golden.json").read_text())
MODEL = ROOT / "tests" / "models" / "clocks.py"
REF_SCRIPT = Path("/root/reference/tests/test_Clocks.py")
CASES = ["clocks_basic", "clocks_ref_run"] + [k for k in GOLD if "opts" in GOLD[k]]


def model_for(key):
    return wireup.wire(config.build_graph(str(MODEL), GOLD[key].get("opts", "")))


def oracle_lines(m, run):
    o = run.output
    rec = np.stack([o["time"], o["comp"].astype(np.uint64) | (o["code"].astype(np.uint64) << np.uint64(32)), o["a"], o["b"]],
                   axis=1)
    return output.console_lines(m, rec, run.end_time), rec


def per_component(rec):
    d = defaultdict(list)
    for r in rec:
        d[int(r[1]) & 0xFFFFFFFF].append((int(r[0]), int(r[1]) >> 32, int(r[2]), int(r[3])))
    return d


@pytest.mark.skipif(not REF_SCRIPT.exists(), reason="the reference tree is only present in the build container")
def test_model_script_builds_the_same_graph_as_the_reference_script():
    ours, ref = config.build_graph(str(MODEL)), config.build_graph(str(REF_SCRIPT))
    assert [(c.name, c.type, c.params) for c in ours.components] == [(c.name, c.type, c.params) for c in ref.components]
    ends = lambda g: sorted(tuple(sorted((o.name, p, lat) for o, p, lat in l.ends)) for l in g.links)
    assert ends(ours) == ends(ref)
    a, b = wireup.wire(ours), wireup.wire(ref)
    for f in ("comp_type", "comp_param", "port_off", "peer_port", "latency", "clock_period", "port_self"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f


def test_wiring_has_the_self_link_last_and_out_of_the_lookahead():
    m = model_for("clocks_basic")
    assert m.ncomp == 8 and list(np.diff(m.port_off)) == [3] * 8
    assert list(m.port_self.reshape(8, 3)[0]) == [0, 0, 1] and int(m.min_latency()) == 6000
    sp = np.nonzero(m.port_self)[0]
    assert np.array_equal(m.peer_port[sp], sp) and (m.latency[sp] == 0).all()


@pytest.mark.parametrize("key", CASES)
def test_oracle_reproduces_the_reference_output(key):
    m = model_for(key)
    run = O.run_model(m)
    lines, _ = oracle_lines(m, run)
    lines += output.finish_lines(m, run.results, run.end_time)
    assert sorted(lines) == sorted(GOLD[key]["lines"])


def test_oracle_serial_order_is_the_reference_order_up_to_the_first_two_active_groups():
    # the serial interleaving of components is not part of the contract (groups are swapped around inside the reference's
    # Clock, clock.cc:289-303), but until two components have ordered handlers running the oracle's order is the
    # reference's, line for line
    m = model_for("clocks_basic")
    lines, _ = oracle_lines(m, O.run_model(m))
    ref = GOLD["clocks_basic"]["lines"]
    first = next(i for i, (a, b) in enumerate(zip(lines, ref)) if a != b)
    assert first > 100 and "ordered clock result" in ref[first]


@pytest.mark.gpu
@pytest.mark.parametrize("key", CASES[1:])
def test_gpu_engine_reproduces_the_reference_output_and_the_oracle_per_component(key):
    from sst_core_b200.run import simulate_model
    m = model_for(key)
    extras = {}
    res, st = simulate_model(m, extras=extras)
    run = O.run_model(m)
    assert st.end_time == run.end_time and st.events_delivered == run.events
    # the engine finishes the lookahead window in which the last component said OKToEndSim -- like a parallel run of the
    # reference, which checks Exit at its sync points (exit.cc:111-132) -- so a master clock due in the rest of that
    # window is still called once (it finds done_ set and unregisters itself, changing nothing)
    assert run.ticks <= st.clock_ticks <= run.ticks + m.ncomp
    assert np.array_equal(res[:, :13], run.results[:, :13])
    lines = output.console_lines(m, extras["output"], st.end_time) + output.finish_lines(m, res, st.end_time)
    assert sorted(lines) == sorted(GOLD[key]["lines"])
    _, orec = oracle_lines(m, run)
    mine = per_component(extras["output"][extras["output"][:, 0] <= st.end_time][:, :4])
    theirs = per_component(orec)
    assert mine == theirs  # every component printed the same lines in the same order


@pytest.mark.gpu
def test_gpu_clocker_chain_in_three_partitions_on_one_gpu():
    from test_gpu_multirank import _check_local  # steps N ranks of one model on one GPU and compares with the oracle
    for peer in (True, False):
        _check_local(model_for("clocks_ref_run"), 3, peer=peer)

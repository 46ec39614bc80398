"""Open element registration (eli/elementinfo.h:426-435, elemLoader.cc:159-177): an element library from
outside the tree -- integration/example_elements -- is compiled into a second libsstb200 through csrc/elements.def, its
host half registered from Python, and a model of it runs on that library.  No file of the engine is edited.
CPU tier: the library builds (nvcc cross-compiles) and its registry lists the element next to the built-in ones;
GPU tier: a ring of relays against a 10-line Python restatement of the element."""
import os
import subprocess
import sys
import textwrap
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
EX = ROOT / "integration" / "example_elements"


@pytest.fixture(scope="module")
def relay_lib(tmp_path_factory):
    from sst_core_b200 import build
    out = ROOT / "integration" / "example_elements" / "libsstb200_relay.so"  # in-tree: travels to the GPU box
    return build.build(extra_elements=(EX / "relay_elements.cuh", EX / "relay_elements.def"), out=out)


def run_py(code, lib):
    env = dict(os.environ, SSTB200_LIB=str(lib), PYTHONPATH=f"{ROOT}:{EX}")
    return subprocess.run([sys.executable, "-c", textwrap.dedent(code)], capture_output=True, text=True, env=env, timeout=600)


def test_library_with_an_outside_element_builds_and_lists_it(relay_lib):
    r = run_py("""
        import relay
        from sst_core_b200 import engine
        relay.register()
        engine.check_registry()
        print(sorted((code, name) for name, code, *_ in engine.component_types()))
    """, relay_lib)
    assert r.returncode == 0, r.stderr
    assert "(16, 'demo.relay')" in r.stdout and "(1, 'coreTestElement.message_mesh.enclosing_component')" in r.stdout
    # without the host half the mismatch is reported, not ignored
    r = run_py("from sst_core_b200 import engine; engine.check_registry()", relay_lib)
    assert r.returncode != 0 and "element registry mismatch" in r.stderr


@pytest.mark.gpu
def test_ring_of_outside_elements_runs_on_the_gpu(relay_lib, tmp_path):
    model = tmp_path / "ring.py"
    model.write_text(textwrap.dedent("""
        import sst
        N = 7
        sst.setProgramOption("stop-at", "100ns")
        c = [sst.Component(f"r{i}", "demo.relay") for i in range(N)]
        c[2].addParams({"start": 1})
        for i in range(N):
            sst.Link(f"l{i}").connect((c[i], "port0", f"{1 + i % 3}ns"), (c[(i + 1) % N], "port1", f"{1 + i % 3}ns"))
    """))
    r = run_py(f"""
        import relay
        relay.register()
        from sst_core_b200 import config, wireup, output
        from sst_core_b200.run import simulate_model
        m = wireup.wire(config.build_graph({str(model)!r}))
        res, st = simulate_model(m)
        print("\\n".join(output.finish_lines(m, res, st.end_time)))
    """, relay_lib)
    assert r.returncode == 0, r.stderr
    # the token leaves r2 on port0 at time 0 and walks the ring: hop k arrives after the latency of link (2 + k) mod 7
    hops, total, t, at, tok = [0] * 7, [0] * 7, 0, 2, 1
    while True:
        t += 1 + at % 3
        at = (at + 1) % 7
        if t >= 100:
            break
        hops[at] += 1
        total[at] += tok
        tok += 1
    want = [f"r{i}: {hops[i]} hops, sum {total[i]}" for i in range(7)] + ["Simulation is complete, simulated time: 100 ns"]
    assert r.stdout.strip().splitlines() == want

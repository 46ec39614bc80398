"""The reference-side binding of the C ABI (integration/sst_b200_partitioner.cc): the UNMODIFIED reference binary loads
it as an ELI partitioner plugin (`--partitioner=b200.linear`), the plugin calls sstb200_partition_linear in
libsstb200.so.  Partition dumps and results must equal those of the reference's own sst.linear, on meshes whose no-cut
groups do not divide evenly among the threads.  Runs wherever oracle/_ref was built (oracle/build_ref.py)."""
import json
from pathlib import Path

import pytest

import oracle_lib as O

ROOT = Path(__file__).resolve().parent.parent
GOLD = json.loads((ROOT / "tests" / "golden" / "reference_golden.json").read_text())
SHIM = ROOT / "oracle" / "_ref" / "lib" / "sst-elements-library" / "libb200.so"
pytestmark = pytest.mark.skipif(not (O.have_ref() and SHIM.exists()), reason="oracle/_ref with libb200.so is not built")


def run(part, opts, threads, tmp_path, stop="200ns"):
    out = O.run_ref(ROOT / "tests/models/mesh.py", opts, threads=threads, cwd=tmp_path,
                    extra=[f"--stop-at={stop}", f"--partitioner={part}", f"--output-partition={part}.txt"])
    return sorted(out.strip().splitlines()), (tmp_path / f"{part}.txt").read_text()


@pytest.mark.parametrize("opts,threads", [("16 16", 4), ("5 3", 4), ("7 3", 3), ("4 4", 7), ("3 1", 8)])
def test_reference_with_the_b200_partitioner_plugin_equals_sst_linear(tmp_path, opts, threads):
    ours, part_ours = run("b200.linear", opts, threads, tmp_path)
    ref, part_ref = run("sst.linear", opts, threads, tmp_path)
    assert part_ours == part_ref
    assert ours == ref


def test_plugin_run_reproduces_the_committed_golden_counts(tmp_path):
    out, _ = run("b200.linear", "16 16", 4, tmp_path)
    want = sorted([f"{i} received {c} messages" for i, c in GOLD["mesh_16x16_200ns_n4"]["counts"].items()]
                  + ["Simulation is complete, simulated time: 200 ns"])
    assert out == want

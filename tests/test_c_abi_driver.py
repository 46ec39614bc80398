"""The C ABI from C: integration/c_abi_example.c is compiled with gcc against include/sstb200.h alone.
CPU tier: the struct layouts the C compiler sees equal the ctypes mirror in sst_core_b200/engine.py (an ABI drift
between header and Python driver would corrupt options silently), and without a GPU the driver fails loudly.
GPU tier: the C driver runs the 6 x 6 MessageMesh and prints the reference's golden counts."""
import ctypes as C
import json
import subprocess
from pathlib import Path

import pytest

from sst_core_b200 import engine

ROOT = Path(__file__).resolve().parent.parent
GOLD = json.loads((ROOT / "tests" / "golden" / "reference_golden.json").read_text())


@pytest.fixture(scope="module")
def driver(tmp_path_factory):
    engine.lib()  # makes sure libsstb200.so is built
    exe = tmp_path_factory.mktemp("cabi") / "mesh_c"
    lib = ROOT / "sst_core_b200"
    subprocess.check_call(["gcc", "-O2", "-Wall", "-Werror", f"-I{ROOT / 'include'}", str(ROOT / "integration/c_abi_example.c"),
                           f"-L{lib}", "-lsstb200", f"-Wl,-rpath,{lib}", "-o", str(exe)])
    return exe


def test_struct_layouts_match_the_ctypes_mirror(driver):
    out = subprocess.run([str(driver), "--layout"], capture_output=True, text=True, check=True).stdout
    for line, cls in zip(out.strip().splitlines(), (engine._Model, engine._Options, engine._Status)):
        tok = line.split()
        assert int(tok[1]) == C.sizeof(cls), tok[0]
        fields = dict(zip(tok[2::2], map(int, tok[3::2])))
        assert set(fields) == {f[0] for f in cls._fields_}, tok[0]
        for name, off in fields.items():
            assert getattr(cls, name).offset == off, (tok[0], name)


def test_c_driver_fails_loudly_without_a_gpu(driver):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([str(driver), "2", "2", "1000"], capture_output=True, text=True)
    assert r.returncode == 1 and "no CUDA device available" in r.stderr and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_c_driver_reproduces_the_reference_golden_file(driver):
    r = subprocess.run([str(driver), "6", "6", "10000000"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    want = sorted(f"{i} received {c} messages" for i, c in GOLD["mesh_6x6_10us"]["counts"].items())
    assert sorted(lines[:-1]) == want
    assert lines[-1].startswith("Simulation is complete, simulated time: 10000000 ps (1439856 events, ")

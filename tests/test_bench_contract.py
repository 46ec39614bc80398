"""bench.py's driver contract where it can be checked without a GPU: the reference arm (CPU) prints ONE JSON line with
the contract's keys, under plain python and under torchrun (rank 0 alone works); our arm fails loudly without a GPU."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

import oracle_lib as O

ROOT = Path(__file__).resolve().parent.parent
KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
        "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "cpu_baseline", "impl"}


def _json_lines(text):
    return [json.loads(l) for l in text.splitlines() if l.startswith("{")]


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref is not built")
def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--cpu-size", "32", "--cpu-ns", "40"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = _json_lines(r.stdout)
    assert len(lines) == 1
    d = lines[0]
    assert KEYS <= set(d) and d["impl"] == "reference" and d["metric"] == "committed events/sec" and d["unit"] == "events/s"
    assert d["higher_is_better"] is True and d["steps"] == 1 and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref is not built")
def test_reference_arm_under_torchrun_prints_once():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                        "127.0.0.1", "--master-port", "29741", str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "1", "--cpu-size", "32", "--cpu-ns", "40"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = _json_lines(r.stdout)
    assert len(lines) == 1 and lines[0]["impl"] == "reference" and lines[0]["n_gpus"] == 2


def test_our_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--x", "8", "--y", "8", "--steps", "1", "--warmup", "3",
                        "--no-cpu-baseline"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0 and not _json_lines(r.stdout)  # no silent CPU fallback, no bench line

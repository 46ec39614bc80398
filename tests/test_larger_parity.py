"""Larger parity sizes against the THREADED reference: digests of per-component results
written by `sstsim.x -n 8` / `-n 4` (tests/golden/make_golden.py).  CPU tier: the oracle; GPU tier: the engine on one
GPU and cut into 3 partitions."""
import hashlib
import json
from pathlib import Path

import numpy as np
import pytest

import oracle_lib as O
from sst_core_b200 import models

ROOT = Path(__file__).resolve().parent.parent
GOLD = json.loads((ROOT / "tests" / "golden" / "reference_golden.json").read_text())


def mesh_digest(res):
    arr = res[:, 0].astype("<u8").astype("<u4").view("<i4")
    return {"sum": int(arr.sum()), "sha256": hashlib.sha256(np.ascontiguousarray(arr).tobytes()).hexdigest()}


def phold_digest(res):
    blob = "\n".join(f"{i} {int(r[0])} {int(r[1]):016x}" for i, r in enumerate(res)).encode()
    return {"components": len(res), "recv_sum": int(res[:, 0].sum()), "sha256": hashlib.sha256(blob).hexdigest()}


def want(key):
    return {k: v for k, v in GOLD[key].items() if k != "source"}


def test_oracle_matches_the_threaded_reference_on_larger_models():
    assert mesh_digest(O.run_model(models.mesh_model(128, 128, stop_at="150ns")).results) == want("mesh_128x128_150ns_n8_digest")
    assert phold_digest(O.run_model(models.phold_model(48, 40, stop_at="250ns")).results) == want("phold_48x40_250ns_n4_digest")


@pytest.mark.gpu
def test_gpu_matches_the_threaded_reference_on_larger_models():
    from sst_core_b200.engine import LocalMultiRankRunner
    from sst_core_b200.run import simulate_model
    res, _ = simulate_model(models.mesh_model(128, 128, stop_at="150ns"))
    assert mesh_digest(res) == want("mesh_128x128_150ns_n8_digest")
    res, _ = simulate_model(models.phold_model(48, 40, stop_at="250ns"), calendar_slots=16)
    assert phold_digest(res) == want("phold_48x40_250ns_n4_digest")
    r = LocalMultiRankRunner(models.phold_model(48, 40, stop_at="250ns"), 3, calendar_slots=16)
    r.setup()
    r.run()
    out = r.results()
    back = np.empty_like(out)
    back[r.pm.orig_id] = out
    r.close()
    assert phold_digest(back) == want("phold_48x40_250ns_n4_digest")

"""Host half of checkpoint / restart (sst_core_b200/checkpoint.py) on CPU: directory layout and registry as in the
reference (checkpointAction.cc:155-240, config.cc:195-250), checkpoint scheduling at window boundaries, fingerprint."""
import json
from types import SimpleNamespace

import numpy as np
import pytest

from sst_core_b200 import checkpoint, models


class FakeEngine:
    """windows of `w` cycles; finishes at `end`"""

    def __init__(self, w, end, now=0):
        self.w, self.end, self.now, self.calls = w, end, now, []

    def window(self):
        return self.w

    def status(self):
        return SimpleNamespace(now=self.now, finished=self.now >= self.end, error=0)

    def raise_on_error(self, st):
        pass

    def run(self, max_windows=0, check_every=64):
        self.calls.append(max_windows)
        self.now = min(self.end, self.now + max_windows * self.w)
        return self.now >= self.end

    def checkpoint(self):
        return np.arange(16, dtype=np.uint8) + (self.now % 200)


def test_checkpoints_fall_on_the_first_window_boundary_at_or_after_each_period(tmp_path):
    m = models.mesh_model(3, 3, stop_at="100ns")
    recipe = {"script": "x.py", "model_options": "3 3", "stop_at": None, "timebase": None}
    w = checkpoint.CheckpointWriter(str(tmp_path), "cp", recipe, m)
    e = FakeEngine(w=1000, end=10500)
    checkpoint.run_with_checkpoints(e, 2500, w, {"calendar_slots": 4, "stream": 123})
    dirs = sorted(p.name for p in (tmp_path / "cp").iterdir())
    assert dirs == ["cp_1_3000", "cp_2_5000", "cp_3_8000", "cp_4_10000"]  # 2500 -> 3000, 5000, 7500 -> 8000, 10000
    reg = checkpoint.load_registry(str(tmp_path / "cp" / "cp_3_8000" / "cp_3_8000.sstcpt"))
    assert reg["sim_time"] == 8000 and reg["checkpoint_id"] == 3 and reg["recipe"] == recipe
    assert reg["engine_opts"] == {"calendar_slots": 4} and reg["partition_files"] == ["cp_3_8000_0_0.bin"]
    assert reg["model_fingerprint"] == checkpoint.model_fingerprint(m)
    assert np.array_equal(checkpoint.read_partition(reg), np.arange(16, dtype=np.uint8) + (8000 % 200))


def test_period_that_is_a_multiple_of_the_window_is_hit_exactly():
    e = FakeEngine(w=1000, end=9000)
    seen = []
    w = SimpleNamespace(write=lambda eng, t, o: seen.append(t))
    checkpoint.run_with_checkpoints(e, 4000, w, {})
    assert seen == [4000, 8000] and e.calls == [4, 4, 4]


def test_fingerprint_depends_on_every_model_array_and_the_run_options():
    a = models.phold_model(4, 4, stop_at="50ns")
    fp = checkpoint.model_fingerprint(a)
    assert fp == checkpoint.model_fingerprint(models.phold_model(4, 4, stop_at="50ns"))
    b = models.phold_model(4, 4, stop_at="50ns")
    b.latency[3] += 1
    c = models.phold_model(4, 4, stop_at="60ns")
    d = models.phold_model(4, 4, stop_at="50ns", stats_enabled=True)
    e = models.phold_model(4, 4, stop_at="50ns", delay_mod=7000)
    assert len({fp, *(checkpoint.model_fingerprint(x) for x in (b, c, d, e))}) == 5


def test_a_foreign_file_is_not_a_registry(tmp_path):
    p = tmp_path / "x.sstcpt"
    p.write_text(json.dumps({"format": "something else"}))
    with pytest.raises(ValueError, match="not a checkpoint registry"):
        checkpoint.load_registry(str(p))

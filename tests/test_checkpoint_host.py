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
    # the state is taken at the first window boundary at or after each multiple of the period (2500 -> 3000, 5000,
    # 7500 -> 8000, 10000); names and the registry's sim_time carry the multiple itself, like the reference's
    # CheckpointAction (checkpointAction.cc:143-152), the boundary is kept as state_time
    assert dirs == ["cp_1_2500", "cp_2_5000", "cp_3_7500", "cp_4_10000"]
    reg = checkpoint.load_registry(str(tmp_path / "cp" / "cp_3_7500" / "cp_3_7500.sstcpt"))
    assert reg["sim_time"] == 7500 and reg["state_time"] == 8000 and reg["checkpoint_id"] == 3 and reg["recipe"] == recipe
    assert reg["engine_opts"] == {"calendar_slots": 4} and reg["partition_files"] == ["cp_3_7500_0_0.bin"]
    assert reg["model_fingerprint"] == checkpoint.model_fingerprint(m)
    assert np.array_equal(checkpoint.read_partition(reg), np.arange(16, dtype=np.uint8) + (8000 % 200))


def test_period_that_is_a_multiple_of_the_window_is_hit_exactly():
    e = FakeEngine(w=1000, end=9000)
    seen = []
    w = SimpleNamespace(write=lambda eng, t, o, nominal_time=None: seen.append(t))
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


# ---- repartitioned restart: the blob transformation on synthetic blobs (no engine needed)
def _synthetic_blob(m, c0, c1, rank, n_ranks, W, cap, ss, as_, rng, has_payload, with_prof, outbox_cap=16):
    from sst_core_b200.checkpoint import _ABI, _CTRL, _build_blob, _XHDR
    n, nb = c1 - c0, max(1, -(-(c1 - c0) // 256))
    po = m.port_off.astype(np.int64)
    ctrl = np.zeros(1, dtype=_CTRL)
    ctrl["k"], ctrl["T"], ctrl["T_end"], ctrl["w"], ctrl["stop_at"] = 7, 7000, 8000, 1000, 50000
    ctrl["events_delivered"], ctrl["events_sent"], ctrl["clock_ticks"] = 100 + rank, 130 + rank, 5 * rank
    ctrl["local"][0, 0], ctrl["local"][0, 1], ctrl["had_primary"], ctrl["windows"] = 30 - 50 * rank, n, 1, 7
    ctrl["report_at"] = 0xFFFFFFFFFFFFFFFF
    ctl = rng.integers(0, 2**32, size=(n, 32), dtype=np.uint64).astype(np.uint8)
    ctl.view("<u8").reshape(-1, 4)[::3, 0] = 0xFFFFFFFFFFFFFFFF  # some components without an active clock
    state = rng.integers(0, 256, size=(n, ss), dtype=np.uint64).astype(np.uint8)
    aux = rng.integers(0, 256, size=(n, as_), dtype=np.uint64).astype(np.uint8)
    counts = rng.integers(0, cap // 3 + 1, size=W * nb).astype("<u4")  # re-cut blocks may gather several bins
    ev = rng.integers(0, 2**40, size=(W * nb, cap, 2), dtype=np.uint64)
    # live records must address ports of components in their bin's block
    for b in range(W * nb):
        blk = b % nb
        lo, hi = po[c0 + blk * 256], po[min(c1, c0 + (blk + 1) * 256)]
        if hi > lo:
            ports = rng.integers(lo, hi, size=cap, dtype=np.uint64)
            ev[b, :, 1] = (ports << np.uint64(32)) | (ev[b, :, 1] & np.uint64(0xFFFFFFFF))
        else:
            counts[b] = 0
    secs = [ctrl.view(np.uint8), ctl, state, aux, counts, ev]
    if has_payload:
        secs.append(rng.integers(0, 2**60, size=(W * nb, cap), dtype=np.uint64))
    if n_ranks > 1:
        xw = _XHDR + 4 * outbox_cap
        secs += [np.zeros(n_ranks * xw, "<u8"), np.zeros(n_ranks * xw, "<u8"), np.zeros(n_ranks * 8, "<i8")]
    if with_prof:
        secs += [rng.integers(0, 1000, size=int(po[c1] - po[c0]), dtype=np.uint64), rng.integers(0, 99, size=n, dtype=np.uint64)]
    geom = {"n_local": n, "nport_local": int(po[c1] - po[c0]), "c0": c0, "P0": int(po[c0]), "W": W, "cap": cap,
            "state_stride": ss, "aux_stride": as_, "has_payload": int(has_payload), "n_ranks": n_ranks, "rank": rank,
            "window": 1000}
    return _build_blob(_ABI, geom, secs)


def _contents(blobs):
    """partition-independent view of a set of rank blobs"""
    from sst_core_b200.checkpoint import _CTRL, _parse_blob
    P = [_parse_blob(b) for b in blobs]
    g = P[0]["geom"]
    per_comp = {k: np.concatenate([p[k] for p in P]) for k in ("ctl", "state", "aux") + (("prof_recv", "prof_clock") if "prof_recv" in P[0] else ())}
    recs = []
    for p in P:
        nb = max(1, -(-p["geom"]["n_local"] // 256))
        cnt = p["bin_count"].view("<u4").reshape(g["W"], nb)
        ev = p["ev"].view("<u8").reshape(g["W"], nb, g["cap"], 2)
        pl = p["ev_payload"].view("<u8").reshape(g["W"], nb, g["cap"]) if g["has_payload"] else None
        for s in range(g["W"]):
            for b in range(nb):
                for i in range(int(cnt[s, b])):
                    recs.append((s, int(ev[s, b, i, 0]), int(ev[s, b, i, 1]), int(pl[s, b, i]) if pl is not None else 0))
    ctrls = [p["ctrl"].view(_CTRL)[0] for p in P]
    sums = {f: sum(int(c[f]) for c in ctrls) for f in ("events_delivered", "events_sent", "clock_ticks")}
    sums["pending"] = sum(int(c["local"][0]) for c in ctrls)
    sums["primary"] = sum(int(c["local"][1]) for c in ctrls)
    sums["clocks"] = sum(int(c["local"][2]) for c in ctrls)
    return per_comp, sorted(recs), sums, [(int(c["k"]), int(c["T"]), int(c["w"]), int(c["stop_at"])) for c in ctrls]


@pytest.mark.parametrize("has_payload,with_prof", [(False, False), (True, True)])
def test_repartitioned_blobs_hold_the_same_state(has_payload, with_prof):
    from sst_core_b200.engine import partition_linear
    rng = np.random.default_rng(5)
    m = models.phold_model(40, 17, stop_at="50ns")  # 680 components: blocks of 256 do not line up with any cut
    def cut(n):
        r = partition_linear(m.ncomp, None, n).astype(np.int64)
        return r, np.searchsorted(r, np.arange(n + 1))
    r3, b3 = cut(3)
    blobs3 = [_synthetic_blob(m, int(b3[i]), int(b3[i + 1]), i, 3, 4, 24, 48, 64, rng, has_payload, with_prof) for i in range(3)]
    # the synthetic clock counts must be what repartition recomputes
    want = _contents(blobs3)
    for n_new in (1, 2, 5):
        rn, _ = cut(n_new)
        got = _contents(checkpoint.repartition_checkpoints(blobs3, m, rn, n_new, outbox_capacity=16))
        for k in want[0]:
            assert np.array_equal(got[0][k], want[0][k]), k
        assert got[1] == want[1]
        assert {k: v for k, v in got[2].items() if k != "clocks"} == {k: v for k, v in want[2].items() if k != "clocks"}
        assert got[2]["clocks"] == int((want[0]["ctl"].view("<u8").reshape(-1, 4)[:, 0] != np.uint64(0xFFFFFFFFFFFFFFFF)).sum())
        assert len(set(got[3])) == 1 and got[3][0] == want[3][0]
    # and back: 3 -> 5 -> 3 gives the per-rank arrays of the original (records possibly in another order inside a bin)
    r5, _ = cut(5)
    back = checkpoint.repartition_checkpoints(checkpoint.repartition_checkpoints(blobs3, m, r5, 5, 16), m, r3, 3, 16)
    a, b = _contents(back), want
    assert all(np.array_equal(a[0][k], b[0][k]) for k in b[0]) and a[1] == b[1]


def test_repartition_refuses_what_it_cannot_do():
    from sst_core_b200.engine import partition_linear
    rng = np.random.default_rng(6)
    m = models.phold_model(20, 15, stop_at="50ns")
    blob = _synthetic_blob(m, 0, m.ncomp, 0, 1, 2, 8, 48, 0, rng, False, False)
    r2 = partition_linear(m.ncomp, None, 2).astype(np.int64)
    with pytest.raises(ValueError, match="contiguous component ranges"):
        checkpoint.repartition_checkpoints([blob], m, r2[::-1].copy(), 2)
    with pytest.raises(ValueError, match="do not cover the model"):
        checkpoint.repartition_checkpoints([blob], models.phold_model(21, 15, stop_at="50ns"), np.zeros(315, np.int64), 1)
    with pytest.raises(ValueError, match="not a checkpoint"):
        checkpoint.repartition_checkpoints([np.zeros(256, np.uint8)], m, r2, 2)
    # bins of the new partition that would not fit are reported, not truncated
    from sst_core_b200.engine import partition_linear as pl
    m2 = models.phold_model(40, 17, stop_at="50ns")
    r3 = pl(m2.ncomp, None, 3).astype(np.int64)
    b3 = np.searchsorted(r3, np.arange(4))
    full = []
    for i in range(3):
        blob = _synthetic_blob(m2, int(b3[i]), int(b3[i + 1]), i, 3, 2, 8, 48, 0, rng, False, False)
        nb = max(1, -(-(int(b3[i + 1]) - int(b3[i])) // 256))
        off = 15 * 8 + sum(8 + ((len(checkpoint._parse_blob(blob)[k]) + 7) & ~7) for k in ("ctrl", "ctl", "state", "aux")) + 8
        blob[off:off + 4 * 2 * nb].view("<u4")[:] = 8  # every bin full
        full.append(blob)
    with pytest.raises(ValueError, match="would overflow"):
        checkpoint.repartition_checkpoints(full, m2, np.zeros(m2.ncomp, np.int64), 1)

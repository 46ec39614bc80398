"""Checkpoint / restart of a run on the device engine, laid out like the reference's checkpoint directories.

Reference: ``--checkpoint-sim-period`` schedules a CheckpointAction every PERIOD of simulated time
(checkpointAction.cc:143-152); each checkpoint is a directory ``<prefix>/<prefix>_<id>_<simtime>/`` holding the registry
``<prefix>_<id>_<simtime>.sstcpt`` and one ``..._<rank>_<thread>.bin`` per partition (checkpointAction.cc:155-240,
config.cc:195-250); ``sst --load-checkpoint <registry>`` restarts from it (simulation.cc:2073-2200) and the run ends
with the same output as an uninterrupted one (testsuite_default_Checkpoint.py).

Here the partition files are the engine's state blobs (``sstb200_engine_checkpoint_save``: element state incl. RNGs
and statistics, calendar, control blocks) and the registry is a small JSON document with the *recipe* of the model
(script, model options, stop time, timebase, engine options) plus a fingerprint of the wired arrays -- the graph is
rebuilt from the script on restart and must hash to the same value.  Checkpoints are taken at lookahead-window
boundaries: the first boundary at or after each multiple of the period.
"""
from __future__ import annotations

import hashlib
import json
import os
import time
from typing import Optional

import numpy as np

from .timelord import TimeLord
from .wireup import WiredModel


def model_fingerprint(m: WiredModel) -> str:
    h = hashlib.sha256()
    for a in (m.comp_type, m.comp_param, m.port_off, m.peer_port, m.latency, m.clock_period):
        h.update(np.ascontiguousarray(a).view(np.uint8).data)
    h.update(f"{m.stop_at}|{m.timebase}|{int(bool(m.stats_enabled))}".encode())
    return h.hexdigest()


class CheckpointWriter:
    """Writes ``<dir>/<prefix>/<prefix>_<id>_<simtime>/{<base>.sstcpt, <base>_<rank>_0.bin}``."""

    def __init__(self, directory: str, prefix: str, recipe: dict, model: WiredModel, rank: int = 0, n_ranks: int = 1):
        self.root = os.path.join(directory, prefix)
        self.prefix, self.recipe, self.rank, self.n_ranks = prefix, dict(recipe), rank, n_ranks
        self.fingerprint = model_fingerprint(model)
        self.timebase = model.timebase
        self.next_id = 1
        self.last = time.process_time()
        if rank == 0:
            os.makedirs(self.root, exist_ok=True)

    def write(self, engine, sim_time: int, engine_opts: dict, nominal_time: Optional[int] = None) -> str:
        """``sim_time``: the window boundary the state was taken at; ``nominal_time``: the multiple of the checkpoint
        period it stands for (CheckpointAction runs exactly there, checkpointAction.cc:143-152) -- directory names and
        the console line use it, so that they agree with the reference's whenever period % window != 0."""
        actual_time = int(sim_time)
        sim_time = int(nominal_time) if nominal_time is not None else actual_time
        base = f"{self.prefix}_{self.next_id}_{sim_time}"
        d = os.path.join(self.root, base)
        os.makedirs(d, exist_ok=True)
        if self.rank == 0:
            now = time.process_time()
            print(f"# Simulation Checkpoint: Simulated Time {TimeLord(self.timebase).best_si(sim_time)} "
                  f"(Real CPU time since last checkpoint {now - self.last:.5f} seconds)")
            self.last = now
        blob = engine.checkpoint()
        blob.tofile(os.path.join(d, f"{base}_{self.rank}_0.bin"))
        reg = os.path.join(d, f"{base}.sstcpt")
        if self.rank == 0:
            with open(reg, "w") as f:
                json.dump({"format": "sst_core_b200 checkpoint registry", "version": 1, "checkpoint_id": self.next_id,
                           "sim_time": int(sim_time), "state_time": actual_time, "ranks": self.n_ranks, "recipe": self.recipe,
                           "engine_opts": {k: v for k, v in engine_opts.items() if k != "stream"},
                           "model_fingerprint": self.fingerprint,
                           "partition_files": [f"{base}_{r}_0.bin" for r in range(self.n_ranks)]}, f, indent=1)
        self.next_id += 1
        return reg


def load_registry(path: str) -> dict:
    with open(path) as f:
        reg = json.load(f)
    if reg.get("format") != "sst_core_b200 checkpoint registry":
        raise ValueError(f"{path} is not a checkpoint registry of this engine")
    reg["directory"] = os.path.dirname(os.path.abspath(path))
    return reg


def read_partition(reg: dict, rank: int = 0) -> np.ndarray:
    return np.fromfile(os.path.join(reg["directory"], reg["partition_files"][rank]), dtype=np.uint8)


def blob_for_rank(reg: dict, model: WiredModel, rank: int, n_ranks: int, outbox_capacity: int = 0) -> np.ndarray:
    """The state blob rank ``rank`` of ``n_ranks`` restores from: its own partition file when the checkpoint was taken
    on the same number of ranks, else the old ranks' blobs re-cut for the new partition (``model`` = the partitioned
    model of the NEW run; both partitions must keep the script's component numbering)."""
    if int(reg["ranks"]) == n_ranks:
        return read_partition(reg, rank)
    if model.orig_id is not None and not np.array_equal(model.orig_id, np.arange(model.ncomp)):
        raise ValueError("this partition renumbers the components: a checkpoint cannot be re-cut onto it")
    new_rank = model.rank if model.rank is not None else np.zeros(model.ncomp, dtype=np.int64)
    blobs = [read_partition(reg, r) for r in range(int(reg["ranks"]))]
    return repartition_checkpoints(blobs, model, new_rank, n_ranks, outbox_capacity or 65536)[rank]


def run_with_checkpoints(engine, period: int, writer: Optional[CheckpointWriter], engine_opts: dict,
                         first_at: Optional[int] = None, runner=None) -> None:
    """Run a set-up (or restored) engine to the end, writing a checkpoint at the first window boundary at or after
    every multiple of ``period`` cycles (CheckpointAction::execute, checkpointAction.cc:143-152).  ``runner``: what
    steps the windows -- the engine itself on one GPU, a DistributedRunner when every rank of a torch.distributed world
    calls this (each rank writes its own partition file, rank 0 the registry)."""
    runner = runner or engine
    st = engine.status()
    w = engine.window()
    nxt = first_at if first_at is not None else (st.now // period + 1) * period
    while True:
        windows = max(1, -(-(nxt - st.now) // w))
        finished = runner.run(max_windows=windows, check_every=min(windows, 64))
        st = engine.status()
        engine.raise_on_error(st)
        if finished or st.finished:
            return
        if st.now >= nxt:
            if writer is not None:
                writer.write(engine, st.now, engine_opts, nominal_time=(st.now // period) * period)
            nxt = (st.now // period + 1) * period


# ------------------------------------------------------------------------------------------------ repartitioned restart
# A checkpoint written on N ranks restarted on M (testsuite_default_Checkpoint.py:175-193 restarts MessageMesh runs on
# other rank/thread counts).  The blob format is that of sstb200_engine_checkpoint_save (sst_core_b200/csrc/engine.cu):
# 15 u64 header words -- magic, ABI version, number of sections, then n_local, nport_local, c0, P0, W, cap,
# state_stride, aux_stride, has_payload, n_ranks, rank, window -- then per section a byte count and the bytes, padded
# to 8: ctrl, ctl, state, aux, bin_count, ev, [ev_payload], [outbox, inbox, gathered], [prof_recv, prof_clock].
# Per-component arrays are re-sliced; calendar records carry GLOBAL destination ports, so each goes to the bin
# (same window slot, block of its destination component) of the rank that owns that component now; counters are summed
# onto rank 0.  Requires both partitions to keep the model's component numbering (sst.linear over contiguous groups).
_MAGIC = 0x4330303242545353
_ABI = 2  # SSTB200_ABI_VERSION of include/sstb200.h: the layout of Ctrl and of the sections below belongs to it
_HDR = 15
_CTRL = np.dtype([("k", "<u8"), ("T", "<u8"), ("T_end", "<u8"), ("stop_at", "<u8"), ("w", "<u8"),
                  ("events_delivered", "<u8"), ("clock_ticks", "<u8"), ("events_sent", "<u8"), ("max_bin_fill", "<u8"),
                  ("max_due", "<u8"), ("end_time", "<u8"), ("done", "<u4"), ("error", "<u4"), ("error_info", "<u8", 4),
                  ("local", "<i8", 8), ("trace_n", "<u8"), ("windows", "<u8"), ("had_primary", "<u4"), ("slot", "<u4"),
                  ("report_at", "<u8"), ("fb_n", "<u4"), ("maint_n", "<u4"), ("out_n", "<u8"),
                  ("remote_events", "<u8"), ("sync_wait_ns", "<u8"), ("tail_done", "<u4"), ("pf_next", "<u4"), ("fb_taken", "<u4"), ("pad_fb", "<u4")])
assert _CTRL.itemsize == 272  # sizeof(Ctrl) in csrc/engine.cu (static_assert there)
_NEVER = np.uint64(0xFFFFFFFFFFFFFFFF)
_BLOCK = 256
_XHDR = 10


def _parse_blob(blob: np.ndarray) -> dict:
    blob = np.ascontiguousarray(blob, dtype=np.uint8)
    h = blob[:_HDR * 8].view("<u8")
    if h[0] != _MAGIC:
        raise ValueError("not a checkpoint of this engine")
    names = ["n_local", "nport_local", "c0", "P0", "W", "cap", "state_stride", "aux_stride", "has_payload", "n_ranks",
             "rank", "window"]
    g = {n: int(v) for n, v in zip(names, h[3:15])}
    secs, pos = [], _HDR * 8
    for _ in range(int(h[2])):
        n = int(blob[pos:pos + 8].view("<u8")[0])
        secs.append(blob[pos + 8:pos + 8 + n])
        pos += 8 + ((n + 7) & ~7)
    order = ["ctrl", "ctl", "state", "aux", "bin_count", "ev"] + (["ev_payload"] if g["has_payload"] else []) + \
            (["outbox", "inbox", "gathered"] if g["n_ranks"] > 1 else [])
    rest = len(secs) - len(order)
    if rest not in (0, 2) or int(h[1]) != _ABI:
        raise ValueError("unexpected checkpoint layout")
    order += ["prof_recv", "prof_clock"] if rest == 2 else []
    return {"abi": int(h[1]), "geom": g, **dict(zip(order, secs))}


def _build_blob(abi: int, geom: dict, sections: list) -> np.ndarray:
    names = ["n_local", "nport_local", "c0", "P0", "W", "cap", "state_stride", "aux_stride", "has_payload", "n_ranks",
             "rank", "window"]
    parts = [np.array([_MAGIC, abi, len(sections)] + [geom[n] for n in names], dtype="<u8").view(np.uint8)]
    for s in sections:
        b = np.ascontiguousarray(s).view(np.uint8).reshape(-1)
        parts.append(np.array([b.size], dtype="<u8").view(np.uint8))
        parts.append(b)
        if b.size % 8:
            parts.append(np.zeros(8 - b.size % 8, dtype=np.uint8))
    return np.concatenate(parts)


def repartition_checkpoints(blobs, model: WiredModel, new_rank_of_component: np.ndarray, n_new: int,
                            outbox_capacity: int = 65536):
    """``blobs``: the old ranks' checkpoints in rank order; ``model``: the wired model in the numbering both partitions
    use; ``new_rank_of_component``: the new owner of every component (non-decreasing).  Returns the ``n_new`` blobs to
    ``Engine.restore`` into engines created for the new partition with the same calendar_slots / bin_capacity /
    handler_counts and ``outbox_capacity``."""
    old = [_parse_blob(b) for b in blobs]
    g0 = old[0]["geom"]
    for i, o in enumerate(old):
        g = o["geom"]
        if g["rank"] != i or g["n_ranks"] != len(old) or any(g[k] != g0[k] for k in
                                                             ("W", "cap", "state_stride", "aux_stride", "has_payload", "window")):
            raise ValueError("the checkpoints are not the rank-ordered parts of one run")
    ncomp = model.ncomp
    if sum(o["geom"]["n_local"] for o in old) != ncomp or old[0]["geom"]["c0"] != 0:
        raise ValueError("the checkpoints do not cover the model")
    new_rank = np.asarray(new_rank_of_component, dtype=np.int64)
    if new_rank.shape != (ncomp,) or (np.diff(new_rank) < 0).any():
        raise ValueError("the new partition must own contiguous component ranges in the model's numbering")
    begins = np.searchsorted(new_rank, np.arange(n_new + 1)).astype(np.int64)
    port_off = model.port_off.astype(np.int64)
    W, cap, ss, as_, has_pl = g0["W"], g0["cap"], g0["state_stride"], g0["aux_stride"], g0["has_payload"]

    ctrl_old = [o["ctrl"].view(_CTRL)[0] for o in old]
    if any(c["error"] for c in ctrl_old):
        raise ValueError("a rank had reported an error when the checkpoint was taken")
    if sum(int(c["local"][1]) for c in ctrl_old) != ncomp:
        raise ValueError("some primary components have already finished: their ranks cannot be told apart any more")
    cat = lambda key: np.concatenate([o[key] for o in old])
    ctl, state, aux = cat("ctl").reshape(ncomp, 32), cat("state").reshape(ncomp, ss), cat("aux").reshape(ncomp, as_)
    prof = "prof_recv" in old[0]
    if prof:
        prof_recv, prof_clock = cat("prof_recv").view("<u8"), cat("prof_clock").view("<u8")

    # every record of every bin of every old rank, with its window slot
    rec_all, pay_all, slot_all = [], [], []
    for o in old:
        nb = max(1, -(-o["geom"]["n_local"] // _BLOCK))
        cnt = o["bin_count"].view("<u4").reshape(W, nb).astype(np.int64)
        ev = o["ev"].view("<u8").reshape(W, nb, cap, 2)
        live = np.arange(cap)[None, None, :] < cnt[:, :, None]
        rec_all.append(ev[live])
        slot_all.append(np.broadcast_to(np.arange(W)[:, None, None], live.shape)[live])
        if has_pl:
            pay_all.append(o["ev_payload"].view("<u8").reshape(W, nb, cap)[live])
    rec, slot = np.concatenate(rec_all), np.concatenate(slot_all)
    pay = np.concatenate(pay_all) if has_pl else None
    dst_comp = np.searchsorted(port_off, (rec[:, 1] >> np.uint64(32)).astype(np.int64), side="right") - 1
    dst_rank = new_rank[dst_comp]

    out = []
    for r in range(n_new):
        c0, c1 = int(begins[r]), int(begins[r + 1])
        n_local, nb = c1 - c0, max(1, -(-(c1 - c0) // _BLOCK))
        sel = np.nonzero(dst_rank == r)[0]
        b = (dst_comp[sel] - c0) // _BLOCK
        key = slot[sel] * nb + b
        order = np.argsort(key, kind="stable")
        sel, key = sel[order], key[order]
        counts = np.bincount(key, minlength=W * nb)
        if counts.size and counts.max() > cap:
            raise ValueError(f"a calendar bin of the new partition would overflow ({counts.max()} > {cap})")
        first = np.concatenate([[0], np.cumsum(counts)[:-1]])
        within = np.arange(sel.size) - first[key]
        ev = np.zeros((W * nb, cap, 2), dtype="<u8")
        ev[key, within] = rec[sel]
        secs = []
        c = np.zeros(1, dtype=_CTRL)[0]
        for f in ("k", "T", "T_end", "stop_at", "w", "end_time", "done", "windows", "had_primary", "slot"):
            c[f] = ctrl_old[0][f]
        c["report_at"] = _NEVER
        c["max_bin_fill"] = max(int(x["max_bin_fill"]) for x in ctrl_old)
        c["max_due"] = max(int(x["max_due"]) for x in ctrl_old)
        if r == 0:
            for f in ("events_delivered", "clock_ticks", "events_sent"):
                c[f] = sum(int(x[f]) for x in ctrl_old)
            c["local"][0] = sum(int(x["local"][0]) for x in ctrl_old)  # pending = sent - delivered, only the sum counts
        ctl_r = np.ascontiguousarray(ctl[c0:c1])
        c["local"][1] = n_local
        c["local"][2] = int((ctl_r.view("<u8").reshape(-1, 4)[:, 0] != _NEVER).sum()) if n_local else 0
        c["local"][4] = max(int(x["local"][4]) for x in ctrl_old)
        secs += [np.array([c], dtype=_CTRL).view(np.uint8), ctl_r, state[c0:c1], aux[c0:c1],
                 counts.astype("<u4"), ev]
        if has_pl:
            pl = np.zeros((W * nb, cap), dtype="<u8")
            pl[key, within] = pay[sel]
            secs.append(pl)
        if n_new > 1:
            xw = _XHDR + 4 * outbox_capacity
            secs += [np.zeros(n_new * xw, dtype="<u8"), np.zeros(n_new * xw, dtype="<u8"), np.zeros(n_new * 8, dtype="<i8")]
        if prof:
            secs += [prof_recv[int(port_off[c0]):int(port_off[c1])], prof_clock[c0:c1]]
        geom = {"n_local": n_local, "nport_local": int(port_off[c1] - port_off[c0]), "c0": c0, "P0": int(port_off[c0]),
                "W": W, "cap": cap, "state_stride": ss, "aux_stride": as_, "has_payload": has_pl, "n_ranks": n_new,
                "rank": r, "window": g0["window"]}
        out.append(_build_blob(old[0]["abi"], geom, secs))
    return out

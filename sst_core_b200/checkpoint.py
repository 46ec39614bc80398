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

    def write(self, engine, sim_time: int, engine_opts: dict) -> str:
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
                           "sim_time": int(sim_time), "ranks": self.n_ranks, "recipe": self.recipe,
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


def run_with_checkpoints(engine, period: int, writer: Optional[CheckpointWriter], engine_opts: dict,
                         first_at: Optional[int] = None) -> None:
    """Run a set-up (or restored) single-rank engine to the end, writing a checkpoint at the first window boundary at
    or after every multiple of ``period`` cycles (CheckpointAction::execute, checkpointAction.cc:143-152)."""
    st = engine.status()
    w = engine.window()
    nxt = first_at if first_at is not None else (st.now // period + 1) * period
    while True:
        windows = max(1, -(-(nxt - st.now) // w))
        finished = engine.run(max_windows=windows, check_every=min(windows, 64))
        st = engine.status()
        engine.raise_on_error(st)
        if finished or st.finished:
            return
        if st.now >= nxt:
            if writer is not None:
                writer.write(engine, st.now, engine_opts)
            nxt = (st.now // period + 1) * period

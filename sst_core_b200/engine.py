"""Host driver of the CUDA engine through the C ABI (include/sstb200.h).

`Engine` is what replaces ``Simulation::run`` (simulation.cc:1072-1193) for a wired model.  One process drives one
GPU.  With ``torch.distributed`` initialised and world_size > 1 the same class steps a partition of the model and
exchanges boundary events over NCCL once per lookahead window (SyncManager::execute, sync/syncManager.cc:547-732).

There is NO CPU fallback: if libsstb200.so is missing this module raises at import; if no CUDA device is visible,
engine creation fails with the library's error text.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from pathlib import Path
from typing import Optional

import numpy as np

from .elements import NRESULT, TYPE_NAMES
from .wireup import WiredModel

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libsstb200.so"


class EngineError(RuntimeError):
    pass


class _Model(C.Structure):
    _fields_ = [("ncomp", C.c_uint32), ("comp_type", C.c_void_p), ("comp_param", C.c_void_p),
                ("port_off", C.c_void_p), ("peer_port", C.c_void_p), ("latency", C.c_void_p),
                ("clock_period", C.c_void_p), ("stop_at", C.c_uint64), ("stats_enabled", C.c_int),
                ("port_self", C.c_void_p), ("xparam_off", C.c_void_p), ("xparam", C.c_void_p)]


class _Options(C.Structure):
    _fields_ = [("device", C.c_int), ("rank", C.c_uint32), ("n_ranks", C.c_uint32), ("comp_begin", C.c_uint32),
                ("comp_end", C.c_uint32), ("rank_comp_begin", C.c_void_p), ("calendar_slots", C.c_uint32),
                ("bin_capacity", C.c_uint32), ("smem_events", C.c_uint32), ("outbox_capacity", C.c_uint32),
                ("window", C.c_uint64), ("trace_capacity", C.c_uint64), ("stream", C.c_void_p),
                ("parallel_max_events", C.c_uint32), ("handler_counts", C.c_uint32),
                ("model_type_mask", C.c_uint32), ("model_uniform_ports", C.c_uint32), ("output_capacity", C.c_uint64)]


class _Status(C.Structure):
    _fields_ = [("now", C.c_uint64), ("end_time", C.c_uint64), ("windows", C.c_uint64),
                ("events_delivered", C.c_uint64), ("clock_ticks", C.c_uint64), ("events_sent", C.c_uint64),
                ("max_bin_fill", C.c_uint64), ("max_due", C.c_uint64), ("finished", C.c_uint32),
                ("error", C.c_uint32), ("error_info", C.c_uint64 * 4), ("launches", C.c_uint64)]


EXPORTS = [
    "sstb200_abi_version", "sstb200_last_error", "sstb200_num_component_types", "sstb200_component_type_info",
    "sstb200_component_type_lookup", "sstb200_partition_linear", "sstb200_engine_create", "sstb200_engine_destroy",
    "sstb200_engine_setup", "sstb200_engine_run", "sstb200_engine_dispatch", "sstb200_engine_ingest",
    "sstb200_engine_advance", "sstb200_engine_exchange_buffers", "sstb200_engine_control_words",
    "sstb200_engine_status", "sstb200_engine_results", "sstb200_engine_trace", "sstb200_engine_handler_counts",
    "sstb200_engine_checkpoint_size", "sstb200_engine_checkpoint_save", "sstb200_engine_checkpoint_load",
    "sstb200_engine_run_to_report", "sstb200_engine_report_begin", "sstb200_engine_report_end",
    "sstb200_rng_streams",
    "sstb200_rng_ticks", "sstb200_rng_distribution",
    "sstb200_engine_peer_export", "sstb200_engine_peer_connect", "sstb200_engine_peer_connect_local",
    "sstb200_engine_stream", "sstb200_engine_form",
    "sstb200_peer_region_bytes", "sstb200_peer_arena_create", "sstb200_peer_arena_open", "sstb200_peer_arena_destroy",
    "sstb200_engine_in_arena", "sstb200_engine_peer_connect_arena", "sstb200_engine_output",
    "sstb200_engine_handler_times", "sstb200_engine_sync_profile", "sstb200_engine_prepare_run",
]

_lib = None


def lib():
    """Load libsstb200.so; raises if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    import os
    path = Path(os.environ["SSTB200_LIB"]) if os.environ.get("SSTB200_LIB") else LIB_PATH  # a build with more elements
    if not path.exists():
        raise EngineError(f"{path} is missing: build it with `python -m sst_core_b200.build` "
                          f"(the CUDA engine is the only execution path)")
    L = C.CDLL(str(path))
    L.sstb200_last_error.restype = C.c_char_p
    L.sstb200_component_type_info.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_uint32),
                                              C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_int)]
    L.sstb200_component_type_lookup.argtypes = [C.c_char_p]
    L.sstb200_partition_linear.argtypes = [C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
    L.sstb200_engine_create.argtypes = [C.POINTER(_Model), C.POINTER(_Options), C.POINTER(C.c_void_p)]
    L.sstb200_engine_destroy.argtypes = [C.c_void_p]
    L.sstb200_engine_destroy.restype = None
    L.sstb200_engine_setup.argtypes = [C.c_void_p]
    L.sstb200_engine_run.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.POINTER(C.c_int)]
    L.sstb200_engine_prepare_run.argtypes = [C.c_void_p, C.c_uint32]
    for f in ("dispatch", "ingest", "advance"):
        getattr(L, f"sstb200_engine_{f}").argtypes = [C.c_void_p]
    L.sstb200_engine_exchange_buffers.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                                  C.POINTER(C.c_uint64)]
    L.sstb200_engine_control_words.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
    L.sstb200_engine_peer_export.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_void_p)]
    L.sstb200_engine_peer_connect.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.sstb200_engine_peer_connect_local.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.sstb200_engine_stream.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    L.sstb200_engine_form.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
    L.sstb200_engine_handler_times.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.sstb200_engine_sync_profile.argtypes = [C.c_void_p, C.c_void_p]
    L.sstb200_engine_output.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.c_void_p]
    L.sstb200_peer_region_bytes.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]
    L.sstb200_peer_region_bytes.restype = C.c_uint64
    L.sstb200_peer_arena_create.argtypes = [C.c_int, C.c_uint64, C.c_void_p]
    L.sstb200_peer_arena_open.argtypes = [C.c_int, C.c_uint32, C.c_uint32, C.c_void_p]
    L.sstb200_peer_arena_destroy.argtypes = [C.c_int]
    L.sstb200_engine_in_arena.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
    L.sstb200_engine_peer_connect_arena.argtypes = [C.c_void_p, C.c_void_p]
    L.sstb200_engine_status.argtypes = [C.c_void_p, C.POINTER(_Status)]
    L.sstb200_engine_results.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
    L.sstb200_engine_trace.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p]
    L.sstb200_engine_handler_counts.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.sstb200_engine_run_to_report.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32, C.POINTER(C.c_int),
                                               C.POINTER(C.c_int)]
    L.sstb200_engine_report_begin.argtypes = [C.c_void_p, C.c_uint64]
    L.sstb200_engine_report_end.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
    L.sstb200_engine_checkpoint_size.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    L.sstb200_engine_checkpoint_save.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
    L.sstb200_engine_checkpoint_load.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
    L.sstb200_rng_streams.argtypes = [C.c_int, C.c_void_p, C.c_int] + [C.c_uint64] * 6 + [C.c_void_p, C.c_void_p]
    L.sstb200_rng_ticks.argtypes = [C.c_int, C.c_int] + [C.c_uint64] * 4 + [C.c_void_p]
    L.sstb200_rng_distribution.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_uint64,
                                           C.c_uint64, C.c_void_p]
    _lib = L
    return L


def last_error() -> str:
    return lib().sstb200_last_error().decode()


def _check(rc: int, what: str):
    if rc != 0:
        raise EngineError(f"{what} failed ({rc}): {last_error()}")


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def component_types():
    """The native ELI table: [(eli name, type code, state bytes, aux bytes, has payload)]."""
    L = lib()
    out = []
    for i in range(L.sstb200_num_component_types()):
        name, code, sb, ab, hp = C.c_char_p(), C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_int()
        _check(L.sstb200_component_type_info(i, C.byref(name), C.byref(code), C.byref(sb), C.byref(ab),
                                             C.byref(hp)), "component_type_info")
        out.append((name.value.decode(), code.value, sb.value, ab.value, bool(hp.value)))
    return out


def check_registry():
    """The Python element descriptors and the native type table must agree (names and codes)."""
    native = {code: name for name, code, *_ in component_types()}
    if native != TYPE_NAMES:
        raise EngineError(f"element registry mismatch: native {native} vs host {TYPE_NAMES} (an element library compiled "
                          f"into the native side needs its host half: elements.register_element)")


def partition_linear(ncomp: int, nocut_pairs: Optional[np.ndarray], n_parts: int) -> np.ndarray:
    """sst.linear over no-cut groups (linpart.cc:30-82, configGraph.cc:764-850) -- native host C++."""
    pairs = np.ascontiguousarray(nocut_pairs if nocut_pairs is not None else np.zeros((0, 2)), dtype=np.uint32)
    out = np.zeros(ncomp, dtype=np.uint32)
    _check(lib().sstb200_partition_linear(ncomp, _p(pairs), pairs.shape[0], n_parts, _p(out)), "partition_linear")
    return out


def partition_model(m: WiredModel, n_ranks: int) -> WiredModel:
    """Assign ranks with the linear partitioner and renumber components so that every rank owns one contiguous id
    range (the engine's layout); per-component port order -- hence delivery order -- is unchanged."""
    rank = partition_linear(m.ncomp, m.nocut_pairs, n_ranks)
    if m.ncomp < 2 or bool(np.all(rank[:-1] <= rank[1:])):  # already contiguous per rank (the usual case): no renumbering
        m.rank, m.n_ranks = rank, n_ranks
        m.orig_id = np.arange(m.ncomp, dtype=np.uint32)
        return _model_facts(m)
    order = np.argsort(rank, kind="stable")  # new id -> old id
    nports = np.diff(m.port_off.astype(np.int64))
    new_nports = nports[order]
    new_port_off = np.zeros(m.ncomp + 1, dtype=np.int64)
    np.cumsum(new_nports, out=new_port_off[1:])
    # old directed port id -> new id
    old_first = m.port_off.astype(np.int64)[order]
    port_map = np.empty(m.nport, dtype=np.int64)
    idx_new = np.arange(m.nport, dtype=np.int64)
    comp_of_new = np.repeat(np.arange(m.ncomp), new_nports)
    old_ids = old_first[comp_of_new] + (idx_new - new_port_off[comp_of_new])
    port_map[old_ids] = idx_new
    peer_old = m.peer_port[old_ids]
    conn = peer_old != 0xFFFFFFFF
    new_peer = np.full(m.nport, 0xFFFFFFFF, dtype=np.uint32)
    new_peer[conn] = port_map[peer_old[conn].astype(np.int64)].astype(np.uint32)
    out = WiredModel(
        comp_type=np.ascontiguousarray(m.comp_type[order]), comp_param=np.ascontiguousarray(m.comp_param[order]),
        port_off=new_port_off.astype(np.uint32), peer_port=new_peer, latency=np.ascontiguousarray(m.latency[old_ids]),
        clock_period=np.ascontiguousarray(m.clock_period[order]), stop_at=m.stop_at, timebase=m.timebase,
        stats_enabled=m.stats_enabled, rank=rank[order], n_ranks=n_ranks, orig_id=order.astype(np.uint32),
        names=[m.names[i] for i in order] if m.names else None,
        stats=[m.stats[i] for i in order] if m.stats else None,
        finish=[m.finish[i] for i in order] if m.finish else None, nocut_pairs=None, stat_output=m.stat_output,
        port_info=[m.port_info[int(i)] for i in old_ids] if m.port_info else None,
        type_names=[m.type_names[i] for i in order] if m.type_names else None, stat_rate=m.stat_rate,
        port_self=np.ascontiguousarray(m.port_self[old_ids]) if m.port_self is not None else None,
        output_fmt=[m.output_fmt[i] for i in order] if m.output_fmt else None)
    out.validate()
    return _model_facts(out)


def _model_facts(m: WiredModel) -> WiredModel:
    """What the partitioner knows about the WHOLE model and every rank's engine needs: the lookahead (minimum link
    latency, ConfigGraph::getMinimumPartitionLatency configGraph.cc:711-730 -- the reference computes it on the graph,
    before any Simulation object exists), the element types present and whether all components have the same number
    of ports.  With these in the options a rank's engine_create scans its own slice of the model only."""
    m._min_latency = m.min_latency()
    m._rank_begins = np.searchsorted(m.rank, np.arange(m.n_ranks + 1)).astype(np.uint32)
    mask = 0
    for t in np.unique(m.comp_type):
        mask |= 1 << int(t)
    m._type_mask = mask
    np_ = np.diff(m.port_off.astype(np.int64))
    m._uniform_ports = 1 if (np_.size == 0 or bool(np.all(np_ == np_[0]))) else 2
    return m


@dataclass
class Status:
    now: int
    end_time: int
    windows: int
    events_delivered: int
    clock_ticks: int
    events_sent: int
    max_bin_fill: int
    max_due: int
    finished: bool
    error: int
    error_info: tuple
    launches: int


_ERRORS = {
    1: "calendar bin overflow (raise bin_capacity or calendar_slots): slot {0}, block {1}, fill {2}",
    2: "shared-memory staging overflow (raise smem_events): block {0} had {1} due events, capacity {2}",
    3: "time fault: event for time {0} while the window starts at {1} (component {2}); a send violated the lookahead",
    4: "outbox overflow (raise outbox_capacity): to rank {0}, fill {1}, capacity {2}",
    5: "trace overflow",
    6: "peer mode: the mailbox entry of rank {0} for window sequence {1} did not arrive (found {2})",
    7: "output log overflow (raise output_capacity): record {0}, capacity {1}, component {2}",
    8: "self-link queue overflow: component {0} has {1} self events pending inside one window (time {2})",
    0x80000000: "another rank reported an error",
}


def _global_window(model: WiredModel, c0: int, c1: int, rank: int, n_ranks: int, device: int) -> int:
    """Lookahead window of a partitioned model = the minimum link latency over ALL ranks.  Inside a torch.distributed
    job of n_ranks processes every rank scans its own ports and one MIN all-reduce gives the result (the reference does
    the same with MPI_Allreduce, sync/syncManager.cc:276-377); otherwise the model's ports are scanned once."""
    try:
        import torch
        import torch.distributed as dist
        in_job = dist.is_available() and dist.is_initialized() and dist.get_world_size() == n_ranks and \
            dist.get_rank() == rank
    except Exception:
        in_job = False
    if not in_job:
        w = getattr(model, "_min_latency", None)
        if w is None:
            w = model._min_latency = model.min_latency()
        return w or 0
    p0, p1 = int(model.port_off[c0]), int(model.port_off[c1])
    lat = model.latency[p0:p1][model.peer_port[p0:p1] != 0xFFFFFFFF]
    local = int(lat.min()) if lat.size else (1 << 62)
    if dist.get_backend() == "nccl":
        t = torch.tensor([local], dtype=torch.int64, device=torch.device("cuda", device))
    else:
        t = torch.tensor([local], dtype=torch.int64)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    w = int(t.item())
    return 0 if w >= (1 << 62) else w


class Engine:
    """One rank's engine.  ``Engine(model).run()`` for a single GPU; see :func:`run_distributed` for N GPUs."""

    def __init__(self, model: WiredModel, device: int = 0, rank: int = 0, n_ranks: int = 1, calendar_slots: int = 2,
                 bin_capacity: int = 0, smem_events: int = 0, outbox_capacity: int = 0, window: int = 0,
                 trace_capacity: int = 0, stream: int = 0, parallel_max_events: int = 0, handler_counts: bool = False,
                 output_capacity: int = 0):
        L = lib()
        check_registry()
        self.model = model
        self.rank, self.n_ranks = rank, n_ranks
        self._keep = [np.ascontiguousarray(a) for a in (model.comp_type, model.comp_param, model.port_off,
                                                        model.peer_port, model.latency, model.clock_period)]
        ps = getattr(model, "port_self", None)
        self._port_self = np.ascontiguousarray(ps, dtype=np.uint8) if ps is not None else None
        cm = _Model(model.ncomp, *[_p(a) for a in self._keep], int(model.stop_at), int(bool(model.stats_enabled)),
                    _p(self._port_self) if self._port_self is not None else None, None, None)
        xo = getattr(model, "xparam_off", None)
        if xo is not None:
            self._xparam = (np.ascontiguousarray(xo, dtype=np.uint32), np.ascontiguousarray(model.xparam, dtype=np.int64))
            cm.xparam_off, cm.xparam = _p(self._xparam[0]), _p(self._xparam[1])
        if n_ranks > 1:
            if model.rank is None or model.n_ranks != n_ranks:
                raise EngineError("multi-rank engine needs a model prepared by partition_model(model, n_ranks)")
            begins = getattr(model, "_rank_begins", None)
            if begins is None or len(begins) != n_ranks + 1:  # once per partitioned model, not once per engine
                begins = np.searchsorted(model.rank, np.arange(n_ranks + 1)).astype(np.uint32)
                model._rank_begins = begins
            self._begins = begins
            c0, c1 = int(begins[rank]), int(begins[rank + 1])
            if not window:
                window = getattr(model, "_min_latency", 0) or _global_window(model, c0, c1, rank, n_ranks, device)
        else:
            self._begins = np.array([0, model.ncomp], dtype=np.uint32)
            c0, c1 = 0, model.ncomp
        self.comp_begin, self.comp_end = c0, c1
        opt = _Options(device, rank, n_ranks, c0, c1, _p(self._begins), calendar_slots, bin_capacity, smem_events,
                       outbox_capacity, window, trace_capacity, stream, parallel_max_events, int(handler_counts),
                       int(getattr(model, "_type_mask", 0)) if n_ranks > 1 else 0,
                       int(getattr(model, "_uniform_ports", 0)) if n_ranks > 1 else 0, output_capacity)
        h = C.c_void_p()
        _check(L.sstb200_engine_create(C.byref(cm), C.byref(opt), C.byref(h)), "engine_create")
        self._h = h
        self._L = L
        self.device = device
        self._window = window
        self.peer = None  # "ipc" / "local" once the ranks are connected in peer mode

    def close(self):
        if getattr(self, "_h", None):
            self._L.sstb200_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def setup(self):
        _check(self._L.sstb200_engine_setup(self._h), "engine_setup")

    def status(self) -> Status:
        s = _Status()
        _check(self._L.sstb200_engine_status(self._h, C.byref(s)), "engine_status")
        return Status(s.now, s.end_time, s.windows, s.events_delivered, s.clock_ticks, s.events_sent,
                      s.max_bin_fill, s.max_due, bool(s.finished), s.error, tuple(s.error_info), s.launches)

    def raise_on_error(self, st: Optional[Status] = None):
        st = st or self.status()
        if st.error:
            msg = _ERRORS.get(st.error, f"error code {st.error}").format(*st.error_info)
            raise EngineError(f"engine error at window {st.error_info[3]}: {msg}")

    def run(self, max_windows: int = 0, check_every: int = 64) -> bool:
        """Run to the end (or for max_windows lookahead windows).  Returns True when the simulation has ended."""
        fin = C.c_int(0)
        rc = self._L.sstb200_engine_run(self._h, max_windows, check_every, C.byref(fin))
        if rc == -4:
            self.raise_on_error()
        _check(rc, "engine_run")
        return bool(fin.value)

    def prepare_run(self, check_every: int = 64):
        """Capture the CUDA graph of a chunk of `check_every` windows ahead of the first run() with that chunk size."""
        _check(self._L.sstb200_engine_prepare_run(self._h, check_every), "engine_prepare_run")

    def dispatch(self):
        _check(self._L.sstb200_engine_dispatch(self._h), "engine_dispatch")

    def ingest(self):
        _check(self._L.sstb200_engine_ingest(self._h), "engine_ingest")

    def advance(self):
        _check(self._L.sstb200_engine_advance(self._h), "engine_advance")

    # ---- peer mode (include/sstb200.h): boundary events stored straight into the owning rank's calendar over NVLink
    def peer_export(self):
        """(64-byte CUDA IPC handle, number of component blocks, device pointer) of this rank's peer-visible region."""
        h = (C.c_uint8 * 64)()
        nb, reg = C.c_uint32(0), C.c_void_p()
        _check(self._L.sstb200_engine_peer_export(self._h, h, C.byref(nb), C.byref(reg)), "engine_peer_export")
        return bytes(h), int(nb.value), reg.value

    def peer_connect(self, handles, nbins):
        """handles: n_ranks x 64 bytes, nbins: n_ranks block counts, both in rank order (from every rank's peer_export).
        Call on every rank before setup()."""
        hb = np.frombuffer(b"".join(handles), dtype=np.uint8).copy()
        nb = np.ascontiguousarray(nbins, dtype=np.uint32)
        _check(self._L.sstb200_engine_peer_connect(self._h, _p(hb), _p(nb)), "engine_peer_connect")
        self.peer = "ipc"

    def in_arena(self) -> bool:
        v = C.c_int(0)
        _check(self._L.sstb200_engine_in_arena(self._h, C.byref(v)), "engine_in_arena")
        return bool(v.value)

    def peer_connect_arena(self, nbins):
        """Every rank's peer-visible region lies in its (already mapped) peer arena: connect without mapping anything."""
        nb = np.ascontiguousarray(nbins, dtype=np.uint32)
        _check(self._L.sstb200_engine_peer_connect_arena(self._h, _p(nb)), "engine_peer_connect_arena")
        self.peer = "ipc"

    def peer_connect_local(self, regions, nbins):
        """The same for ranks of one process on one GPU: regions = their device pointers."""
        rg = (C.c_void_p * len(regions))(*regions)
        nb = np.ascontiguousarray(nbins, dtype=np.uint32)
        _check(self._L.sstb200_engine_peer_connect_local(self._h, rg, _p(nb)), "engine_peer_connect_local")
        self.peer = "local"

    def handler_times(self):
        """(ns inside the event handlers per port, ns inside the clock handlers per component) of this rank; the engine
        must have been created with handler_counts=3."""
        recv = np.zeros(int(self.model.port_off[self.comp_end]) - int(self.model.port_off[self.comp_begin]), dtype=np.float64)
        clk = np.zeros(self.comp_end - self.comp_begin, dtype=np.float64)
        _check(self._L.sstb200_engine_handler_times(self._h, _p(recv), _p(clk)), "engine_handler_times")
        return recv, clk

    def sync_profile(self):
        """(rank syncs, events sent to other ranks, their bytes, ns waited at the window barrier) of this rank"""
        out = np.zeros(4, dtype=np.uint64)
        _check(self._L.sstb200_engine_sync_profile(self._h, _p(out)), "engine_sync_profile")
        return tuple(int(v) for v in out)

    def output_records(self) -> np.ndarray:
        """Console-output records of this rank's device components: u64 [n, 8] = {time, component | code << 32,
        v0..v4, 0}; a component's records appear in the order it printed them."""
        n = C.c_uint64(0)
        _check(self._L.sstb200_engine_output(self._h, C.byref(n), None), "engine_output")
        out = np.zeros((int(n.value), 8), dtype=np.uint64)
        if n.value:
            _check(self._L.sstb200_engine_output(self._h, C.byref(n), _p(out)), "engine_output")
        return out[:int(n.value)]

    FORMS = ("sequential", "event-parallel", "ordered", "clock")

    def form(self) -> str:
        """Which window kernel runs this rank's windows: one of ``Engine.FORMS``."""
        f = C.c_int()
        _check(self._L.sstb200_engine_form(self._h, C.byref(f)), "engine_form")
        return self.FORMS[f.value]

    def stream_handle(self) -> int:
        st = C.c_void_p()
        _check(self._L.sstb200_engine_stream(self._h, C.byref(st)), "engine_stream")
        return int(st.value or 0)

    def exchange_buffers(self):
        ob, ib, w = C.c_void_p(), C.c_void_p(), C.c_uint64()
        _check(self._L.sstb200_engine_exchange_buffers(self._h, C.byref(ob), C.byref(ib), C.byref(w)), "exchange")
        return ob.value, ib.value, w.value

    def control_words(self):
        a, b = C.c_void_p(), C.c_void_p()
        _check(self._L.sstb200_engine_control_words(self._h, C.byref(a), C.byref(b)), "control_words")
        return a.value, b.value

    def exchange_tensors(self):
        """(outbox, inbox) as zero-copy int64 torch tensors [n_ranks, words] over engine-owned device memory."""
        import torch
        ob, ib, words = self.exchange_buffers()
        dev = torch.device("cuda", self.device)
        n = self.n_ranks
        return (_as_tensor(torch, ob, n * words, dev).view(n, words), _as_tensor(torch, ib, n * words, dev).view(n, words))

    def control_tensors(self):
        """(local [8], gathered [n_ranks*8]) control words as zero-copy int64 torch tensors."""
        import torch
        lw, gw = self.control_words()
        dev = torch.device("cuda", self.device)
        return _as_tensor(torch, lw, 8, dev), _as_tensor(torch, gw, 8 * self.n_ranks, dev)

    def sync(self):
        """Wait for the engine's stream (status() does a device->host read on it)."""
        self.status()

    def results(self, words: int = 0, compact: bool = False, out: Optional[np.ndarray] = None) -> np.ndarray:
        """Result records of this rank's components.  `words` > 0 downloads only the first `words` of the 32 words per
        component; the returned array is [n, 32] (rest zero) unless `compact`, then [n, words] with no host-side copy
        (what a large run wants: 16.7 M components x 32 words is 4.3 GB of mostly zeros).  `out` (compact form only):
        a caller-owned C-contiguous uint64 [n, words] array to download into, e.g. ``wireup.pinned_empty`` memory."""
        n = self.comp_end - self.comp_begin
        if out is not None:
            w = words if words and words < NRESULT else NRESULT
            if out.dtype != np.uint64 or out.shape != (n, w) or not out.flags.c_contiguous:
                raise ValueError(f"results(out=): need a C-contiguous uint64 array of shape ({n}, {w})")
            _check(self._L.sstb200_engine_results(self._h, _p(out), 0 if w == NRESULT else w), "engine_results")
            return out
        if not words or words >= NRESULT:
            out = np.empty((n, NRESULT), dtype=np.uint64)
            _check(self._L.sstb200_engine_results(self._h, _p(out), 0), "engine_results")
            return out
        part = np.empty((n, words), dtype=np.uint64)
        _check(self._L.sstb200_engine_results(self._h, _p(part), words), "engine_results")
        if compact:
            return part
        out = np.zeros((n, NRESULT), dtype=np.uint64)
        out[:, :words] = part
        return out

    def run_to_report(self, report_time: int, words: int = 0):
        """Run up to and including the window that contains ``report_time`` (cycles).  Returns (records, finished):
        records = the result records [n, 32] as of the END of that time (periodic statistic output), or None when
        the simulation ended first or the time is at/after stop-at."""
        n = self.comp_end - self.comp_begin
        w = words if words and words < NRESULT else NRESULT
        part = np.zeros((n, w), dtype=np.uint64)
        rep, fin = C.c_int(0), C.c_int(0)
        rc = self._L.sstb200_engine_run_to_report(self._h, int(report_time), _p(part), 0 if w == NRESULT else w,
                                                  C.byref(rep), C.byref(fin))
        if rc == -4:
            self.raise_on_error()
        _check(rc, "engine_run_to_report")
        if not rep.value:
            return None, bool(fin.value)
        out = np.zeros((n, NRESULT), dtype=np.uint64)
        out[:, :w] = part
        return out, bool(fin.value)

    def report_begin(self, report_time: int):
        """Bracket the next window (dispatch .. advance) that contains ``report_time``: see run_to_report."""
        _check(self._L.sstb200_engine_report_begin(self._h, int(report_time)), "engine_report_begin")

    def report_end(self, words: int = 0) -> np.ndarray:
        n = self.comp_end - self.comp_begin
        w = words if words and words < NRESULT else NRESULT
        part = np.zeros((n, w), dtype=np.uint64)
        _check(self._L.sstb200_engine_report_end(self._h, _p(part), 0 if w == NRESULT else w), "engine_report_end")
        out = np.zeros((n, NRESULT), dtype=np.uint64)
        out[:, :w] = part
        return out

    def window(self) -> int:
        """Lookahead window in cycles (the option, else the model's minimum link latency)."""
        if not getattr(self, "_window", 0):
            self._window = self.model.min_latency() or 1000000
        return self._window

    def checkpoint(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        """This rank's complete simulation state between two windows as a uint8 array (``out``: a caller-owned,
        e.g. pinned, buffer of at least ``checkpoint_size()`` bytes)."""
        n = self.checkpoint_size()
        blob = out if out is not None else np.empty(n, dtype=np.uint8)
        if blob.dtype != np.uint8 or blob.size < n or not blob.flags.c_contiguous:
            raise ValueError(f"checkpoint(out=): need a C-contiguous uint8 array of at least {n} bytes")
        _check(self._L.sstb200_engine_checkpoint_save(self._h, _p(blob), blob.size), "engine_checkpoint_save")
        return blob[:n]

    def checkpoint_size(self) -> int:
        n = C.c_uint64(0)
        _check(self._L.sstb200_engine_checkpoint_size(self._h, C.byref(n)), "engine_checkpoint_size")
        return int(n.value)

    def restore(self, blob: np.ndarray):
        """Load a checkpoint written by an engine with the same model slice and calendar geometry; call it INSTEAD of
        setup().  The run continues bit-identically to the one that wrote the checkpoint."""
        blob = np.ascontiguousarray(blob, dtype=np.uint8)
        _check(self._L.sstb200_engine_checkpoint_load(self._h, _p(blob), blob.size), "engine_checkpoint_load")

    def handler_counts(self):
        """(event_recv [local directed ports], clock_calls [local components]) u64 -- the data of the reference's
        ``sst.profile.handler.event.count`` / ``sst.profile.handler.clock.count`` tools (needs handler_counts=True)."""
        m = self.model
        np_local = int(m.port_off[self.comp_end]) - int(m.port_off[self.comp_begin])
        recv = np.zeros(np_local, dtype=np.uint64)
        clk = np.zeros(self.comp_end - self.comp_begin, dtype=np.uint64)
        _check(self._L.sstb200_engine_handler_counts(self._h, _p(recv), _p(clk)), "engine_handler_counts")
        return recv, clk

    def trace(self):
        n = C.c_uint64(0)
        _check(self._L.sstb200_engine_trace(self._h, C.byref(n), None, None, None, None), "engine_trace")
        k = n.value
        t = dict(time=np.zeros(k, np.uint64), comp=np.zeros(k, np.uint32), port=np.zeros(k, np.uint32),
                 payload=np.zeros(k, np.uint64))
        n = C.c_uint64(k)
        _check(self._L.sstb200_engine_trace(self._h, C.byref(n), _p(t["time"]), _p(t["comp"]), _p(t["port"]),
                                            _p(t["payload"])), "engine_trace")
        return t


# ------------------------------------------------------------------------------------------------ multi-GPU stepping
def connect_peers(engine) -> bool:
    """Put the ranks of a torch.distributed job (NCCL backend, one process per GPU of one node) into PEER MODE: every
    rank maps the others' calendars over CUDA IPC, boundary events are stored straight into the owner's bins over NVLink
    and the per-window exchange is a mailbox write -- no collective in the window loop (include/sstb200.h, "peer
    mode").  Collective: every rank calls it after creating its Engine and BEFORE setup().  Returns False (and changes
    nothing) outside such a job or when SSTB200_NO_PEER is set; the ranks then exchange outboxes with one all-to-all per
    window (DistributedRunner)."""
    import os

    try:
        import torch
        import torch.distributed as dist
    except Exception:
        return False
    n = engine.n_ranks
    if os.environ.get("SSTB200_NO_PEER") or n < 2 or not (dist.is_available() and dist.is_initialized()):
        return False
    if dist.get_backend() != "nccl" or dist.get_world_size() != n or dist.get_rank() != engine.rank:
        return False
    handle, nbins, _ = engine.peer_export()
    dev = torch.device("cuda", engine.device)
    mine = np.concatenate([np.frombuffer(handle, dtype=np.uint8),
                           np.array([nbins, int(engine.in_arena())], dtype=np.uint32).view(np.uint8)])
    allt = torch.empty(n * 72, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(allt, torch.from_numpy(mine.copy()).to(dev))  # also: every rank's region is cleared
    a = allt.cpu().numpy().reshape(n, 72)
    meta = np.ascontiguousarray(a[:, 64:]).view(np.uint32).reshape(n, 2)
    if bool(meta[:, 1].all()):  # every rank's region lies in its peer arena, mapped when the arena was opened
        engine.peer_connect_arena(meta[:, 0].copy())
    else:
        engine.peer_connect([bytes(a[r, :64]) for r in range(n)], meta[:, 0].copy())
    return True


def init_peer_arena(device: int, nbytes: int) -> bool:
    """Collective, once per process (like creating a communicator): allocate this rank's peer arena of `nbytes`, exchange
    the IPC handles and map every other rank's arena.  Engines created afterwards whose peer-visible region fits are
    placed inside it, and connect_peers() then maps nothing (cudaIpcOpenMemHandle costs 10-25 ms per peer and
    allocation).  Returns False outside a torch.distributed NCCL job."""
    try:
        import torch
        import torch.distributed as dist
    except Exception:
        return False
    if not (dist.is_available() and dist.is_initialized()) or dist.get_backend() != "nccl" or dist.get_world_size() < 2:
        return False
    L = lib()
    n, rank = dist.get_world_size(), dist.get_rank()
    h = (C.c_uint8 * 64)()
    _check(L.sstb200_peer_arena_create(device, int(nbytes), h), "peer_arena_create")
    dev = torch.device("cuda", device)
    allt = torch.empty(n * 64, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(allt, torch.from_numpy(np.frombuffer(bytes(h), dtype=np.uint8).copy()).to(dev))
    hb = allt.cpu().numpy().copy()
    _check(L.sstb200_peer_arena_open(device, rank, n, _p(hb)), "peer_arena_open")
    dist.barrier()
    return True


def peer_region_bytes(n_local: int, calendar_slots: int = 0, bin_capacity: int = 0, has_payload: bool = False) -> int:
    return int(lib().sstb200_peer_region_bytes(int(n_local), int(calendar_slots), int(bin_capacity), int(has_payload)))


class DistributedRunner:
    """One lookahead window per step across all ranks of the torch.distributed world.

    Peer mode (the engine was connected with :func:`connect_peers`):  dispatch -> advance.  The boundary events are in
    their owners' calendars when dispatch ends; advance waits on the device for every rank's mailbox entry.  ``run``
    hands whole chunks of windows to the native driver (CUDA-graph replay).
    Outbox mode:  dispatch -> ONE all-to-all of the outboxes, control words piggy-backed in the headers (NCCL) ->
    ingest -> advance.

    Replaces SyncManager::execute + RankSyncSerialSkip::exchange and its MPI_Allreduce calls
    (sync/syncManager.cc:547-732, sync/rankSyncSerialSkip.cc:209-343).  The engine may be the CUDA `Engine` or any
    object with the same stepping interface (the CPU tests drive the oracle through this class over gloo)."""

    def __init__(self, engine):
        import torch.distributed as dist

        self.dist, self.e = dist, engine
        self.peer = getattr(engine, "peer", None) == "ipc"
        self._stream_ctx = None
        if self.peer:
            return
        self.outbox, self.inbox = engine.exchange_tensors()
        self.local, self.gathered = engine.control_tensors()
        if hasattr(engine, "stream_handle"):
            # the all-to-all must run on the ENGINE's stream, between its dispatch and ingest kernels: NCCL orders
            # itself behind torch's current stream only
            import torch
            ext = torch.cuda.ExternalStream(engine.stream_handle(), device=torch.device("cuda", engine.device))
            self._stream_ctx = lambda: torch.cuda.stream(ext)

    def step(self):
        e = self.e
        e.dispatch()  # ... which also publishes this rank's control words (outbox headers / the ranks' mailboxes)
        if self.peer:
            e.advance()
            return
        if self._stream_ctx:
            with self._stream_ctx():
                self.dist.all_to_all_single(self.inbox, self.outbox)  # the ONE collective of a window
        else:
            self.dist.all_to_all_single(self.inbox, self.outbox)
        e.ingest()  # bins received events, collects every rank's control words from the inbox headers
        e.advance()

    def run(self, max_windows: int = 0, check_every: int = 32) -> bool:
        if self.peer:
            return self.e.run(max_windows=max_windows, check_every=check_every)
        done_windows = 0
        while True:
            chunk = check_every if not max_windows else min(check_every, max_windows - done_windows)
            for _ in range(chunk):
                self.step()
            done_windows += chunk
            st = self.e.status()
            self.e.raise_on_error(st)
            if st.finished:
                return True
            if max_windows and done_windows >= max_windows:
                return False

    def run_to_report(self, report_time: int, words: int = 0):
        """Periodic statistic output across the ranks: run up to and including the window that contains
        ``report_time`` and return (this rank's result records as of the END of that time, finished) -- (None, finished)
        when the simulation ends first or the time is at/after stop-at.  Every rank calls it with the same time."""
        e = self.e
        st = e.status()
        if st.finished:
            return None, True
        stop_at = int(e.model.stop_at)
        if stop_at and report_time >= stop_at:
            return None, False
        w = e.window()
        before = report_time // w - st.now // w
        if before < 0:
            raise EngineError("run_to_report: the report time lies before the current window")
        if before and self.run(max_windows=before, check_every=min(before, 32)):
            return None, True
        e.report_begin(report_time)
        self.step()
        rec = e.report_end(words)
        st = e.status()
        e.raise_on_error(st)
        return rec, bool(st.finished)


class LocalMultiRankRunner:
    """All partitions of a model stepped by ONE process on ONE GPU: the same device code path as a multi-GPU run
    (outbox emission, ingest kernel, gathered control words) with the exchange done by device-to-device copies.
    Used to validate the multi-rank path where fewer GPUs than ranks are available, and usable as a way to run more
    partitions than GPUs.  Kernels of different ranks never wait on each other, so running them back to back on one
    stream is safe."""

    def __init__(self, model: WiredModel, n_ranks: int, device: int = 0, peer: bool = True, **engine_opts):
        import torch

        self.torch = torch
        self.pm = partition_model(model, n_ranks)
        self.engines = [Engine(self.pm, device=device, rank=r, n_ranks=n_ranks, **engine_opts) for r in range(n_ranks)]
        self.n = n_ranks
        self.peer = bool(peer) and n_ranks > 1
        if self.peer:
            # peer mode inside one process: every engine writes boundary events straight into the others' calendars
            ex = [e.peer_export() for e in self.engines]
            for e in self.engines:
                e.peer_connect_local([x[2] for x in ex], [x[1] for x in ex])
        else:
            self.xt = [e.exchange_tensors() for e in self.engines]
            self.ct = [e.control_tensors() for e in self.engines]

    def setup(self):
        for e in self.engines:
            e.setup()
        for e in self.engines:
            e.sync()

    def step(self):
        t = self.torch
        for e in self.engines:
            e.dispatch()
        for e in self.engines:
            e.sync()
        if not self.peer:
            for r in range(self.n):  # inbox[r][src] = outbox[src][r]
                for src in range(self.n):
                    self.xt[r][1][src].copy_(self.xt[src][0][r])
            t.cuda.synchronize()
            for e in self.engines:
                e.ingest()
        for e in self.engines:
            e.advance()
        for e in self.engines:
            e.sync()

    def run(self, max_windows: int = 0, check_every: int = 16) -> bool:
        done = 0
        while True:
            for _ in range(check_every):
                self.step()
                done += 1
                if max_windows and done >= max_windows:
                    break
            sts = [e.status() for e in self.engines]
            for e, st in zip(self.engines, sts):
                e.raise_on_error(st)
            if all(st.finished for st in sts):
                return True
            if max_windows and done >= max_windows:
                return False

    def run_to_report(self, report_time: int):
        """Periodic statistic output with every rank in this process: (records of all components in ENGINE id order as
        of the END of ``report_time``, finished), or (None, finished) -- see DistributedRunner.run_to_report."""
        st = self.engines[0].status()
        if st.finished:
            return None, True
        stop_at = int(self.pm.stop_at)
        if stop_at and report_time >= stop_at:
            return None, False
        w = self.engines[0].window()
        before = report_time // w - st.now // w
        if before < 0:
            raise EngineError("run_to_report: the report time lies before the current window")
        if before and self.run(max_windows=before, check_every=min(before, 16)):
            return None, True
        for e in self.engines:
            e.report_begin(report_time)
        self.step()
        rec = np.concatenate([e.report_end() for e in self.engines], axis=0)
        sts = [e.status() for e in self.engines]
        for e, s_ in zip(self.engines, sts):
            e.raise_on_error(s_)
        return rec, all(s_.finished for s_ in sts)

    def results(self) -> np.ndarray:
        """Results of all components in ENGINE id order (use self.pm.orig_id to map back)."""
        return np.concatenate([e.results() for e in self.engines], axis=0)

    def close(self):
        for e in self.engines:
            e.close()


def _as_tensor(torch, ptr: int, n_words: int, device):
    """Zero-copy int64 view of engine-owned device memory (``__cuda_array_interface__``)."""

    class _Holder:
        pass

    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (n_words,), "typestr": "<i8", "data": (ptr, False), "version": 3}
    return torch.as_tensor(h, device=device)


# ------------------------------------------------------------------------------------------------ RNG streams
def rng_streams(kind: int, s1: int, s2: int, n_streams: int, draws: int, stride1: int = 0, stride2: int = 0,
                want_output: bool = True, device: int = 0):
    """SST::RNG streams on the device (BASELINE config 2).  Returns (draws [n_streams, draws] u32 or None,
    checksums [n_streams] u64) as numpy arrays."""
    import torch

    dev = torch.device("cuda", device)
    out = torch.empty((n_streams, draws), dtype=torch.int32, device=dev) if want_output else None
    cs = torch.zeros(n_streams, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    _check(lib().sstb200_rng_streams(device, stream, kind, s1, s2, stride1, stride2, n_streams, draws,
                                     out.data_ptr() if want_output else None, cs.data_ptr()), "rng_streams")
    torch.cuda.synchronize(dev)
    o = out.cpu().numpy().view(np.uint32) if want_output else None
    return o, cs.cpu().numpy().view(np.uint64)


def rng_ticks(kind: int, s1: int, s2: int, first: int, n: int, device: int = 0) -> np.ndarray:
    out = np.zeros((n, 5), dtype=np.uint64)
    _check(lib().sstb200_rng_ticks(device, kind, s1, s2, first, n, _p(out)), "rng_ticks")
    return out


def rng_distribution(dist: int, params, rng_kind: int, s1: int, s2: int, n: int, device: int = 0) -> np.ndarray:
    p = np.ascontiguousarray(params, dtype=np.float64)
    out = np.zeros(n, dtype=np.float64)
    _check(lib().sstb200_rng_distribution(device, dist, _p(p), len(p), rng_kind, s1, s2, n, _p(out)), "rng_dist")
    return out

"""ConfigGraph -> WiredModel: the flat arrays the engine's C-ABI consumes.

Follows the reference's wire-up (``Simulation::prepareLinks`` simulation.cc:668-820, ``performWireUp`` :824-859,
``BaseComponent::configureLink_impl`` baseComponent.cc:541-642): components are constructed in id order, every
``configureLink`` call claims the next delivery-order tag, each link end sends with its OWN latency
(``lp.getLeft()->setLatency(clink->latency_[0])``), and the lookahead is the minimum link latency
(``ConfigGraph::getMinimumPartitionLatency`` configGraph.cc:711-730).

WiredModel layout (all numpy, C-contiguous; this is what include/sstb200.h ``sstb200_engine_create`` takes):
    comp_type   u32 [ncomp]        native type code
    comp_param  i64 [ncomp, 8]     type-specific construction params
    port_off    u32 [ncomp+1]      CSR over directed ports, ports in configureLink order
    peer_port   u32 [nport]        directed port that receives what this port sends (0xFFFFFFFF = unconnected)
    latency     u64 [nport]        cycles added when sending FROM this port
    clock_period u64 [ncomp]       0 = no clock
    rank        u32 [ncomp]        partition (GPU) owning the component; components are renumbered so that each
                                   rank owns one contiguous id range (orig_id maps back)
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from .config import ConfigGraph, ModelError
from .elements import ELEMENTS, NPARAM, StatDesc
from .timelord import TimeLord

NO_PEER = np.uint32(0xFFFFFFFF)


def pinned_empty(shape, dtype) -> np.ndarray:
    """numpy array backed by page-locked host memory (a pinned torch byte tensor owns the storage)."""
    import torch

    dtype = np.dtype(dtype)
    shape = tuple(int(v) for v in (shape if isinstance(shape, (tuple, list)) else (shape,)))
    nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
    t = torch.empty(max(nbytes, 1), dtype=torch.uint8, pin_memory=True)
    return t.numpy()[:nbytes].view(dtype).reshape(shape)


def pinned_like(a: np.ndarray) -> np.ndarray:
    out = pinned_empty(a.shape, a.dtype)
    np.copyto(out, a)
    return out


@dataclass
class WiredModel:
    comp_type: np.ndarray
    comp_param: np.ndarray
    port_off: np.ndarray
    peer_port: np.ndarray
    latency: np.ndarray
    clock_period: np.ndarray
    stop_at: int = 0
    timebase: str = "1ps"
    stats_enabled: bool = False
    rank: Optional[np.ndarray] = None
    n_ranks: int = 1
    orig_id: Optional[np.ndarray] = None  # engine id -> id in the ConfigGraph
    names: Optional[List[str]] = None
    stats: Optional[List[List[StatDesc]]] = None
    finish: Optional[list] = None
    nocut_pairs: Optional[np.ndarray] = None  # [k,2] component pairs joined by no-cut links
    stat_output: tuple = ("sst.statOutputConsole", {})
    # per directed port: (full name of the (sub)component that configured the link, its ELI type, port name) -- the
    # handler metadata the reference hands to profile tools (baseComponent.cc:593-600); None for unwired ports
    port_info: Optional[list] = None
    type_names: Optional[List[str]] = None  # ELI type of every top-level component
    stat_rate: int = 0  # cycles between periodic statistic outputs ("rate" statistic parameter), 0 = only at the end
    # u8 [nport] or None: 1 = configureSelfLink port (peer_port[p] == p, latency may be 0, delivered after the events of
    # the component's other ports at equal time; never constrains the lookahead)
    port_self: Optional[np.ndarray] = None
    output_fmt: Optional[list] = None  # per component: (code, a, b) -> console line, or None
    xparam_off: Optional[np.ndarray] = None  # u32 [ncomp+1] CSR over xparam, None when no component has any
    xparam: Optional[np.ndarray] = None      # i64
    stat_slots: Optional[list] = None        # per component: list of elements.StatSlot / None (device statistics engine)
    construct_console: Optional[list] = None  # per component: lines its constructor prints

    @property
    def ncomp(self) -> int:
        return int(self.comp_type.shape[0])

    @property
    def nport(self) -> int:
        return int(self.port_off[-1])

    def pin(self) -> "WiredModel":
        """Move the six model arrays into page-locked host memory (in place), so that ``sstb200_engine_create``'s
        uploads run at PCIe speed instead of through the driver's pageable staging (2.2 GB for a 4096x4096 torus).
        Needs a CUDA device; the arrays stay ordinary numpy arrays."""
        for f in ("comp_type", "comp_param", "port_off", "peer_port", "latency", "clock_period"):
            setattr(self, f, pinned_like(getattr(self, f)))
        return self

    def min_latency(self) -> int:
        keep = self.peer_port != NO_PEER
        if self.port_self is not None:
            keep &= self.port_self == 0
        m = self.latency[keep]
        return int(m.min()) if m.size else 0

    def validate(self):
        assert self.comp_type.dtype == np.uint32 and self.comp_param.dtype == np.int64
        assert self.comp_param.shape == (self.ncomp, NPARAM)
        assert self.port_off.dtype == np.uint32 and self.port_off.shape == (self.ncomp + 1,)
        assert self.peer_port.dtype == np.uint32 and self.latency.dtype == np.uint64
        assert self.peer_port.shape == (self.nport,) and self.latency.shape == (self.nport,)
        assert self.clock_period.dtype == np.uint64 and self.clock_period.shape == (self.ncomp,)
        conn = self.peer_port != NO_PEER
        if conn.any():
            pp = self.peer_port[conn].astype(np.int64)
            assert pp.max() < self.nport
            # links are symmetric: my peer's peer is me
            idx = np.nonzero(conn)[0]
            assert np.array_equal(self.peer_port[pp].astype(np.int64), idx), "peer_port is not an involution"
            real = conn if self.port_self is None else conn & (self.port_self == 0)
            if (self.latency[real] == 0).any():
                raise ModelError("zero-latency links are not allowed (link.cc / configGraph checks)")
        if self.port_self is not None:
            assert self.port_self.dtype == np.uint8 and self.port_self.shape == (self.nport,)
            sp = np.nonzero(self.port_self)[0]
            assert np.array_equal(self.peer_port[sp].astype(np.int64), sp), "a self-link port must be its own peer"


def wire(graph: ConfigGraph, stop_at_override: Optional[str] = None) -> WiredModel:
    tl = TimeLord(graph.program_options.get("timebase", "1ps"))
    wired = []
    for c in graph.components:
        fn = ELEMENTS.get(c.type)
        if fn is None:
            raise ModelError(
                f"can't find requested component or subcomponent '{c.type}' among the GPU-eligible element types "
                f"({', '.join(sorted(ELEMENTS))})")
        wired.append(fn(c, tl))
    ncomp = len(wired)
    port_off = np.zeros(ncomp + 1, dtype=np.uint32)
    for i, w in enumerate(wired):
        port_off[i + 1] = port_off[i] + len(w.ports) + len(w.self_links)
    nport = int(port_off[-1])
    peer_port = np.full(nport, NO_PEER, dtype=np.uint32)
    latency = np.zeros(nport, dtype=np.uint64)
    port_self = None
    if any(w.self_links for w in wired):
        port_self = np.zeros(nport, dtype=np.uint8)
        for i, w in enumerate(wired):
            for k in range(len(w.self_links)):  # after the configureLink ports: a self link's events sort last
                p = int(port_off[i]) + len(w.ports) + k
                peer_port[p], port_self[p] = p, 1
    # (owner ConfigComponent id(), port name) -> directed port id
    where = {}
    for i, w in enumerate(wired):
        for k, (owner, pname) in enumerate(w.ports):
            if pname is not None:
                where[(id(owner), pname)] = int(port_off[i]) + k
    used_ends = 0
    for l in graph.links:
        if len(l.ends) != 2:
            continue
        pids = []
        for owner, pname, lat in l.ends:
            pid = where.get((id(owner), pname))
            if pid is None:
                raise ModelError(f"port {pname} of {owner.full_name()} (link {l.name}) is connected in the model "
                                 f"but is not a port the element configures")
            pids.append(pid)
        for k in (0, 1):
            peer_port[pids[k]] = pids[1 - k]
            latency[pids[k]] = tl.cycles(l.ends[k][2])
            used_ends += 1
    stop = stop_at_override if stop_at_override is not None else graph.program_options.get("stop-at")
    stop_at = tl.cycles(stop) if stop not in (None, "", "0 ns") else 0
    nocut = [(l.ends[0][0].top().id, l.ends[1][0].top().id) for l in graph.links if l.no_cut and len(l.ends) == 2]
    # periodic output: the "rate" parameter of enableAllStatistics... (statengine.cc:152-200 registers a statistic clock
    # per distinct rate); one rate for the whole model is supported here
    rates = set()
    if graph.enable_all_stats and graph.enable_all_stats.get("rate"):
        rates.add(str(graph.enable_all_stats["rate"]))
    legacy = any(w.stats for w in wired)  # statistics output through the result records (one rate per model)
    if not legacy:
        rates.clear()
    for c, w in zip(graph.components, wired):
        if w.stats and c.stat_params.get("rate"):
            rates.add(str(c.stat_params["rate"]))
    rates = {tl.cycles(r) for r in rates if r not in ("0", "0ns", "0 ns")}
    if len(rates) > 1:
        raise ModelError("statistics with different output rates in one model are not supported on the device engine")
    # only ENABLED statistics exist for the output (a statistic that is not enabled is a NullStatistic in the reference,
    # baseComponent.cc registerStatistic): everything when enableAllStatisticsForAllComponents was called, else what was
    # enabled on the (sub)component that registers it -- all of its statistics, or the named ones
    owners = {}

    def index(c):
        owners[c.full_name()] = c
        for sc in c.subs:
            index(sc)

    for c in graph.components:
        index(c)

    def enabled(sd):
        if graph.enable_all_stats is not None:
            return True
        o = owners.get(sd.owner_name)
        return o is not None and (o.stats_enabled == "all" or (isinstance(o.stats_enabled, list) and sd.name in o.stats_enabled))

    for w in wired:
        w.stats = [sd for sd in w.stats if enabled(sd)]
    m = WiredModel(
        comp_type=np.array([w.type_code for w in wired], dtype=np.uint32),
        comp_param=np.array([w.params for w in wired], dtype=np.int64).reshape(ncomp, NPARAM),
        port_off=port_off, peer_port=peer_port, latency=latency,
        clock_period=np.array([w.clock_period for w in wired], dtype=np.uint64),
        stop_at=stop_at, timebase=tl.timebase_string,
        stats_enabled=any(w.stats for w in wired),
        names=[c.name for c in graph.components],
        stats=[w.stats for w in wired], finish=[w.finish for w in wired],
        nocut_pairs=np.array(nocut, dtype=np.uint32).reshape(-1, 2),
        stat_output=graph.stat_output,
        port_info=[(owner.full_name(), owner.type, pname) if pname is not None else None
                   for c, w in zip(graph.components, wired)
                   for owner, pname in list(w.ports) + [(c, n) for n in w.self_links]],
        port_self=port_self, output_fmt=[w.output_fmt for w in wired],
        stat_slots=[w.stat_slots for w in wired] if any(w.stat_slots for w in wired) else None,
        construct_console=[w.console for w in wired] if any(w.console for w in wired) else None,
        type_names=[c.type for c in graph.components],
        stat_rate=rates.pop() if rates else 0,
    )
    if any(w.xparams for w in wired):
        m.xparam_off = np.zeros(ncomp + 1, dtype=np.uint32)
        np.cumsum([len(w.xparams) for w in wired], out=m.xparam_off[1:])
        m.xparam = np.array([v for w in wired for v in w.xparams], dtype=np.int64)
    m.validate()
    return m

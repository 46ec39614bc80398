"""Handler-count profile tools on the device: ``--enable-profiling`` / ``--profiling-output`` for the hot path.

The reference attaches ``ProfileTool`` objects to every event and clock handler (profile/eventHandlerProfileTool.h:34-66,
clockHandlerProfileTool.h) and prints one record per tool at the end of the run (Simulation::printProfilingInfo
simulation.cc:2624-2638, util/perfReporter.cc).  Here the counting happens inside ``dispatch_kernel`` -- deliveries per
receiving port, clock-handler calls per component (``sstb200_options.handler_counts``) -- and this module does the two
host-side halves: parsing the option string exactly as ``Simulation::initializeProfileTools`` does
(simulation.cc:1607-1760; same fatal messages) and folding the counters into the reference's keys and text format.

In scope: ``sst.profile.handler.event.count`` (params ``level`` = global|type|component|subcomponent, ``track_ports``,
``profile_receives``), ``sst.profile.handler.clock.count`` (``level``), their ``.time.high_resolution`` / ``.time.steady``
variants (eventHandlerProfileTool.h:76-120: the time reported is SM time inside the handlers, measured with the SM cycle
counter around every handler call -- the device has no per-call wall clock) and the sync tools ``sst.profile.sync.count``
/ ``sst.profile.sync.time.*`` (syncProfileTool.h:60-140: rank syncs = lookahead windows of a multi-rank run, events and
bytes sent to other ranks, time waited at the window barrier).  ``profile_sends`` is not provided.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

from .config import ModelError
from .timelord import TimeLord
from .wireup import WiredModel

FORMAT_ERROR = ("ERROR: Invalid format for argument passed to --enable-profiling.  Argument should be a semi-colon "
                "separated list where each item specified details for a given profiling point using the following "
                "format: point=type(key=value,key=value,...).  Params are optional and can only be specified if a "
                "type is supplied.  Type is also optional and a default type will be used if not specified.")

EVENT_COUNT = "sst.profile.handler.event.count"
CLOCK_COUNT = "sst.profile.handler.clock.count"
EVENT_TIME = ("sst.profile.handler.event.time.high_resolution", "sst.profile.handler.event.time.steady")
CLOCK_TIME = ("sst.profile.handler.clock.time.high_resolution", "sst.profile.handler.clock.time.steady")
SYNC_COUNT = "sst.profile.sync.count"
SYNC_TIME = ("sst.profile.sync.time.high_resolution", "sst.profile.sync.time.steady")
ALL_TYPES = (EVENT_COUNT, CLOCK_COUNT, SYNC_COUNT) + EVENT_TIME + CLOCK_TIME + SYNC_TIME
LEVELS = ("global", "type", "component", "subcomponent")


@dataclass
class Tool:
    name: str
    type: str
    params: Dict[str, str] = field(default_factory=dict)
    points: List[str] = field(default_factory=list)

    def flag(self, key: str, default: bool) -> bool:
        v = self.params.get(key)
        if v is None:
            return default
        return v.strip().lower() in ("1", "true", "t", "yes", "y", "on")


def parse_enable_profiling(spec: str) -> List[Tool]:
    """``name:type(key=value,...)[point,point];...`` -> tools, in name order (the reference keeps a std::map)."""
    tools: Dict[str, Tool] = {}
    for item in (t for t in spec.split(";") if t):
        end = item.find(":")
        if end < 0:
            raise ModelError(FORMAT_ERROR)
        name = item[:end].strip()
        lb = item.find("[", end + 1)
        info = (item[end + 1:lb] if lb >= 0 else item[end + 1:]).strip()
        rb = item.find("]", lb + 1) if lb >= 0 else -1
        points = item[lb + 1:rb if rb >= 0 else len(item)].strip() if lb >= 0 else ""
        lp = info.find("(")
        params: Dict[str, str] = {}
        if lp < 0:
            typ = info
        else:
            typ = info[:lp].strip()
            rp = info.find(")", lp + 1)
            for pair in (p for p in info[lp + 1:rp if rp >= 0 else len(info)].strip().split(",") if p):
                kv = [s.strip() for s in pair.split("=") if s.strip()]
                if len(kv) < 2:
                    raise ModelError(f"ERROR: Invalid format for params ({pair}), format should be key=value")
                params[kv[0]] = kv[1]
        if name in tools:
            raise ModelError(f"ERROR: Cannot reuse tool name: {name}")
        pts = [p.strip() for p in points.split(",") if p.strip()]
        for p in pts:
            if p not in ("clock", "event", "sync"):
                raise ModelError(f"ERROR: Invalid profile point specified: {p}")
        if typ not in ALL_TYPES:
            raise ModelError(f"profile tool type '{typ}' is not available on the device engine "
                             f"(supported: {', '.join(ALL_TYPES)})")
        level = params.get("level", "type")
        if level not in LEVELS and "sync" not in typ:
            what = "EventHandlerProfileTool" if ".event." in typ else "ClockHandlerProfileTool"
            raise ModelError(f"ERROR: unsupported level specified for {what}: {level}")
        tool = Tool(name, typ, params, pts)
        if ".event." in typ and tool.flag("profile_sends", False):
            raise ModelError("profile_sends is not available on the device engine")
        tools[name] = tool
    return [tools[k] for k in sorted(tools)]


def _key(level: str, owner_name: str, owner_type: str) -> str:
    """EventHandlerProfileTool::getKeyForHandler (eventHandlerProfileTool.cc:49-76)."""
    if level == "global":
        return "global"
    if level == "type":
        return owner_type
    if level == "component":
        return owner_name.split(":", 1)[0]
    return owner_name


def event_counts(m: WiredModel, recv: np.ndarray, tool: Tool) -> Dict[str, int]:
    """Fold per-port delivery counts (ENGINE port order) into the tool's keys.  Every configured handler gets a key,
    also when it never ran (the reference registers the key at configureLink time)."""
    if m.port_info is None:
        raise ModelError("the model carries no port names (built by a vectorised builder): handler keys are unknown")
    level, track = tool.params.get("level", "type"), tool.flag("track_ports", False)
    out: Dict[str, int] = {}
    for pid, info in enumerate(m.port_info):
        if info is None:
            continue
        owner_name, owner_type, pname = info
        k = _key(level, owner_name, owner_type)
        if track:
            k = f"{k}:{pname}"
        out[k] = out.get(k, 0) + int(recv[pid])
    return out


def clock_counts(m: WiredModel, calls: np.ndarray, tool: Tool, only: Optional[np.ndarray] = None) -> Dict[str, int]:
    """ClockHandlerProfileTool::getKeyForHandler (clockHandlerProfileTool.cc:44-68): one handler per clocked component
    (``only``: boolean mask of the components to include, e.g. those of one rank)."""
    level = tool.params.get("level", "type")
    out: Dict[str, int] = {}
    for i in range(m.ncomp):
        if int(m.clock_period[i]) == 0 or (only is not None and not only[i]):
            continue
        name = m.names[i] if m.names else f"component{i}"
        k = _key(level, name, m.type_names[i] if m.type_names else "")
        out[k] = out.get(k, 0) + int(calls[i])
    return out


def needs_times(tools: List[Tool]) -> bool:
    return any(t.type in EVENT_TIME + CLOCK_TIME for t in tools)


def _si_time(seconds: float) -> str:
    """UnitAlgebra::toStringBestSI of a time: '1.788 ms', '311 ns', '0 s'"""
    if seconds <= 0:
        return "0 s"
    for pre, scale in (("", 1.0), ("m", 1e-3), ("u", 1e-6), ("n", 1e-9), ("p", 1e-12)):
        if seconds >= scale:
            return f"{seconds / scale:.4g} {pre}s"
    return f"{seconds / 1e-12:.4g} ps"


def _aligned(indent: str, items: Dict[str, str]) -> str:
    width = max(len(k) for k in items) + 1
    return "".join(f"{indent}{(k + ':').ljust(width)} {v}\n" for k, v in sorted(items.items()))


def profiling_text(m: WiredModel, tools: List[Tool], recv: np.ndarray, calls: np.ndarray, end_time: int,
                   input_file: str, ranks: int = 1, threads: int = 1, times=None, sync=None) -> str:
    """The file ``--profiling-output`` names, laid out as the reference's PerfReporter does (observed from the
    reference binary; tests/golden holds the fixtures): one ``name:`` record per tool in name order, clock tools
    grouped under ``<rank>_<thread>:`` (one group per GPU here), then the run summary.  ``recv`` / ``calls`` are in
    ENGINE port / component order."""
    records = {}
    for t in tools:
        out = []
        if t.type == EVENT_COUNT:
            if t.flag("profile_receives", True):
                for k, v in sorted(event_counts(m, recv, t).items()):
                    out.append(f"  {k}:\n    recv_count: {v}\n")
        elif t.type in EVENT_TIME:
            # recv_count, recv_time (seconds, best SI), avg_recv_time (whole ns) -- eventHandlerProfileTool.cc:140-155
            if times is None:
                raise ModelError("handler-time tools need the engine's handler times (handler_counts=3)")
            if t.flag("profile_receives", True):
                cnt = event_counts(m, recv, t)
                ns = event_counts(m, np.rint(times[0]).astype(np.int64), t)
                for k in sorted(cnt):
                    items = {"recv_count": str(cnt[k]), "recv_time": _si_time(ns[k] * 1e-9)}
                    if cnt[k]:
                        items["avg_recv_time"] = _si_time((ns[k] // cnt[k]) * 1e-9)
                    out.append(f"  {k}:\n" + _aligned("    ", items))
        elif t.type in CLOCK_TIME:
            if times is None:
                raise ModelError("handler-time tools need the engine's handler times (handler_counts=3)")
            for r in (range(m.n_ranks) if m.rank is not None else [0]):
                out.append(f"  {r}_0:\n")
                only = (m.rank == r) if m.rank is not None else None
                cnt = clock_counts(m, calls, t, only)
                ns = clock_counts(m, np.rint(times[1]).astype(np.int64), t, only)
                for k in sorted(cnt):
                    items = {"count": str(cnt[k]), "handler_time": _si_time(ns[k] * 1e-9)}
                    if cnt[k]:
                        items["avg_handler_time"] = _si_time((ns[k] // cnt[k]) * 1e-9)
                    out.append(f"    {k}:\n" + _aligned("      ", items))
        elif t.type == SYNC_COUNT or t.type in SYNC_TIME:
            # one child per rank with the tool's keys (syncProfileTool.cc:58-120); threads do not exist here
            for r, (syncs, events, nbytes, wait_ns) in enumerate(sync or [(0, 0, 0, 0)]):
                items = {"ranksync_total_count": str(syncs), "threadsync_total_count": "0",
                         "ranksync_total_events": str(events), "ranksync_total_data": f"{nbytes} B"}
                if t.type in SYNC_TIME:
                    items.update({"ranksync_total_time": _si_time(wait_ns * 1e-9),
                                  "ranksync_avg_time": _si_time((wait_ns // syncs if syncs else 0) * 1e-9),
                                  "threadsync_total_time": "0 s", "threadsync_avg_time": "0 s"})
                out.append(f"  rank{r}:\n" + _aligned("    ", items))
        else:
            for r in (range(m.n_ranks) if m.rank is not None else [0]):
                out.append(f"  {r}_0:\n")
                only = (m.rank == r) if m.rank is not None else None
                for k, v in sorted(clock_counts(m, calls, t, only).items()):
                    out.append(f"    {k}:\n      count: {v}\n")
        records[t.name] = f"\n\n{t.name}:\n" + "".join(out)
    # the run summary is one more record, named "metadata" (main.cc:1640-1655); records print in name order
    records["metadata"] = ("\n\nSimulation Summary:\n"
                           f"  Simulation Input File: {input_file}\n"
                           f"  Ranks:                 {ranks}\n"
                           f"  Simulated time:        {TimeLord(m.timebase).best_si(end_time)}\n"
                           f"  Threads:               {threads}\n")
    return "".join(records[k] for k in sorted(records))

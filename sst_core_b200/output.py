"""finish() console lines and end-of-simulation statistic output.

Formats follow the reference so its testsuites' sorted-diff comparison can be pointed at this engine:
* per-component finish() lines come from the element descriptors (e.g. ``"%d received %d messages"``,
  enclosingComponent.cc:84-89), then ``Simulation is complete, simulated time: <best SI>`` (main.cc);
* ``StatisticOutput.csv`` rows (statapi/statoutputcsv.cc, statengine.cc:476-548): header
  ``ComponentName, StatisticName, StatisticSubId, StatisticType, SimTime, Rank, Sum.<t>, SumSQ.<t>, Count.u64,
  Min.<t>, Max.<t>``; one row per registered statistic, components in id order, statistics in registration order.
"""
from __future__ import annotations

from typing import List

import numpy as np

from .config import ModelError
from .timelord import TimeLord
from .wireup import WiredModel


def _signed(v: int, bits: int) -> int:
    v &= (1 << bits) - 1
    return v - (1 << bits) if v >> (bits - 1) else v


def console_lines(m: WiredModel, records: np.ndarray, end_time: int, comp_begin: int = 0) -> List[str]:
    """Lines device components printed through ctx.output (records u64 [n, 8] of Engine.output_records or the five
    arrays of the oracle; component ids are engine ids).  Nothing at a later time than the end of the simulation is
    shown: the engine finishes its last lookahead window, the reference stops right after the Exit time."""
    if records is None or not m.output_fmt:
        return []
    lines = []
    for r in records:
        t, comp, code = int(r[0]), int(r[1]) & 0xFFFFFFFF, int(r[1]) >> 32
        fn = m.output_fmt[comp]
        if fn is not None and t <= end_time:
            lines.append(fn(code, int(r[2]), int(r[3])))
    return lines


def _stat_values(dtype: str, raw) -> list:
    """Sum, SumSQ, Count, Min, Max of an output record as the numbers the reference would print"""
    import struct

    def one(v):
        v = int(v)
        if dtype == "u32":
            return v & 0xFFFFFFFF
        if dtype == "i32":
            return _signed(v & 0xFFFFFFFF, 32)
        if dtype == "i64":
            return _signed(v, 64)
        if dtype == "f32":
            return struct.unpack("<f", struct.pack("<I", v & 0xFFFFFFFF))[0]
        if dtype == "f64":
            return struct.unpack("<d", struct.pack("<Q", v))[0]
        return v
    return [one(raw[0]), one(raw[1]), int(raw[2]), one(raw[3]), one(raw[4])]


def _fmt(dtype: str, v) -> str:
    return f"{v:f}" if dtype in ("f32", "f64") else str(v)


def _flag(params: dict, key: str, default: bool) -> bool:
    v = params.get(key)
    return default if v is None else str(v).strip().lower() in ("1", "true", "t", "yes", "y", "on")


def stat_engine_outputs(m: WiredModel, records: np.ndarray, end_time: int):
    """Outputs of the statistics that run under the device statistics engine (elements.StatSlot; records as in
    console_lines, code = slot | end_of_simulation << 8, values = Sum, SumSQ, Count, Min, Max).  Returns
    (console lines, {file path: lines}) formatted like statapi/statoutputtxt.cc (console, TXT) and statoutputcsv.cc:
    the model's default statistic output for ungrouped statistics, each group's own output for the others."""
    if not m.stat_slots or records is None:
        return [], {}
    default = (m.stat_output[0], dict(m.stat_output[1]))
    outs = {}  # id -> (type, params, [(time, comp, slot, values)])

    def out_of(slot):
        key = id(slot.output) if slot.output is not None else 0
        if key not in outs:
            t, p = (slot.output.type, slot.output.params) if slot.output is not None else default
            outs[key] = (t.lower(), p, [])
        return outs[key]

    # field registration order = statistic registration order: components in id order, slots in order
    order = np.argsort(m.orig_id) if m.orig_id is not None else np.arange(m.ncomp)
    fields = {}
    for i in order:
        for sl in (m.stat_slots[i] or []):
            if sl is not None and sl.alias_of is None:
                t, p, _ = out_of(sl)
                fl = fields.setdefault(id(sl.output) if sl.output is not None else 0, [])
                for f in (f"Sum.{sl.dtype}", f"SumSQ.{sl.dtype}", "Count.u64", f"Min.{sl.dtype}", f"Max.{sl.dtype}"):
                    if f not in fl:
                        fl.append(f)
    for r in records:
        t, comp, code = int(r[0]), int(r[1]) & 0xFFFFFFFF, int(r[1]) >> 32
        slots = m.stat_slots[comp]
        if not slots or t > end_time:
            continue
        sl = slots[code & 0xFF]
        out_of(sl)[2].append((t, comp, sl, _stat_values(sl.dtype, r[2:7])))
    console, files = [], {}
    for key, (otype, params, recs) in outs.items():
        lines = []
        if otype in ("sst.statoutputconsole", "sst.statoutputtxt"):
            is_con = otype.endswith("console")
            simtime, rank = _flag(params, "outputsimtime", not is_con), _flag(params, "outputrank", not is_con)
            for t, comp, sl, v in recs:
                name = f"{sl.comp_name}.{sl.name}" + (f".{sl.subid}" if sl.subid != "" else "")
                rk = int(m.rank[comp]) if m.rank is not None else 0
                d = sl.dtype
                lines.append((" " if is_con else "") + f"{name} : Accumulator : " + (f"SimTime = {t}; " if simtime else "")
                             + (f"Rank = {rk}; " if rank else "")
                             + f"Sum.{d} = {_fmt(d, v[0])}; SumSQ.{d} = {_fmt(d, v[1])}; Count.u64 = {v[2]}; "
                               f"Min.{d} = {_fmt(d, v[3])}; Max.{d} = {_fmt(d, v[4])}; ")
            if is_con:
                console += lines
            else:
                files.setdefault(params.get("filepath", "StatisticOutput.txt"), []).extend(lines)
        elif otype == "sst.statoutputcsv":
            sep = params.get("separator", ", ")
            simtime, rank = _flag(params, "outputsimtime", True), _flag(params, "outputrank", True)
            fl = fields.get(key, [])
            head = ["ComponentName", "StatisticName", "StatisticSubId", "StatisticType"] + (["SimTime"] if simtime else []) \
                + (["Rank"] if rank else []) + fl
            lines.append(sep.join(head))
            for t, comp, sl, v in recs:
                d = sl.dtype
                mine = dict(zip((f"Sum.{d}", f"SumSQ.{d}", "Count.u64", f"Min.{d}", f"Max.{d}"),
                                (_fmt(d, v[0]), _fmt(d, v[1]), str(v[2]), _fmt(d, v[3]), _fmt(d, v[4]))))
                rk = int(m.rank[comp]) if m.rank is not None else 0
                lines.append(sep.join([sl.comp_name, sl.name, sl.subid, "Accumulator"] + ([str(t)] if simtime else [])
                                      + ([str(rk)] if rank else []) + [mine.get(f, "0") for f in fl]))
            files.setdefault(params.get("filepath", "StatisticOutput.csv"), []).extend(lines)
        else:
            raise ModelError(f"statistic output {otype} is not supported for device statistics")
    return console, files


def stat_end_records(m: WiredModel, results: np.ndarray, end_time: int) -> np.ndarray:
    """The end-of-simulation output of every statistic under the device statistics engine
    (StatisticProcessingEngine::endOfSimulation statengine.cc:296-311), as output records made from the result records
    ([0] ticks run, [1 + 5 i ..] the fields of statistic i).  A statistic registered at run time (register_dynamic) only
    exists once its tick has been reached."""
    rows = []
    if m.stat_slots:
        for c in range(m.ncomp):
            for i, sl in enumerate(m.stat_slots[c] or []):
                if sl is None or sl.alias_of is not None:
                    continue
                if sl.name == "stat5_dyn":
                    dyn = int(m.comp_param[c, 4])
                    if dyn <= 0 or int(results[c, 0]) < dyn:
                        continue
                rows.append([end_time, c | ((i | 0x100) << 32)] + [int(v) for v in results[c, 1 + 5 * i:6 + 5 * i]] + [0])
    return np.array(rows, dtype=np.uint64).reshape(-1, 8)


def finish_lines(m: WiredModel, results: np.ndarray, end_time: int) -> List[str]:
    """results are indexed by engine id; lines are returned in ConfigGraph id order."""
    order = np.argsort(m.orig_id) if m.orig_id is not None else np.arange(m.ncomp)
    lines = []
    if m.finish:
        for i in order:
            fn = m.finish[i]
            if fn is not None:
                s = fn([int(x) for x in results[i]])
                if s is not None:
                    lines.append(s)
    lines.append(f"Simulation is complete, simulated time: {TimeLord(m.timebase).best_si(end_time)}")
    return lines


def stat_rows_csv(m: WiredModel, results: np.ndarray, end_time: int, rank_of=None, periodic=()) -> List[str]:
    """Header + one block of rows per output: the periodic outputs first (``periodic`` = [(sim time, records), ...],
    the statistic clock of a "rate" parameter, statengine.cc:346-376), then the end-of-simulation block."""
    if not m.stats_enabled or not m.stats:
        return []
    if periodic:
        rows = stat_rows_csv(m, periodic[0][1], periodic[0][0])
        for t_p, rec in periodic[1:]:
            rows += stat_rows_csv(m, rec, t_p)[1:]
        return rows + stat_rows_csv(m, results, end_time)[1:]
    order = np.argsort(m.orig_id) if m.orig_id is not None else np.arange(m.ncomp)
    # output fields are registered as the (enabled) statistics are created, component by component: an accumulator of
    # type t brings Sum.t, SumSQ.t, Count.u64, Min.t, Max.t; Count.u64 is shared by all types (statoutput.cc
    # registerField; observed: "Sum.u64, SumSQ.u64, Count.u64, Min.u64, Max.u64, Sum.i32, SumSQ.i32, Min.i32, Max.i32")
    fields = []
    for i in order:
        for s in m.stats[i]:
            for f in (f"Sum.{s.dtype}", f"SumSQ.{s.dtype}", "Count.u64", f"Min.{s.dtype}", f"Max.{s.dtype}"):
                if f not in fields:
                    fields.append(f)
    rows = ["ComponentName, StatisticName, StatisticSubId, StatisticType, SimTime, Rank, " + ", ".join(fields)]
    for i in order:
        rk = int(m.rank[i]) if m.rank is not None else 0
        for s in m.stats[i]:
            if s.offset is None:
                vals = [0, 0, 0, 0, 0]
            else:
                vals = [int(x) for x in results[i, s.offset:s.offset + 5]]
                if s.dtype == "i32":
                    vals = [_signed(vals[0], 32), _signed(vals[1], 32), vals[2], _signed(vals[3], 32),
                            _signed(vals[4], 32)]
            mine = dict(zip((f"Sum.{s.dtype}", f"SumSQ.{s.dtype}", "Count.u64", f"Min.{s.dtype}", f"Max.{s.dtype}"), vals))
            rows.append(f"{s.owner_name}, {s.name}, {s.subid}, Accumulator, {end_time}, {rk}, "
                        + ", ".join(str(mine.get(f, 0)) for f in fields))
    return rows


def stat_lines_console(m: WiredModel, results: np.ndarray) -> List[str]:
    """``sst.statOutputConsole`` (statapi/statoutputconsole.cc): one line per statistic, e.g.
    `` component0:route.msg_count : Accumulator : Sum.u64 = 118; SumSQ.u64 = 118; Count.u64 = 118; Min.u64 = 1; Max.u64 = 1; ``"""
    if not m.stats_enabled or not m.stats:
        return []
    order = np.argsort(m.orig_id) if m.orig_id is not None else np.arange(m.ncomp)
    out = []
    for i in order:
        for s in m.stats[i]:
            t = s.dtype
            if s.offset is None:
                vals = [0, 0, 0, 0, 0]
            else:
                vals = [int(x) for x in results[i, s.offset:s.offset + 5]]
                if t == "i32":
                    vals = [_signed(vals[0], 32), _signed(vals[1], 32), vals[2], _signed(vals[3], 32),
                            _signed(vals[4], 32)]
            name = f"{s.owner_name}.{s.name}" + (f".{s.subid}" if s.subid != "" else "")
            out.append(f" {name} : Accumulator : Sum.{t} = {vals[0]}; SumSQ.{t} = {vals[1]}; Count.u64 = {vals[2]}; "
                       f"Min.{t} = {vals[3]}; Max.{t} = {vals[4]}; ")
    return out


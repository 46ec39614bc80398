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


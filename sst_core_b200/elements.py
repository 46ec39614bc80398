"""Host-side descriptors of the GPU-eligible element types (the ELI-facing half of each device component).

In the reference a component library registers builders + metadata with static initialisers
(``SST_ELI_REGISTER_COMPONENT`` component.h:93-95, ``SST_ELI_DOCUMENT_PARAMS/PORTS/STATISTICS``), and the
constructor reads ``Params`` and calls ``configureLink`` / ``registerClock`` / ``registerStatistic``
(baseComponent.cc:501-642).  A device component has the same two halves:

* device half  -- POD state + ``__device__`` handler functors, compiled into libsstb200.so
  (``sst_core_b200/csrc/components.cuh``; the native type table is queried through
  ``sstb200_component_type_*`` and cross-checked against this file at import time);
* host half    -- THIS file: for each ELI name, the rule that turns a ConfigComponent (params, subcomponent tree,
  connected ports) into the wired record the engine consumes: type code, 8 int64 params, the ordered list of
  ports exactly in the order the reference constructor would call ``configureLink`` (that order defines the
  delivery order tags, baseComponent.cc:636), the clock period, and the statistics it registers.

Errors mirror the reference constructors' ``Output::fatal`` calls (raised as ModelError).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Tuple

from .config import ConfigComponent, ModelError
from .timelord import TimeLord

TYPE_MESH, TYPE_PHOLD, TYPE_CLOCKNODE, TYPE_RNGCOMP, TYPE_CORETEST, TYPE_CLOCKER = 1, 2, 3, 4, 5, 6
TYPE_STATS_INT, TYPE_STATS_FLOAT = 7, 8
NPARAM = 8
NRESULT = 32
# words of the result record each element actually fills (the download can stop there)
RESULT_WORDS = {1: 6, 2: 7, 3: 24, 4: 7, 5: 24, 6: 13, 7: 26, 8: 16}
# ... and when statistics are disabled (the accumulator words stay zero)
RESULT_WORDS_NOSTATS = {1: 1, 2: 2, 3: 4, 4: 7, 5: 1, 6: 13, 7: 26, 8: 16}


@dataclass
class StatDesc:
    owner_name: str  # full name of the (sub)component that registered it
    name: str
    subid: str
    dtype: str  # "u64" | "i32"
    offset: Optional[int]  # word offset of the 5-word accumulator in the result record; None = never updated


@dataclass
class Wired:
    type_code: int
    params: List[int]
    ports: List[Tuple[ConfigComponent, str]]  # (owner (sub)component, port name) in configureLink order
    clock_period: int = 0
    stats: List[StatDesc] = field(default_factory=list)
    finish: Optional[Callable] = None  # (result record) -> console line or None
    # ports the constructor adds with configureSelfLink (baseComponent.cc:374-378), after its configureLink calls: links
    # from the component to itself, latency 0, delivered after the events of every other port of the same time
    self_links: List[str] = field(default_factory=list)
    # (code, a, b) of a console-output record -> the line getSimulationOutput().output() would have printed
    output_fmt: Optional[Callable] = None
    xparams: List[int] = field(default_factory=list)  # further int64 construction parameters (sstb200_model.xparam)
    # statistics that run under the device statistics engine (rate / startat / stopat / resetOnOutput per statistic):
    # one StatSlot per slot of the element, None = NullStatistic.  Their outputs arrive as output-log records.
    stat_slots: Optional[list] = None
    console: List[str] = field(default_factory=list)  # lines the constructor prints


def _find_int(c: ConfigComponent, key: str, default=None, required_msg=None) -> int:
    if key not in c.params:
        if required_msg is not None:
            raise ModelError(required_msg)
        return default
    v = c.params[key].strip()
    try:
        return int(v, 0)
    except ValueError:
        try:
            return int(float(v))
        except ValueError:
            low = v.lower()
            if low in ("true", "t", "yes"):
                return 1
            if low in ("false", "f", "no"):
                return 0
            raise ModelError(f"parameter {key}={v!r} of {c.full_name()} is not an integer")


def _connected_numbered_ports(c: ConfigComponent, stem="port") -> List[str]:
    """`while (isPortConnected("port%d"))` -- enclosingComponent.cc:129-137."""
    have = {p for p, _, _ in c.links}
    out, i = [], 0
    while f"{stem}{i}" in have:
        out.append(f"{stem}{i}")
        i += 1
    return out


# ---- coreTestElement.message_mesh.enclosing_component (testElements/message_mesh/enclosingComponent.cc:29-69)
def _wire_mesh(c: ConfigComponent, tl: TimeLord) -> Wired:
    my_id = _find_int(c, "id", -1)
    if my_id == -1:
        raise ModelError("Must specify param 'id' in EnclosingComponent")
    verbose = _find_int(c, "verbose", 1)
    mod = _find_int(c, "mod", 1)
    if mod < 1:
        raise ModelError("Modulus must be at least 1")
    nstats = _find_int(c, "stats", 0)
    slots = c.sub_in_slot("ports")
    if not slots:
        raise ModelError("Must specify at least one PortInterface SubComponent for slot 'ports' in EnclosingComponent")
    ports, counts = [], []
    for s in slots:
        mp = s
        if s.type == "coreTestElement.message_mesh.port_slot":
            inner = s.sub_in_slot("port")
            if not inner:
                raise ModelError(f"port_slot {s.full_name()} has no 'port' subcomponent")
            mp = inner[0]
        if mp.type != "coreTestElement.message_mesh.message_port":
            raise ModelError(f"unsupported PortInterface type {mp.type}")
        names = _connected_numbered_ports(mp)
        counts.append(len(names))
        ports += [(mp, n) for n in names]
    route = c.sub_in_slot("route")
    if not route or route[0].type != "coreTestElement.message_mesh.route_message":
        raise ModelError("Must specify The RouteInterface SubComponent to use for slot 'route' in EnclosingComponent")
    if len(counts) > 8 or any(k > 255 or k < 1 for k in counts):
        raise ModelError("message_mesh on device supports up to 8 port groups of 1..255 links")
    packed = 0
    for g, k in enumerate(counts):
        packed |= k << (8 * g)
    stats = [StatDesc(route[0].full_name(), "msg_count", "", "u64", 1)]
    stats += [StatDesc(c.full_name(), "stat", str(i), "u64", None) for i in range(nstats)]

    def finish(r, _id=my_id, _v=verbose):
        cnt = r[0] if r[0] < (1 << 31) else r[0] - (1 << 32)
        return f"{_id} received {cnt} messages" if _v else None

    return Wired(TYPE_MESH, [my_id, mod, verbose, nstats, len(counts), packed, 0, 0], ports, 0, stats, finish)


# ---- pdes.phold (oracle/probe/pdes_elements.cc PholdNode)
def _wire_phold(c: ConfigComponent, tl: TimeLord) -> Wired:
    my_id = _find_int(c, "id", -1)
    init = _find_int(c, "init_events", 16)
    dmod = _find_int(c, "delay_mod", 8000)
    verbose = _find_int(c, "verbose", 1)
    if dmod < 1:
        raise ModelError("delay_mod must be at least 1")
    names = _connected_numbered_ports(c)
    if not names:
        raise ModelError(f"pdes.phold {c.name} has no connected port0")

    def finish(r, _id=my_id, _v=verbose):
        return f"phold {_id} recv {r[0]} hash {r[1]:016x}" if _v else None

    return Wired(TYPE_PHOLD, [my_id, init, dmod, verbose, 0, 0, 0, 0], [(c, n) for n in names], 0,
                 [StatDesc(c.full_name(), "delay", "", "u64", 2)], finish)


# ---- pdes.clocknode (oracle/probe/pdes_elements.cc ClockNode; coreTest_Component.cc:25-173 style)
def _wire_clocknode(c: ConfigComponent, tl: TimeLord) -> Wired:
    my_id = _find_int(c, "id", -1)
    comm = _find_int(c, "commFreq", 1000)
    verbose = _find_int(c, "verbose", 1)
    if comm == 0:
        raise ModelError("commFreq must be non-zero")
    period = tl.cycles(c.params.get("clock", "1GHz"))
    have = {p for p, _, _ in c.links}
    # the 4 ports are always "configured" in order; unconnected ones stay unwired (no tag is consumed for them)
    ports = [(c, f"port{i}") if f"port{i}" in have else (c, None) for i in range(4)]
    stats = [StatDesc(c.full_name(), n, "", "i32", 4 + 5 * i) for i, n in enumerate("NSEW")]

    def finish(r, _id=my_id, _v=verbose):
        return (f"clocknode {_id} ticks {r[0]} last_cycle {r[1]} recv {r[2]} hash {r[3]:016x}") if _v else None

    return Wired(TYPE_CLOCKNODE, [my_id, comm, verbose, 0, 0, 0, 0, 0], ports, period, stats, finish)


# ---- coreTestElement.coreTestComponent (testElements/coreTest_Component.cc:25-84; tests/test_Component.py)
def _wire_coretest(c: ConfigComponent, tl: TimeLord) -> Wired:
    _find_int(c, "workPerCycle", required_msg="couldn't find work per cycle")  # a busy loop: no simulated effect
    comm = _find_int(c, "commFreq", required_msg="couldn't find communication frequency")
    _find_int(c, "commSize", 16)  # payload bytes, summed by the receiver and dropped: not carried on the device
    if comm == 0:
        raise ModelError("commFreq must be non-zero (generateNextInt32() % commFreq)")
    period = tl.cycles(c.params.get("clockFrequency", "1GHz"))
    have = {p for p, _, _ in c.links}
    names = ["Nlink", "Slink", "Elink", "Wlink"]
    for n in names:  # coreTest_Component.cc:66-69 assert(N) ... assert(W)
        if n not in have:
            raise ModelError(f"coreTestComponent {c.name}: port {n} is not connected (the reference asserts on it)")
    stats = [StatDesc(c.full_name(), n, "", "i32", 4 + 5 * i) for i, n in enumerate("NSEW")]

    def finish(r):
        return "\n".join(["bad neighbor"] * int(r[0]) + ["Component Finished."])

    return Wired(TYPE_CORETEST, [0, comm, 0, 0, 0, 0, 0, 0], [(c, n) for n in names], period, stats, finish)


# ---- coreTestElement.coreTestRNGComponent (testElements/coreTest_RNGComponent.cc:26-74)
def _wire_rngcomp(c: ConfigComponent, tl: TimeLord) -> Wired:
    count = _find_int(c, "count", 1000)
    verbose = _find_int(c, "verbose", 0)
    kind_s = c.params.get("rng", "mersenne")
    if kind_s == "mersenne":
        kind, s1, s2 = 0, _find_int(c, "seed", 1447) & 0xFFFFFFFF, 0
    elif kind_s == "marsaglia":
        w = _find_int(c, "seed_w", 0) & 0xFFFFFFFF
        z = _find_int(c, "seed_z", 0) & 0xFFFFFFFF
        if w == 0 or z == 0:
            raise ModelError("marsaglia generator without seeds is time-seeded in the reference; not reproducible")
        kind, s1, s2 = 1, z, w
    elif kind_s == "xorshift":
        kind, s1, s2 = 2, _find_int(c, "seed", 57) & 0xFFFFFFFF, 0
        if s1 == 0:
            raise ModelError("XORShiftRNG seed must be non-zero (xorshift.cc:44-50 asserts)")
    else:
        kind, s1, s2 = 0, 1447, 0
    return Wired(TYPE_RNGCOMP, [kind, s1, s2, count, verbose, 0, 0, 0], [], tl.cycles("1GHz"), [], None)


# ---- coreTestElement.coreTestClockerComponent (testElements/coreTest_ClockerComponent.cc:75-165; tests/test_Clocks.py)
def _s32(v: int) -> int:
    v &= 0xFFFFFFFF
    return v - (1 << 32) if v >= (1 << 31) else v


_CLOCKER_FMT = {
    0: "{i}: starting test sequence at {b}", 1: "{i}: Starting test sequence at {b}",
    2: "{i}: Clock {x} will restart at cycle {b}", 3: "{i}: Stopping Clock {x} at time {b}",
    4: "{i}: Terminating test sequence at {b}", 5: "{i}: Clock {x} at cycle {b}",
    6: "{i}: Self stopping Clock {x} at time {b}", 7: "{i}: ordered clock result = {x} (cycle = {b})",
}


def _clocker_line(code: int, a: int, b: int) -> str:
    return _CLOCKER_FMT[code].format(i=_s32(a), x=_s32(a >> 32), b=b)


def _wire_clocker(c: ConfigComponent, tl: TimeLord) -> Wired:
    if tl.cycles("1ns") != 1000:
        raise ModelError("coreTestClockerComponent on the device assumes the default 1 ps time base")
    user = c.sub_in_slot("user_slot")
    if not user or user[0].type != "coreTestElement.ClockSubComponent":
        raise ModelError("coreTestClockerComponent needs a coreTestElement.ClockSubComponent in slot 'user_slot' "
                         "(the reference dereferences it, coreTest_ClockerComponent.cc:150-153)")
    have = {p for p, _, _ in c.links}
    ports = [(c, n) if n in have else (c, None) for n in ("left", "right")]
    # the clocks are registered by the device constructor (several per component): no single period here
    return Wired(TYPE_CLOCKER, [0] * NPARAM, ports, 0, [], None, self_links=["inst_link"], output_fmt=_clocker_line)


# ---- coreTestElement.StatisticsComponent.int / .float (testElements/coreTest_StatisticsComponent.{h,cc};
# tests/test_StatisticsComponent_basic.py).  Which statistics exist and how they are output is decided here exactly as
# BaseComponent::registerStatistic_impl (baseComponent.h:756-803) and StatisticProcessingEngine::
# registerStatisticWithEngine (statengine.cc:107-204) decide it; the device runs the result.
@dataclass
class StatSlot:
    comp_name: str
    name: str        # statistic name as output (the shared object's name for createStatistic objects)
    subid: str
    dtype: str       # u32 u64 i32 i64 f32 f64
    output: object   # config.ConfigStatOutput or None = the model's default statistic output
    group: Optional[str]
    alias_of: Optional[int] = None  # slot whose statistic object this slot adds to (setStatistic on a shared object)


_STATS_ELI = {
    "coreTestElement.StatisticsComponent.int": [("stat1_U32", "1", "u32", 1), ("stat2_U64", "2", "u64", 2),
                                                ("stat3_I32", "3", "i32", 3), ("stat4_I64", "4", "i64", 4),
                                                ("stat5_dyn", "", "i64", 1)],
    "coreTestElement.StatisticsComponent.float": [("stat1_F32", "1", "f32", 1), ("stat2_F64", "2", "f64", 2),
                                                  ("stat3_F64", "3", "f64", 9)],
}
ST_EXISTS, ST_CLEAR, ST_EVENT = 1, 2, 4


def _truthy(v) -> bool:
    return str(v).strip().lower() in ("1", "true", "t", "yes", "y", "on")


def _value_unit(text: str):
    import re
    m = re.match(r"^\s*([-+]?[0-9]*\.?[0-9]+(?:[eE][-+]?[0-9]+)?)\s*([A-Za-z]*)\s*$", str(text))
    if not m:
        raise ModelError(f"cannot parse '{text}' as a number with a unit")
    return float(m.group(1)), m.group(2).lower()


def _stat_slot_params(p: dict, tl: TimeLord, group_freq: Optional[str]):
    """[flags, rate, startat, stopat] of one statistic from its parameter set (statengine.cc:133-201)"""
    flags = ST_EXISTS | (ST_CLEAR if _truthy(p.get("resetOnOutput", "0")) else 0)
    rate = str(p.get("rate", "0ns"))
    val, unit = _value_unit(rate)
    if unit not in ("s", "ms", "us", "ns", "ps", "hz", "khz", "mhz", "ghz") and not unit.startswith("event"):
        raise ModelError(f"Collection Rate = {rate} not valid")
    if group_freq is not None:  # statistics in a group: periodic at the group's frequency, or only at the end
        if unit.startswith("event"):
            raise ModelError("Statistics in groups must be periodic or dump at end, event-based output triggers are not allowed")
        r = tl.cycles(group_freq) if _value_unit(group_freq)[0] != 0 else 0
    elif unit.startswith("event"):
        flags |= ST_EVENT
        r = int(round(val))
    else:
        r = tl.cycles(rate) if val != 0 else 0

    def when(key):
        v = str(p.get(key, "0ns"))
        return tl.cycles(v) if v != "0ns" and _value_unit(v)[0] != 0 else 0
    return [flags, r, when("startat"), when("stopat")]


def _wire_stats(c: ConfigComponent, tl: TimeLord) -> Wired:
    if tl.cycles("1ns") != 1000:
        raise ModelError("StatisticsComponent on the device assumes the default 1 ps time base")
    is_int = c.type.endswith(".int")
    count = _find_int(c, "count", 1000)
    kind_s = c.params.get("rng", "mersenne")
    console = []
    if kind_s == "mersenne":
        kind, s1, s2 = 0, _find_int(c, "seed", 1447) & 0xFFFFFFFF, 0
        console.append(f"Using Mersenne Random Number Generator with seed = {s1}")
    elif kind_s == "marsaglia":
        w = _find_int(c, "seed_w", 0) & 0xFFFFFFFF
        z = _find_int(c, "seed_z", 0) & 0xFFFFFFFF
        if w == 0 or z == 0:
            raise ModelError("marsaglia generator without seeds is time-seeded in the reference; not reproducible")
        kind, s1, s2 = 1, z, w
        console.append(f"Using Marsaglia Random Number Generator with seeds m_z = {z}, m_w = {w}")
    else:
        kind, s1, s2 = 0, 1447, 0
        console.append(f"RNG provided but unknown {kind_s}, so using Mersenne with seed = 1447...")
    console.append("REGISTER CLOCK #1 at 1 ns")
    dyn = _find_int(c, "register_dynamic", 0) if is_int else 0
    g = c.graph
    level = c.stat_load_level if c.stat_load_level is not None else g.stat_load_level
    groups = [grp for grp in g.stat_groups if c in grp.components]
    slots, xp, first_of_obj = [], [], {}
    for i, (name, subid, dtype, eli_level) in enumerate(_STATS_ELI[c.type]):
        grp = next((gr for gr in groups if name in gr.stat_map), None)
        cfg = c.stat_cfg.get(name)
        params = None
        if grp is not None:  # configGraph.cc:222-233: the group's statistics are enabled on its components
            params = dict(cfg.params if cfg is not None else {})
            params.update(grp.stat_map[name])
        elif cfg is not None:
            params = cfg.params
        elif c.all_stat_params is not None and eli_level <= level:
            params = c.all_stat_params
        if params is None:
            slots.append(None)
            xp += [0, 0, 0, 0, i, 0]
            continue
        stat_type = str(params.get("type", "sst.AccumulatorStatistic"))
        if stat_type != "sst.AccumulatorStatistic":
            raise ModelError(f"statistic type {stat_type} is not available on the device (only sst.AccumulatorStatistic)")
        target = i
        out_name, out_sub = name, subid
        if cfg is not None and cfg.shared and grp is None:
            out_name, out_sub = cfg.name, ""
            if id(cfg) in first_of_obj:
                target = first_of_obj[id(cfg)]
            else:
                first_of_obj[id(cfg)] = i
        xp += _stat_slot_params(params, tl, grp.frequency or "0ns" if grp is not None else None) + [target, 0]
        slots.append(StatSlot(c.full_name(), out_name, out_sub, dtype, grp.output if grp is not None else None,
                              grp.name if grp is not None else None, alias_of=target if target != i else None))
    p = [kind, s1, s2, count, dyn, len(slots), 0, 0]
    have = {q for q, _, _ in c.links}
    ports = [(c, n) if n in have else (c, None) for n in ("left", "right")]  # configured, never used
    return Wired(TYPE_STATS_INT if is_int else TYPE_STATS_FLOAT, p, ports, tl.cycles("1ns"), [], None, xparams=xp,
                 stat_slots=slots, console=console)


ELEMENTS: Dict[str, Callable[[ConfigComponent, TimeLord], Wired]] = {
    "coreTestElement.message_mesh.enclosing_component": _wire_mesh,
    "pdes.phold": _wire_phold,
    "pdes.clocknode": _wire_clocknode,
    "coreTestElement.coreTestRNGComponent": _wire_rngcomp,
    "coreTestElement.coreTestComponent": _wire_coretest,
    "coreTestElement.coreTestClockerComponent": _wire_clocker,
    "coreTestElement.StatisticsComponent.int": _wire_stats,
    "coreTestElement.StatisticsComponent.float": _wire_stats,
}

TYPE_NAMES = {
    TYPE_MESH: "coreTestElement.message_mesh.enclosing_component",
    TYPE_PHOLD: "pdes.phold",
    TYPE_CLOCKNODE: "pdes.clocknode",
    TYPE_RNGCOMP: "coreTestElement.coreTestRNGComponent",
    TYPE_CORETEST: "coreTestElement.coreTestComponent",
    TYPE_CLOCKER: "coreTestElement.coreTestClockerComponent",
    TYPE_STATS_INT: "coreTestElement.StatisticsComponent.int",
    TYPE_STATS_FLOAT: "coreTestElement.StatisticsComponent.float",
}


def register_element(eli_name: str, type_code: int, wire_fn: Callable[[ConfigComponent, TimeLord], Wired],
                     result_words: int = NRESULT, result_words_nostats: Optional[int] = None) -> None:
    """The host half of an element compiled into libsstb200 from outside the tree (csrc/elements.def; type codes
    16..31): the analogue of loading an element library, elemLoader.cc:159-177.  ``wire_fn`` turns a ConfigComponent
    into the Wired record (type code, parameters, ports in configureLink order, clock, statistics)."""
    if not (0 < type_code < 32):
        raise ModelError("element type codes are 1..31")
    if TYPE_NAMES.get(type_code, eli_name) != eli_name:
        raise ModelError(f"type code {type_code} already belongs to {TYPE_NAMES[type_code]}")
    ELEMENTS[eli_name] = wire_fn
    TYPE_NAMES[type_code] = eli_name
    RESULT_WORDS[type_code] = result_words
    RESULT_WORDS_NOSTATS[type_code] = result_words if result_words_nostats is None else result_words_nostats

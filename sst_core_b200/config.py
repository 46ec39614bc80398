"""The model-description surface: a stand-in for the reference's embedded-Python module ``sst``.

Mirrors (names, argument meaning, error behaviour) the subset of
``src/sst/core/model/python/pymodel.cc:990-1045`` (module functions),
``pymodel_comp.cc:578-601,684-703`` (Component / SubComponent methods) and
``pymodel_link.cc:38-120,188-192`` (Link) that GPU-eligible models use, and builds the same graph the reference's
``ConfigGraph`` would (``model/configGraph.cc:320-455``): components numbered from 0 in creation order, links in
``sst.Link()`` creation order, a link's endpoint 0/1 in attach order.

A model script does ``import sst`` exactly as under the reference; :func:`install` puts this module into
``sys.modules['sst']`` for the duration of the script (the reference does the same through
``PyImport_AppendInittab("sst", ...)``, pymodel.cc:1144).
"""
from __future__ import annotations

import sys
import types
from typing import Optional


class ModelError(RuntimeError):
    """Raised where the reference raises a Python exception or calls Output::fatal while building the graph."""


class ConfigLink:
    __slots__ = ("id", "name", "latency", "no_cut", "ends")

    def __init__(self, lid: int, name: str, latency: Optional[str]):
        self.id = lid
        self.name = name
        self.latency = latency
        self.no_cut = False
        self.ends = []  # [(ConfigComponent, port name, latency string)]


class ConfigComponent:
    """A component or subcomponent node of the graph (model/configComponent.h)."""

    __slots__ = ("graph", "id", "name", "type", "params", "links", "subs", "parent", "slot", "slot_num",
                 "rank", "thread", "weight", "stats_enabled", "stat_params", "stat_load_level", "stat_cfg",
                 "all_stat_params", "stat_objs")

    def __init__(self, graph, cid, name, ctype, parent=None, slot=None, slot_num=0):
        self.graph = graph
        self.id = cid
        self.name = name
        self.type = ctype
        self.params = {}
        self.links = []  # [(port name, ConfigLink, endpoint index)]
        self.subs = []  # ConfigComponent children, in creation order
        self.parent = parent
        self.slot = slot
        self.slot_num = slot_num
        self.rank = None
        self.thread = None
        self.weight = 1.0
        self.stats_enabled = None  # None | "all" | [names]
        self.stat_params = {}
        self.stat_load_level = None
        # per statistic (configComponent.cc:285-330 enableStatistic): name -> ConfigStatistic; the enable-all config
        # (allStatConfig, None = not enabled); statistic objects made with createStatistic
        self.stat_cfg = {}
        self.all_stat_params = None
        self.stat_objs = []

    # full name rule: componentInfo.cc:125-144 -- parent:slot, with [n] only when the slot holds several
    def full_name(self) -> str:
        if self.parent is None:
            return self.name
        sibs = [s for s in self.parent.subs if s.slot == self.slot]
        idx = f"[{self.slot_num}]" if len(sibs) > 1 else ""
        return f"{self.parent.full_name()}:{self.slot}{idx}"

    def top(self):
        c = self
        while c.parent is not None:
            c = c.parent
        return c

    def sub_in_slot(self, slot):
        return sorted((s for s in self.subs if s.slot == slot), key=lambda s: s.slot_num)


class ConfigStatistic:
    """model/configStatistic.h ConfigStatistic: the parameters one statistic was enabled with; `shared` objects
    (Component.createStatistic) carry their own name and may be bound to several statistic slots (setStatistic)."""

    def __init__(self, params=None, shared=False, name=None):
        self.params = dict(params or {})
        self.shared = shared
        self.name = name


class ConfigStatOutput:
    def __init__(self, otype, params=None):
        self.type = str(otype)
        self.params = {str(k): _param_str(v) for k, v in (params or {}).items()}


class ConfigStatGroup:
    """model/configStatistic.h ConfigStatGroup (python/pymodel_statgroup.cc)"""

    def __init__(self, name):
        self.name = str(name)
        self.components = []  # ConfigComponent
        self.stat_map = {}  # statistic name -> params
        self.frequency = None
        self.output = None  # ConfigStatOutput, None = the default output


class ConfigGraph:
    def __init__(self):
        self.components = []  # top-level, id order
        self.links = []
        self.link_by_name = {}
        self.program_options = {}
        self.stat_output = ("sst.statOutputConsole", {})
        self.stat_load_level = 0
        self.enable_all_stats = None  # None or params dict
        self.stat_groups = []
        self.name_prefix = []
        self.comp_by_name = {}
        self.n_ranks = 1
        self.my_rank = 0
        self.n_threads = 1


_graph: Optional[ConfigGraph] = None


def _g() -> ConfigGraph:
    if _graph is None:
        raise ModelError("sst module used outside of a model-building context")
    return _graph


# ------------------------------------------------------------------------------------------------ object types
class _CompBase:
    def __init__(self, cfg: ConfigComponent):
        self._c = cfg

    def addParam(self, name, value):
        self._c.params[str(name)] = _param_str(value)

    def addParams(self, d):
        if not isinstance(d, dict):
            raise TypeError("addParams requires a dict")
        for k, v in d.items():
            self._c.params[str(k)] = _param_str(v)

    def addLink(self, link, port, latency=None):
        if not isinstance(link, Link):
            raise TypeError("addLink requires an sst.Link")
        cl = link._l
        lat = latency if latency is not None else cl.latency
        if lat is None:
            raise ModelError(f"no latency specified for link {cl.name} on port {port}")
        if len(cl.ends) >= 2:
            raise ModelError(f"Link {cl.name} already has two endpoints (configGraph.cc addLink)")
        cl.ends.append((self._c, str(port), str(lat)))
        self._c.links.append((str(port), cl, len(cl.ends) - 1))

    def setSubComponent(self, slot, ctype, slot_num=0):
        for s in self._c.subs:
            if s.slot == slot and s.slot_num == slot_num:
                raise ModelError(f"slot {slot}[{slot_num}] of {self._c.full_name()} already populated")
        child = ConfigComponent(self._c.graph, None, None, str(ctype), parent=self._c, slot=str(slot),
                                slot_num=int(slot_num))
        self._c.subs.append(child)
        return SubComponent(child)

    def getFullName(self):
        return self._c.full_name()

    def getType(self):
        return self._c.type

    def _each(self, recursive):
        todo = [self._c]
        while todo:
            c = todo.pop()
            yield c
            if recursive:
                todo += c.subs

    def enableAllStatistics(self, params=None, recursive=False):
        for c in self._each(recursive):
            c.stats_enabled = "all"
            c.stat_params = dict(params or {})
            _enable_stat(c, None, params)

    def enableStatistics(self, names, params=None, recursive=False):
        names = [names] if isinstance(names, str) else list(names)
        for c in self._each(recursive):
            if c.stats_enabled != "all":
                cur = c.stats_enabled if isinstance(c.stats_enabled, list) else []
                c.stats_enabled = cur + list(names)
            c.stat_params = dict(params or {})
            for n in names:
                _enable_stat(c, n, params)

    def createStatistic(self, name, params=None):
        """pymodel_comp.cc compCreateStatistic: a statistic object with its own name that statistic slots can be bound to"""
        cs = ConfigStatistic({str(k): _param_str(v) for k, v in (params or {}).items()}, shared=True, name=str(name))
        self._c.top().stat_objs.append(cs)
        return cs

    def setStatistic(self, stat_name, stat_obj):
        """pymodel_comp.cc compSetStatistic / ConfigComponent::reuseStatistic: bind slot `stat_name` to the object"""
        if not isinstance(stat_obj, ConfigStatistic):
            raise TypeError("setStatistic requires a statistic object made by createStatistic")
        self._c.stat_cfg[str(stat_name)] = stat_obj

    def enableStatistic(self, name, params=None, recursive=False):
        self.enableStatistics([name], params)

    def setStatisticLoadLevel(self, level, recursive=False):
        self._c.stat_load_level = int(level)

    def setCoordinates(self, *a):
        pass

    def addSharedParamSet(self, name):
        shared = _g().program_options.setdefault("__shared__", {}).get(name, {})
        for k, v in shared.items():
            self._c.params.setdefault(k, v)

    addGlobalParamSet = addSharedParamSet


class Component(_CompBase):
    def __init__(self, name, ctype):
        g = _g()
        full = "".join(p + "." for p in g.name_prefix) + str(name)
        if full in g.comp_by_name:
            raise ModelError(f"ERROR: trying to add Component with name that already exists: {full}")
        cfg = ConfigComponent(g, len(g.components), full, str(ctype))
        g.components.append(cfg)
        g.comp_by_name[full] = cfg
        super().__init__(cfg)

    def setRank(self, rank, thread=0):
        self._c.rank = int(rank)
        self._c.thread = int(thread)

    def setWeight(self, w):
        self._c.weight = float(w)


class SubComponent(_CompBase):
    pass


class Link:
    def __init__(self, name, latency=None):
        g = _g()
        full = "".join(p + "." for p in g.name_prefix) + str(name)
        self._l = ConfigLink(len(g.links), full, None if latency is None else str(latency))
        g.links.append(self._l)
        g.link_by_name[full] = self._l

    def connect(self, end0, end1):
        for end in (end0, end1):
            comp, port = end[0], end[1]
            lat = end[2] if len(end) > 2 else None
            comp.addLink(self, port, lat)

    def setNoCut(self):
        self._l.no_cut = True


def _param_str(v) -> str:
    if isinstance(v, bool):
        return "1" if v else "0"
    return str(v)


# ------------------------------------------------------------------------------------------------ module functions
def setProgramOption(name, value):
    _g().program_options[str(name)] = str(value)


def setProgramOptions(d):
    for k, v in d.items():
        setProgramOption(k, v)


def getProgramOptions():
    return {k: v for k, v in _g().program_options.items() if not k.startswith("__")}


def pushNamePrefix(p):
    _g().name_prefix.append(str(p))


def popNamePrefix():
    _g().name_prefix.pop()


def getMPIRankCount():
    return _g().n_ranks


def getMyMPIRank():
    return _g().my_rank


def getThreadCount():
    return _g().n_threads


def setThreadCount(n):
    _g().n_threads = int(n)


def setStatisticOutput(name, params=None):
    _g().stat_output = (str(name), dict(params or {}))


def setStatisticOutputOption(k, v):
    _g().stat_output[1][str(k)] = str(v)


def setStatisticOutputOptions(d):
    for k, v in d.items():
        setStatisticOutputOption(k, v)


def setStatisticLoadLevel(level):
    _g().stat_load_level = int(level)


def _enable_stat(c: ConfigComponent, name, params):
    """ConfigComponent::enableStatistic (configComponent.cc:285-330): parameters are merged into what the statistic
    (or the enable-all entry, name None) already has"""
    p = {str(k): _param_str(v) for k, v in (params or {}).items()}
    if name is None:
        if c.all_stat_params is None:
            c.all_stat_params = {}
        c.all_stat_params.update(p)
    else:
        c.stat_cfg.setdefault(str(name), ConfigStatistic()).params.update(p)


class StatisticOutput:
    """sst.StatisticOutput(type, params) -- python/pymodel_statgroup.cc"""

    def __init__(self, otype, params=None):
        self._o = ConfigStatOutput(otype, params)

    def addParam(self, k, v):
        self._o.params[str(k)] = _param_str(v)

    def addParams(self, d):
        for k, v in d.items():
            self.addParam(k, v)


class StatisticGroup:
    """sst.StatisticGroup(name): components x statistics dumped together at one frequency into one output
    (python/pymodel_statgroup.cc; configGraph.cc:222-233 enables the statistics on the components)"""

    def __init__(self, name):
        self._g = ConfigStatGroup(name)
        _g().stat_groups.append(self._g)

    def addStatistic(self, name, params=None):
        self._g.stat_map[str(name)] = {str(k): _param_str(v) for k, v in (params or {}).items()}
        if self._g.frequency is None and (params or {}).get("rate"):
            self._g.frequency = str(params["rate"])

    def addComponent(self, comp):
        if comp._c not in self._g.components:
            self._g.components.append(comp._c)

    def setFrequency(self, freq):
        self._g.frequency = str(freq)

    def setOutput(self, out):
        if not isinstance(out, StatisticOutput):
            raise TypeError("setOutput requires an sst.StatisticOutput")
        self._g.output = out._o


_NO_ARG = object()


def enableAllStatisticsForAllComponents(params=_NO_ARG):
    # pymodel.cc enableAllStatisticsForAllComponents parses "|O!" with PyDict_Type: the argument may be left out but,
    # when given, must be a dict
    if params is not _NO_ARG and not isinstance(params, dict):
        raise TypeError(f"argument 1 must be dict, not {type(params).__name__}")
    _g().enable_all_stats = dict(params) if params is not _NO_ARG else {}
    # pymodel.cc:559-561: applies to the components that exist NOW, recursively
    todo = list(_g().components)
    while todo:
        c = todo.pop()
        _enable_stat(c, None, params if params is not _NO_ARG else None)
        todo += c.subs


def enableAllStatisticsForComponentName(name, params=None, recursive=False):
    c = _g().comp_by_name.get(name)
    if c is None:
        raise ModelError(f"component name {name} not found")
    c.stats_enabled = "all"
    c.stat_params = dict(params or {})
    _enable_stat(c, None, params)


def enableAllStatisticsForComponentType(ctype, params=None, recursive=False):
    for c in _g().components:
        if c.type == ctype:
            c.stats_enabled = "all"
            c.stat_params = dict(params or {})
            _enable_stat(c, None, params)


def _named(name):
    c = _g().comp_by_name.get(name)
    if c is None:
        raise ModelError(f"component name {name} not found")
    return c


def enableStatisticsForComponentName(name, stats, params=None, recursive=False):
    """pymodel.cc enableStatisticsForComponentName: a list of statistic names (or one name) on the named component"""
    c = _named(name)
    names = [stats] if isinstance(stats, str) else list(stats)
    cur = c.stats_enabled if isinstance(c.stats_enabled, list) else []
    c.stats_enabled = cur + names if c.stats_enabled != "all" else "all"
    c.stat_params = dict(params or {})
    for n in names:
        _enable_stat(c, n, params)


def enableStatisticForComponentName(name, stat, params=None, recursive=False):
    enableStatisticsForComponentName(name, [stat], params, recursive)


def enableStatisticsForComponentType(ctype, stats, params=None, recursive=False):
    names = [stats] if isinstance(stats, str) else list(stats)
    for c in _g().components:
        if c.type == ctype:
            cur = c.stats_enabled if isinstance(c.stats_enabled, list) else []
            c.stats_enabled = cur + names if c.stats_enabled != "all" else "all"
            c.stat_params = dict(params or {})
            for n in names:
                _enable_stat(c, n, params)


def enableStatisticForComponentType(ctype, stat, params=None, recursive=False):
    enableStatisticsForComponentType(ctype, [stat], params, recursive)


def setStatisticLoadLevelForComponentName(name, level, recursive=False):
    _named(name).stat_load_level = int(level)


def setStatisticLoadLevelForComponentType(ctype, level, recursive=False):
    for c in _g().components:
        if c.type == ctype:
            c.stat_load_level = int(level)


_T0 = __import__("time").perf_counter()


def getElapsedExecutionTime():
    """seconds since the front end started, as the reference's UnitAlgebra prints it ("<x> s")"""
    return f"{__import__('time').perf_counter() - _T0:.6f} s"


def getLocalMemoryUsage():
    import resource
    return f"{resource.getrusage(resource.RUSAGE_SELF).ru_maxrss} KB"


def setCallPythonFinalize(flag):
    pass  # interpreter shutdown policy of the embedded CPython: nothing to do here


def findComponentByName(name):
    c = _g().comp_by_name.get(name)
    if c is None:
        return None
    obj = Component.__new__(Component)
    _CompBase.__init__(obj, c)
    return obj


def addSharedParam(set_name, key, value):
    _g().program_options.setdefault("__shared__", {}).setdefault(set_name, {})[str(key)] = _param_str(value)


def addSharedParams(set_name, d):
    for k, v in d.items():
        addSharedParam(set_name, k, v)


addGlobalParam = addSharedParam
addGlobalParams = addSharedParams


def exit():
    raise SystemExit(0)


_API = [
    "Component", "SubComponent", "Link", "setProgramOption", "setProgramOptions", "getProgramOptions",
    "pushNamePrefix", "popNamePrefix", "getMPIRankCount", "getMyMPIRank", "getThreadCount", "setThreadCount",
    "setStatisticOutput", "setStatisticOutputOption", "setStatisticOutputOptions", "setStatisticLoadLevel",
    "enableAllStatisticsForAllComponents", "enableAllStatisticsForComponentName",
    "enableAllStatisticsForComponentType", "findComponentByName", "addSharedParam", "addSharedParams",
    "addGlobalParam", "addGlobalParams", "exit", "enableStatisticForComponentName", "enableStatisticsForComponentName",
    "enableStatisticForComponentType", "enableStatisticsForComponentType", "setStatisticLoadLevelForComponentName",
    "setStatisticLoadLevelForComponentType", "getElapsedExecutionTime", "getLocalMemoryUsage", "setCallPythonFinalize",
    "StatisticGroup", "StatisticOutput",
]


def make_module() -> types.ModuleType:
    m = types.ModuleType("sst")
    m.__doc__ = "sst_core_b200 stand-in for the reference's embedded `sst` module"
    here = sys.modules[__name__]
    for n in _API:
        setattr(m, n, getattr(here, n))
    return m


def build_graph(script_path: str, model_options: str = "", n_ranks: int = 1, my_rank: int = 0) -> ConfigGraph:
    """Run a model script with ``sst`` bound to this module and ``sys.argv`` set from ``--model-options``
    (main.cc:311-386 start_graph_creation).  Returns the ConfigGraph it built."""
    global _graph
    import shlex

    g = ConfigGraph()
    g.n_ranks, g.my_rank = n_ranks, my_rank
    prev_graph, prev_mod, prev_argv = _graph, sys.modules.get("sst"), sys.argv
    _graph = g
    sys.modules["sst"] = make_module()
    sys.argv = [script_path] + shlex.split(model_options)
    try:
        with open(script_path) as f:
            code = compile(f.read(), script_path, "exec")
        exec(code, {"__name__": "__main__", "__file__": script_path})
    finally:
        sys.argv = prev_argv
        _graph = prev_graph
        if prev_mod is not None:
            sys.modules["sst"] = prev_mod
        else:
            sys.modules.pop("sst", None)
    for l in g.links:
        if len(l.ends) == 1:
            # the reference tolerates dangling links only with a warning at checkForStructuralErrors
            pass
    return g


def load_json(path: str, n_ranks: int = 1, my_rank: int = 0) -> ConfigGraph:
    """Load a model in the reference's JSON interchange format: what ``sst --output-json`` writes
    (``model/cfgoutput/jsonConfigOutput.cc``) and ``sst model.json`` reads (``model/json/jsonmodel.cc``).  A graph built
    and checked by stock SST can thus be handed to this engine.  Handled keys: ``program_options``,
    ``statistics_options``, ``components`` (``name``, ``type``, ``params``, ``statistics_options`` /
    ``statistics``, nested ``subcomponents`` with ``slot_name`` / ``slot_number``) and ``links`` (``name``, ``noCut``,
    ``left`` / ``right`` = ``{component: full name, port, latency}``)."""
    import json

    with open(path) as f:
        doc = json.load(f)
    g = ConfigGraph()
    g.n_ranks, g.my_rank = n_ranks, my_rank
    for k, v in doc.get("program_options", {}).items():
        if v not in ("", None):
            g.program_options[str(k)] = str(v)
    so = doc.get("statistics_options", {}) or {}
    if "statisticOutput" in so:
        g.stat_output = (so["statisticOutput"], dict(so.get("params", {}) or {}))
    g.stat_load_level = int(so.get("statisticLoadLevel", 0))

    def stats_of(node, cfg):
        o = node.get("statistics_options", {}) or {}
        if o.get("enable_all_statistics"):
            cfg.stats_enabled = "all"
        names = [st.get("name") for st in (node.get("statistics", []) or []) if st.get("name")]
        if names:
            cfg.stats_enabled = names if cfg.stats_enabled is None else cfg.stats_enabled

    def add_subs(node, parent):
        for sub in node.get("subcomponents", []) or []:
            child = ConfigComponent(g, None, None, str(sub["type"]), parent=parent, slot=str(sub["slot_name"]),
                                    slot_num=int(sub.get("slot_number", 0)))
            child.params = {str(k): _param_str(v) for k, v in (sub.get("params", {}) or {}).items()}
            stats_of(sub, child)
            parent.subs.append(child)
            add_subs(sub, child)

    for node in doc.get("components", []):
        name = str(node["name"])
        if name in g.comp_by_name:
            raise ModelError(f"ERROR: trying to add Component with name that already exists: {name}")
        cfg = ConfigComponent(g, len(g.components), name, str(node["type"]))
        cfg.params = {str(k): _param_str(v) for k, v in (node.get("params", {}) or {}).items()}
        stats_of(node, cfg)
        g.components.append(cfg)
        g.comp_by_name[name] = cfg
        add_subs(node, cfg)
    by_full = {}

    def canon(name: str) -> str:
        # the JSON writer always indexes slots ("comp:ports[2]:port[0]"); statistic names index a slot only when it
        # holds several subcomponents (componentInfo.cc:125-144).  Look names up in the always-indexed form.
        head, *rest = name.split(":")
        return ":".join([head] + [seg if seg.endswith("]") else seg + "[0]" for seg in rest])

    def index(c, key):
        by_full[key] = c
        for sc in c.subs:
            index(sc, f"{key}:{sc.slot}[{sc.slot_num}]")

    for c in g.components:
        index(c, c.name)
    for ln in doc.get("links", []):
        cl = ConfigLink(len(g.links), str(ln["name"]), None)
        cl.no_cut = bool(ln.get("noCut", False))
        for side in ("left", "right"):
            end = ln.get(side)
            if not end:
                continue
            owner = by_full.get(canon(str(end["component"])))
            if owner is None:
                raise ModelError(f"link {cl.name}: unknown component {end['component']}")
            cl.ends.append((owner, str(end["port"]), str(end["latency"])))
            owner.links.append((str(end["port"]), cl, len(cl.ends) - 1))
        g.links.append(cl)
        g.link_by_name[cl.name] = cl
    return g


def dump_json(g: ConfigGraph, path: str) -> None:
    """Write a ConfigGraph in the reference's JSON interchange format (what ``sst --output-json`` writes,
    ``model/cfgoutput/jsonConfigOutput.cc``), so that a model built through this module's ``sst`` surface can be run --
    or partitioned, or inspected -- by stock SST: ``sst model.json``.  The inverse of :func:`load_json`; like the
    reference's writer it records "all statistics enabled" per (sub)component but not their parameters."""
    import json

    all_on = g.enable_all_stats is not None

    def stat_opts(c):
        node = {}
        if all_on or c.stats_enabled == "all":
            node["statistics_options"] = {"enable_all_statistics": True}
        elif isinstance(c.stats_enabled, list) and c.stats_enabled:
            node["statistics"] = [{"name": n} for n in c.stats_enabled]
        return node

    def subs_of(c):
        out = []
        for sc in c.subs:
            node = {"slot_name": sc.slot, "slot_number": int(sc.slot_num), "type": sc.type}
            if sc.params:
                node["params"] = dict(sc.params)
            node.update(stat_opts(sc))
            kids = subs_of(sc)
            if kids:
                node["subcomponents"] = kids
            out.append(node)
        return out

    def indexed_name(c):
        parts = []
        while c.parent is not None:
            parts.append(f"{c.slot}[{c.slot_num}]")
            c = c.parent
        return ":".join([c.name] + parts[::-1])

    comps = []
    for c in g.components:
        node = {"name": c.name, "type": c.type}
        node.update(stat_opts(c))
        if c.params:
            node["params"] = dict(c.params)
        kids = subs_of(c)
        if kids:
            node["subcomponents"] = kids
        comps.append(node)
    links = []
    for l in g.links:
        node = {"name": l.name, "noCut": bool(l.no_cut), "nonlocal": False}
        for side, (owner, port, lat) in zip(("left", "right"), l.ends):
            node[side] = {"component": indexed_name(owner), "port": port, "latency": str(lat)}
        links.append(node)
    so = {"statisticLoadLevel": int(g.stat_load_level or 0), "statisticOutput": g.stat_output[0]}
    if g.stat_output[1]:
        so["params"] = dict(g.stat_output[1])
    doc = {"program_options": dict(g.program_options), "statistics_options": so, "components": comps,
           "statistics_group": [], "links": links}
    with open(path, "w") as f:
        json.dump(doc, f, indent=1)

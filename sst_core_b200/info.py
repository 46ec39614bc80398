"""`sst-info` for the device engine:  python -m sst_core_b200.info [lib.name ...]

The reference's sst-info walks the ELI database of a loaded library and prints, per element, its description,
parameters, ports and statistics (sstinfo.cc; eli/elibase.h:124-147).  Here the database is the native element table of
libsstb200.so (``sstb200_component_type_*``) joined with the host halves in elements.py.  Needs no GPU.
"""
from __future__ import annotations

import sys

from .elements import ELEMENTS, TYPE_NAMES
from .engine import check_registry, component_types

# what the host halves read (SST_ELI_DOCUMENT_PARAMS / PORTS / STATISTICS of the corresponding reference elements)
DOCS = {
    "coreTestElement.message_mesh.enclosing_component": dict(
        params=[("id", "component id (seeds MersenneRNG(id+100) of the route subcomponent)", "required"),
                ("mod", "send the initial event on a link only if draw % mod == 0", "1"), ("verbose", "print at finish", "1"),
                ("stats", "number of extra (never updated) statistics to register", "0")],
        ports=["subcomponent slots 'ports' (message_port, or port_slot wrapping one) with port0..portN", "slot 'route' (route_message)"],
        stats=["route: msg_count (u64 accumulator)", "stat<i> (registered, never updated)"]),
    "coreTestElement.coreTestRNGComponent": dict(
        params=[("rng", "mersenne | marsaglia | xorshift", "mersenne"), ("seed", "mersenne/xorshift seed", "1447 / 57"),
                ("seed_w, seed_z", "marsaglia seeds", "required for marsaglia"), ("count", "ticks to run", "1000"),
                ("verbose", "print every tick", "0")],
        ports=[], stats=[]),
    "coreTestElement.coreTestComponent": dict(
        params=[("workPerCycle", "busy-loop iterations per tick (no simulated effect)", "required"),
                ("commFreq", "send with probability 1/commFreq per tick", "required"),
                ("commSize", "payload bytes (not carried on the device)", "16"), ("clockFrequency", "clock", "1GHz")],
        ports=["Nlink", "Slink", "Elink", "Wlink (all four must be connected)"], stats=["N, S, E, W (i32 accumulators)"]),
    "pdes.phold": dict(
        params=[("id", "seeds XORShiftRNG(id+1)", "required"), ("init_events", "events sent in setup()", "16"),
                ("delay_mod", "extra delay = u32 % delay_mod cycles", "8000"), ("verbose", "print at finish", "1")],
        ports=["port0..portN"], stats=["delay (u64 accumulator of the received delays)"]),
    "pdes.clocknode": dict(
        params=[("id", "seeds MarsagliaRNG(11+id, 272727)", "required"), ("commFreq", "send with probability 1/commFreq per tick", "1000"),
                ("clock", "clock frequency", "1GHz"), ("verbose", "print at finish", "1")],
        ports=["port0..port3 (unconnected ones are skipped)"], stats=["N, S, E, W (i32 accumulators)"]),
    "coreTestElement.StatisticsComponent.int": dict(
        params=[("rng", "mersenne | marsaglia", "mersenne"), ("seed", "mersenne seed", "1447"),
                ("seed_w, seed_z", "marsaglia seeds", "required for marsaglia"), ("count", "ticks to run", "1000"),
                ("register_dynamic", "tick at which stat5_dyn is registered, 0 = never", "0")],
        ports=["left", "right (configured, never used)"],
        stats=["stat1_U32 (level 1)", "stat2_U64 (2)", "stat3_I32 (3)", "stat4_I64 (4)", "stat5_dyn (1, registered at run time)",
               "statistic parameters: rate (time or events), startat, stopat, resetOnOutput"]),
    "coreTestElement.StatisticsComponent.float": dict(
        params=[("rng", "mersenne | marsaglia", "mersenne"), ("seed", "mersenne seed", "1447"),
                ("seed_w, seed_z", "marsaglia seeds", "required for marsaglia"), ("count", "ticks to run", "1000")],
        ports=["left", "right (configured, never used)"],
        stats=["stat1_F32 (level 1)", "stat2_F64 (2)", "stat3_F64 (9)"]),
    "coreTestElement.coreTestClockerComponent": dict(
        params=[],
        ports=["left", "right", "inst_link (self link, configured by the component)",
               "subcomponent slot 'user_slot' (coreTestElement.ClockSubComponent, required)"],
        stats=[]),
}


def main(argv=None) -> int:
    want = set(argv if argv is not None else sys.argv[1:])
    check_registry()
    native = {name: (code, sb, ab, hp) for name, code, sb, ab, hp in component_types()}
    for name in sorted(ELEMENTS):
        if want and name not in want:
            continue
        code, sb, ab, hp = native[name]
        assert TYPE_NAMES[code] == name
        d = DOCS.get(name, {})
        print(f"ELEMENT {name}  (type code {code}; device state {sb} B + {ab} B aux per component; "
              f"{'8-byte payload' if hp else 'no payload'})")
        for p, desc, default in d.get("params", []):
            print(f"      PARAMETER {p}: {desc} [{default}]")
        for p in d.get("ports", []):
            print(f"      PORT {p}")
        for s in d.get("stats", []):
            print(f"      STATISTIC {s}")
    return 0


if __name__ == "__main__":
    sys.exit(main())

"""Host half of the example element library (relay_elements.cuh): what the reference gets from the library's ELI records
-- the element's name, its ports in configureLink order, its parameters -- is registered here."""
from sst_core_b200.elements import Wired, _find_int, register_element

TYPE_RELAY = 16


def _wire_relay(c, tl):
    have = {p for p, _, _ in c.links}
    ports = [(c, n) if n in have else (c, None) for n in ("port0", "port1")]

    def finish(r, _name=c.name):
        return f"{_name}: {r[0]} hops, sum {r[1]}"

    return Wired(TYPE_RELAY, [_find_int(c, "start", 0), 0, 0, 0, 0, 0, 0, 0], ports, 0, [], finish)


def register():
    register_element("demo.relay", TYPE_RELAY, _wire_relay, result_words=2)

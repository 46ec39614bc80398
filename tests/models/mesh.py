# MessageMesh model: 2-D torus of coreTestElement.message_mesh.enclosing_component, one MessagePort per direction
# (x+, x-, y+ and y- through a port_slot), 1 ns links, even-id components mark their x+ link no-cut.
# Same element types, parameters and argument convention as the reference's tests/test_MessageMesh.py, so it runs
# unchanged under the reference `sst` and under sst_core_b200.
# args: X Y [mod] [verbose] [stats] [stat_gen: 0 none, 1 dump at end]
import sys

import sst

sst.setProgramOption("stop-at", "10us")

x_size, y_size = int(sys.argv[1]), int(sys.argv[2])
mod = int(sys.argv[3]) if len(sys.argv) > 3 else 1
verbose = int(sys.argv[4]) if len(sys.argv) > 4 else 1
stats = int(sys.argv[5]) if len(sys.argv) > 5 else 0
stat_gen = int(sys.argv[6]) if len(sys.argv) > 6 else 0

links = {}


def get_link(a, b):
    name = "link_%s_%s" % (a, b)
    if name not in links:
        links[name] = sst.Link(name)
    return links[name]


LIB = "coreTestElement.message_mesh."
for i in range(x_size * y_size):
    mx, my = i % x_size, i // x_size
    comp = sst.Component("component%d" % i, LIB + "enclosing_component")
    comp.addParams({"id": i, "mod": mod, "verbose": verbose, "stats": stats})
    port_xp = comp.setSubComponent("ports", LIB + "message_port", 0)
    port_xn = comp.setSubComponent("ports", LIB + "message_port", 1)
    port_yp = comp.setSubComponent("ports", LIB + "port_slot", 2).setSubComponent("port", LIB + "message_port")
    port_yn = comp.setSubComponent("ports", LIB + "port_slot", 3).setSubComponent("port", LIB + "message_port")
    comp.setSubComponent("route", LIB + "route_message")

    here = "x%dy%d" % (mx, my)
    xp = "x%dy%d" % ((mx + 1) % x_size, my)
    xn = "x%dy%d" % ((mx - 1) % x_size, my)
    yp = "x%dy%d" % (mx, (my + 1) % y_size)
    yn = "x%dy%d" % (mx, (my - 1) % y_size)
    port_xp.addLink(get_link(here, xp), "port0", "1ns")
    if i % 2 == 0:
        get_link(here, xp).setNoCut()
    port_xn.addLink(get_link(xn, here), "port0", "1ns")
    port_yp.addLink(get_link(here, yp), "port0", "1ns")
    port_yn.addLink(get_link(yn, here), "port0", "1ns")

sst.setStatisticOutput("sst.statOutputCSV")
if stat_gen == 2:  # dump at rate (the reference script uses 250ns; an optional 7th argument overrides it)
    sst.enableAllStatisticsForAllComponents({"rate": sys.argv[7] if len(sys.argv) > 7 else "250ns"})
elif stat_gen == 1:
    sst.enableAllStatisticsForAllComponents()
sst.setStatisticLoadLevel(2)

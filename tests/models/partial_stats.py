# 3 x 2 PHOLD torus in which only SOME statistics are enabled: by component object, by name and by type
# (pdes.phold has one statistic, "delay"; pdes.clocknode has N, S, E, W).  Checks that only enabled statistics
# are collected and written.  args: [stop_at]
import sys

import sst

stop_at = sys.argv[1] if len(sys.argv) > 1 else "120ns"
sst.setProgramOption("stop-at", stop_at)
X, Y = 3, 2
nodes = []
for i in range(X * Y):
    kind = "pdes.clocknode" if i == 4 else "pdes.phold"
    c = sst.Component("node%d" % i, kind)
    c.addParams({"id": i, "init_events": 5, "delay_mod": 3000, "commFreq": 4, "verbose": 0})
    nodes.append(c)
for i in range(X * Y):
    mx, my = i % X, i // X
    sst.Link("x%d" % i).connect((nodes[i], "port0", "1ns"), (nodes[((mx + 1) % X) + my * X], "port1", "1ns"))
    sst.Link("y%d" % i).connect((nodes[i], "port2", "1ns"), (nodes[mx + ((my + 1) % Y) * X], "port3", "1ns"))
sst.setStatisticOutput("sst.statOutputCSV")
sst.setStatisticLoadLevel(3)
nodes[0].enableAllStatistics()
sst.enableStatisticForComponentName("node2", "delay")
sst.enableStatisticsForComponentType("pdes.clocknode", ["S", "W"])

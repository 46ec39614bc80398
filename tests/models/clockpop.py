# Clock-driven population: coreTestComponent-style nodes on a 2-D torus,
# clock period chosen by id % 4 from {1GHz, 500MHz, 250MHz, 2GHz}.
# args: X Y [stop_at] [commFreq] [verbose] [stats] [stat_rate]
import sys

import sst

x_size = int(sys.argv[1])
y_size = int(sys.argv[2])
stop_at = sys.argv[3] if len(sys.argv) > 3 else "1us"
comm_freq = int(sys.argv[4]) if len(sys.argv) > 4 else 1000
verbose = int(sys.argv[5]) if len(sys.argv) > 5 else 1
stats = int(sys.argv[6]) if len(sys.argv) > 6 else 0
stat_rate = sys.argv[7] if len(sys.argv) > 7 else ""  # periodic statistic output every <rate>

sst.setProgramOption("stop-at", stop_at)
CLOCKS = ["1GHz", "500MHz", "250MHz", "2GHz"]

links = {}


def get_link(a, b):
    name = "link_%s_%s" % (a, b)
    if name not in links:
        links[name] = sst.Link(name)
    return links[name]


for i in range(x_size * y_size):
    mx, my = i % x_size, i // x_size
    comp = sst.Component("node%d" % i, "pdes.clocknode")
    comp.addParams({"id": i, "clock": CLOCKS[i % 4], "commFreq": comm_freq, "verbose": verbose})
    xp, xn = (mx + 1) % x_size, (mx - 1) % x_size
    yp, yn = (my + 1) % y_size, (my - 1) % y_size
    comp.addLink(get_link("x%dy%d" % (mx, my), "x%dy%d" % (xp, my)), "port0", "1ns")
    comp.addLink(get_link("x%dy%d" % (xn, my), "x%dy%d" % (mx, my)), "port1", "1ns")
    comp.addLink(get_link("x%dy%d" % (mx, my), "x%dy%d" % (mx, yp)), "port2", "1ns")
    comp.addLink(get_link("x%dy%d" % (mx, yn), "x%dy%d" % (mx, my)), "port3", "1ns")

if stats:
    sst.setStatisticOutput("sst.statOutputCSV")
    sst.enableAllStatisticsForAllComponents({"rate": stat_rate} if stat_rate else {})
    sst.setStatisticLoadLevel(2)

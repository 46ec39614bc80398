"""10 x 10 torus of coreTestElement.coreTestComponent: the graph the reference's tests/test_Component.py builds
(a generated 1000-line script there; this is a loop with the same component names, creation order, parameters, port
wiring and link latency).  tests/test_coretest_component.py checks, where /root/reference is present, that both
scripts produce the same configuration graph.

    --model-options "[X Y [commFreq [stop_at [stats]]]]"      defaults: 10 10 1000 25us 0
"""
import sys

import sst

a = sys.argv[1:]
X = int(a[0]) if len(a) > 0 else 10
Y = int(a[1]) if len(a) > 1 else 10
comm_freq = a[2] if len(a) > 2 else "1000"
stop_at = a[3] if len(a) > 3 else "25us"
stats = int(a[4]) if len(a) > 4 else 0

sst.setProgramOption("stop-at", stop_at)

comp = {}
for y in range(Y):
    for x in range(X):
        c = sst.Component(f"c{x}_{y}", "coreTestElement.coreTestComponent")
        c.addParams({"workPerCycle": "1000", "commSize": "100", "commFreq": comm_freq})
        comp[x, y] = c

for y in range(Y):
    for x in range(X):
        north = sst.Link(f"link_ns_{x}_{y}")
        north.connect((comp[x, y], "Nlink", "10000ps"), (comp[x, (y + 1) % Y], "Slink", "10000ps"))
        east = sst.Link(f"link_ew_{x}_{y}")
        east.connect((comp[x, y], "Elink", "10000ps"), (comp[(x + 1) % X, y], "Wlink", "10000ps"))

if stats:
    sst.setStatisticLoadLevel(7)
    sst.setStatisticOutput("sst.statOutputCSV")
    sst.enableAllStatisticsForAllComponents()

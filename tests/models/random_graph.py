# Irregular random model for order-parity checks against the reference core: a mix of pdes.phold and pdes.clocknode
# components, a random number of ports each, links with different latencies at their two ends, several links between
# the same pair, unconnected clocknode ports.  The PHOLD delivery hash depends on the order of deliveries, so equal
# results mean equal per-component event order.   args: seed [components] [stop_at]
import random
import sys

import sst

seed = int(sys.argv[1])
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
stop_at = sys.argv[3] if len(sys.argv) > 3 else "400ns"
rng = random.Random(seed)
sst.setProgramOption("stop-at", stop_at)

comps, free_ports = [], []
for i in range(n):
    if rng.random() < 0.3:
        c = sst.Component("node%d" % i, "pdes.clocknode")
        c.addParams({"id": i, "commFreq": rng.choice([2, 3, 7]), "clock": rng.choice(["1GHz", "500MHz", "2GHz"]), "verbose": 1})
        nports = 4
    else:
        c = sst.Component("node%d" % i, "pdes.phold")
        c.addParams({"id": i, "init_events": rng.randint(1, 9), "delay_mod": rng.choice([1, 700, 2500, 6000]), "verbose": 1})
        nports = rng.randint(1, 6)
    comps.append(c)
    free_ports.append(["port%d" % k for k in range(nports)])

# PHOLD nodes need port0..portK connected without gaps; connect ports in order, partner chosen at random
latencies = ["1ns", "1500ps", "2ns", "3ns", "7ns"]
order = [(i, p) for i in range(n) for p in free_ports[i]]
used = set()
link_id = 0
for i, p in order:
    if (i, p) in used:
        continue
    # candidates: the lowest unused port of some other component (keeps every component's used ports contiguous)
    cands = []
    for j in range(n):
        if j == i:
            continue
        for q in free_ports[j]:
            if (j, q) not in used:
                cands.append((j, q))
                break
    if not cands:
        break
    j, q = rng.choice(cands)
    used.add((i, p))
    used.add((j, q))
    sst.Link("l%d" % link_id).connect((comps[i], p, rng.choice(latencies)), (comps[j], q, rng.choice(latencies)))
    link_id += 1

sst.setStatisticOutput("sst.statOutputCSV")
sst.setStatisticLoadLevel(3)
sst.enableAllStatisticsForAllComponents()

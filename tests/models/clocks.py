"""A chain of coreTestElement.coreTestClockerComponent: the graph the reference's tests/test_Clocks.py builds (same
component names, creation order, user_slot subcomponent, "right" -> "left" links), with the chain length and the link
latency as options.  tests/test_clocks.py checks, where /root/reference is present, that both scripts produce the same
configuration graph.

    --model-options "[N [latency]]"      defaults: 8 6ns
"""
import sys

import sst

a = sys.argv[1:]
N = int(a[0]) if len(a) > 0 else 8
latency = a[1] if len(a) > 1 else "6ns"

chain = []
for i in range(N):
    c = sst.Component(f"clocker{i}", "coreTestElement.coreTestClockerComponent")
    c.setSubComponent("user_slot", "coreTestElement.ClockSubComponent", 0)
    chain.append(c)

for i in range(N - 1):
    link = sst.Link(f"link{i}", latency)
    chain[i].addLink(link, "right")
    chain[i + 1].addLink(link, "left")

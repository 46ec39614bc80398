"""The model of the reference's tests/test_StatisticsComponent_basic.py, written as tables: 30
coreTestElement.StatisticsComponent.{int,float} components chained by 1 ns links, exercising every way of enabling
statistics (all / by name / by component name / by type, load levels), output rates in time and in events, startat /
stopat, resetOnOutput, a statistic registered at run time, shared statistic objects and statistic groups with their own
outputs (console, CSV, TXT).  tests/test_statistics_component.py checks, where /root/reference is present, that this
script and the reference's build the same configuration graph (components, parameters, every statistic's parameters,
groups, outputs, links)."""
import sst

ACC = "sst.AccumulatorStatistic"
sst.setProgramOptions({"partitioner": "roundrobin"})
sst.setStatisticLoadLevel(2)
sst.setStatisticOutput("sst.statOutputConsole", {"outputsimtime": True, "outputrank": False})

chain = []


def comp(name, kind, seed_w, seed_z, extra=None):
    c = sst.Component(name, "coreTestElement.StatisticsComponent." + kind)
    p = {"rng": "marsaglia", "count": "101", "seed_w": str(seed_w), "seed_z": str(seed_z)}
    p.update(extra or {})
    c.addParams(p)
    chain.append(c)
    return c


def acc(**kw):
    return dict({"type": ACC}, **kw)


# enabled through enableAllStatisticsForAllComponents: only the components that exist when it is called
comp("StatGlobal0", "int", 1440, 1046, {"register_dynamic": 20})
comp("StatGlobal1", "int", 1441, 1047).setStatisticLoadLevel(4)
sst.enableAllStatisticsForAllComponents(acc(rate="8 ns", startat="50 ns"))

# statistic by statistic, through the component object (0-3) and through the component's name (4-7):
# time-based rates, event-based rates, two statistics per call
by_time = [(["stat1_U32"], acc(rate="5 ns")), (["stat2_U64"], acc(rate="10 ns", resetOnOutput=True)),
           (["stat3_I32"], acc(rate="7 ns", startat="25 ns", stopat="60 ns")),
           (["stat4_I64"], acc(startat="35 ns", stopat="70 ns", rate="12 ns", resetOnOutput=True))]
by_event = [(["stat1_U32"], acc(rate="6 event")), (["stat2_U64"], acc(rate="11 events", resetOnOutput=True)),
            (["stat3_I32"], acc(rate="8 events", startat="25 ns", stopat="60 ns")),
            (["stat4_I64"], acc(startat="35 ns", stopat="70 ns", rate="14 events", resetOnOutput=True))]
pairs_time = [(["stat1_U32", "stat3_I32"], acc(rate="20 ns")),
              (["stat2_U64", "stat4_I64"], acc(rate="16 ns", startat="25 ns", stopat="60 ns", resetOnOutput=True))]
pairs_event = [(["stat1_U32", "stat3_I32"], acc(rate="9 events")),
               (["stat2_U64", "stat4_I64"], acc(rate="13 events", startat="25 ns", stopat="60 ns", resetOnOutput=True))]
plans = [by_time, by_event, pairs_time, pairs_event]
for i in range(8):
    name = f"StatBasic{i}"
    c = comp(name, "int", 1447 + i, 1053 + i, {"register_dynamic": 30} if i == 0 else None)
    for k, (names, params) in enumerate(plans[i % 4]):
        if i < 4:
            c.enableStatistics(names, params)
        elif len(names) == 1 and k < 2:
            sst.enableStatisticForComponentName(name, names[0], params)
        elif len(names) == 1 and k == 2:
            sst.enableStatisticsForComponentName(name, names[0], params)
        else:
            sst.enableStatisticsForComponentName(name, names, params)
    if i == 0:
        c.enableStatistics(["stat5_dyn"], acc(rate="16 ns"))

# all statistics of a component, with the model's load level, its own, and one set by name
comp("StatBasic8", "int", 1455, 1061).enableAllStatistics(acc(rate="20 ns"))
c = comp("StatBasic9", "int", 1456, 1062)
c.enableAllStatistics(acc(rate="25 ns"))
c.setStatisticLoadLevel(4)
comp("StatBasic10", "int", 1457, 1063)
sst.enableAllStatisticsForComponentName("StatBasic10", acc(rate="25 ns"))
sst.setStatisticLoadLevelForComponentName("StatBasic10", 1)

# by component type
types = [comp(f"StatType{i}", "float", 1457, 1063) for i in range(3)]
sst.enableAllStatisticsForComponentType("coreTestElement.StatisticsComponent.float", acc(rate="20 ns"))
sst.setStatisticLoadLevelForComponentType("coreTestElement.StatisticsComponent.float", 1)
types[2].setStatisticLoadLevel(2)

# statistic objects: one bound to one slot, one shared by two slots
c = comp("StatObjComp0", "float", 1458, 1064)
c.setStatistic("stat2_F64", c.createStatistic("statobj0", acc(rate="17ns")))
c = comp("StatObjComp1", "float", 1459, 1065)
obj = c.createStatistic("statobj1", acc(rate="17ns"))
c.setStatistic("stat2_F64", obj)
c.setStatistic("stat3_F64", obj)

# statistic groups, each with its own dump frequency and output
outputs = [sst.StatisticOutput("sst.statOutputConsole", {"outputsimtime": True, "outputrank": False}),
           sst.StatisticOutput("sst.statOutputCSV", {"filepath": "test_StatisticsComponent_basic_group_stats.csv",
                                                     "separator": ", ", "outputrank": "0"}),
           sst.StatisticOutput("sst.statOutputTXT", {"filepath": "test_StatisticsComponent_basic_group_stats.txt",
                                                     "outputrank": "0"})]
seeds = [(1460, 1066), (1461, 1067), (1462, 1068), (1463, 1069), (1464, 1070), (1466, 1071), (1466, 1072), (1467, 1073),
         (1468, 1074), (1469, 1075), (1470, 1076), (1471, 1077)]
for gi, (freq, out, kind, s1, s2) in enumerate([("19ns", outputs[0], "int", "stat1_U32", "stat2_U64"),
                                                ("23ns", outputs[1], "int", "stat1_U32", "stat2_U64"),
                                                ("27ns", outputs[2], "int", "stat1_U32", "stat2_U64"),
                                                ("29ns", outputs[2], "float", "stat1_F32", "stat2_F64")]):
    grp = sst.StatisticGroup(f"StatGroup{gi}")
    for k in range(3):
        n = 3 * gi + k
        grp.addComponent(comp(f"StatGroupObj{n}", kind, *seeds[n]))
    grp.addStatistic(s1, {"type": ACC})
    grp.addStatistic(s2, {"type": ACC, "resetOnOutput": True})
    grp.setFrequency(freq)
    grp.setOutput(out)

for i in range(len(chain) - 1):
    link = sst.Link(f"link{i}", "1ns")
    chain[i].addLink(link, "left")
    chain[i + 1].addLink(link, "right")

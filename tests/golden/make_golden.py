#!/usr/bin/env python3
"""Generate tests/golden/*.json from (a) the reference's own golden files under /root/reference/tests/refFiles and
(b) runs of the UNMODIFIED reference core built by oracle/build_ref.py (oracle/_ref/libexec/sstsim.x), including our
out-of-tree probe elements (oracle/probe).  Run in the build container (the GPU box has no /root/reference):

    python oracle/build_ref.py && python tests/golden/make_golden.py
"""
import json
import re
import sys
import tempfile
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT / "tests"))
import oracle_lib as O  # noqa: E402

REF_TESTS = Path("/root/reference/tests")
MODELS = ROOT / "tests" / "models"


def mesh_counts(text):
    return {int(m.group(1)): int(m.group(2)) for m in re.finditer(r"^(\d+) received (-?\d+) messages", text, re.M)}


def main():
    out = {}
    # (a) reference golden vectors -------------------------------------------------------------------------------
    out["mesh_6x6_10us"] = {"source": "tests/refFiles/test_MessageMesh.out",
                            "counts": mesh_counts((REF_TESTS / "refFiles/test_MessageMesh.out").read_text())}
    for g in ("mersenne", "marsaglia", "xorshift"):
        lines = (REF_TESTS / f"refFiles/test_RNGComponent_{g}.out").read_text().strip().splitlines()
        out[f"rng_{g}_last5"] = {"source": f"tests/refFiles/test_RNGComponent_{g}.out", "lines": lines[-5:]}
    prof = (REF_TESTS / "refFiles/test_Profiling_event_global.out").read_text()
    out["profiling_4x4_recv_count"] = {"source": "tests/refFiles/test_Profiling_event_global.out",
                                       "recv_count": int(re.search(r"recv_count\D+(\d+)", prof).group(1))}
    disc = (REF_TESTS / "refFiles/test_DistribComponent_discrete.out").read_text()
    out["distrib_discrete"] = {"source": "tests/refFiles/test_DistribComponent_discrete.out (unwired by the "
                                         "reference's testsuites, SURVEY.md section 4)", "text": disc.strip().splitlines()}
    # (b) reference binary runs ----------------------------------------------------------------------------------
    if not O.have_ref():
        raise SystemExit("oracle/_ref is not built; run python oracle/build_ref.py")
    txt = O.run_ref(REF_TESTS / "test_MessageMesh.py", "4 4")
    out["mesh_4x4_10us"] = {"source": "sstsim.x tests/test_MessageMesh.py '4 4'", "counts": mesh_counts(txt)}
    txt = O.run_ref(REF_TESTS / "test_MessageMesh.py", "5 3 3", extra=["--stop-at", "500ns"])
    out["mesh_5x3_mod3_500ns"] = {"source": "sstsim.x test_MessageMesh.py '5 3 3' --stop-at 500ns",
                                  "counts": mesh_counts(txt)}
    txt = O.run_ref(REF_TESTS / "test_MessageMesh.py", "1 2", extra=["--stop-at", "300ns"])
    out["mesh_1x2_300ns"] = {"source": "sstsim.x test_MessageMesh.py '1 2' --stop-at 300ns (x links are loops "
                                       "between two ports of the same component)", "counts": mesh_counts(txt)}
    txt = O.run_ref(REF_TESTS / "test_MessageMesh.py", "2 2", extra=["--stop-at", "300ns"])
    out["mesh_2x2_300ns"] = {"source": "sstsim.x test_MessageMesh.py '2 2' --stop-at 300ns",
                             "counts": mesh_counts(txt)}
    txt = O.run_ref(REF_TESTS / "test_MessageMesh.py", "16 16", extra=["--stop-at", "200ns"], threads=4)
    out["mesh_16x16_200ns_n4"] = {"source": "sstsim.x -n 4 test_MessageMesh.py '16 16' --stop-at 200ns",
                                  "counts": mesh_counts(txt)}
    txt = O.run_ref(MODELS / "phold.py", "8 8 200ns")
    out["phold_8x8_200ns"] = {
        "source": "sstsim.x tests/models/phold.py '8 8 200ns' (pdes.phold probe element)",
        "nodes": {int(m.group(1)): [int(m.group(2)), m.group(3)]
                  for m in re.finditer(r"^phold (\d+) recv (\d+) hash ([0-9a-f]+)", txt, re.M)}}
    txt = O.run_ref(MODELS / "phold.py", "5 4 300ns 7 3000", threads=2)
    out["phold_5x4_300ns_n2"] = {
        "source": "sstsim.x -n 2 tests/models/phold.py '5 4 300ns 7 3000'",
        "nodes": {int(m.group(1)): [int(m.group(2)), m.group(3)]
                  for m in re.finditer(r"^phold (\d+) recv (\d+) hash ([0-9a-f]+)", txt, re.M)}}
    txt = O.run_ref(MODELS / "clockpop.py", "6 5 2us 10")
    out["clockpop_6x5_2us_f10"] = {
        "source": "sstsim.x tests/models/clockpop.py '6 5 2us 10' (pdes.clocknode probe element)",
        "nodes": {int(m.group(1)): [int(m.group(2)), int(m.group(3)), int(m.group(4)), m.group(5)]
                  for m in re.finditer(r"^clocknode (\d+) ticks (\d+) last_cycle (\d+) recv (\d+) hash ([0-9a-f]+)",
                                       txt, re.M)}}
    with tempfile.TemporaryDirectory() as td:
        O.run_ref(MODELS / "phold.py", "3 3 100ns 16 8000 1 1", cwd=td)
        out["phold_3x3_100ns_csv"] = {"source": "sstsim.x phold.py '3 3 100ns 16 8000 1 1' StatisticOutput.csv",
                                      "rows": (Path(td) / "StatisticOutput.csv").read_text().strip().splitlines()}
        O.run_ref(REF_TESTS / "test_MessageMesh.py", "3 3 1 1 2 1", extra=["--stop-at", "100ns"], cwd=td)
        out["mesh_3x3_100ns_csv"] = {"source": "sstsim.x test_MessageMesh.py '3 3 1 1 2 1' --stop-at 100ns",
                                     "rows": (Path(td) / "StatisticOutput.csv").read_text().strip().splitlines()}
        O.run_ref(MODELS / "clockpop.py", "3 3 1us 10 1 1", cwd=td)
        out["clockpop_3x3_1us_csv"] = {"source": "sstsim.x clockpop.py '3 3 1us 10 1 1' StatisticOutput.csv",
                                       "rows": (Path(td) / "StatisticOutput.csv").read_text().strip().splitlines()}
        # per-component delivery order (time, port) captured with the pdes.trace profile tool
        tr = Path(td) / "tr.txt"
        O.run_ref(MODELS / "phold.py", "3 2 40ns 4 3000 0",
                  extra=[f"--enable-profiling=tr:pdes.trace(level=component,track_ports=true,file={tr})[event]"])
        seq = {}
        for line in tr.read_text().splitlines():
            t, prio, tag, q, key = line.split()
            comp, port = key.split(":")
            seq.setdefault(comp, []).append([int(t), int(port.replace("port", ""))])
        out["phold_3x2_40ns_trace"] = {
            "source": "sstsim.x phold.py '3 2 40ns 4 3000 0' + pdes.trace: per-component [time, port] sequences",
            "seq": seq}
    with tempfile.TemporaryDirectory() as td:
        # the reference's JSON interchange format: written by the reference (--output-json), consumed by both
        gj = Path(td) / "g.json"
        O.run_ref(REF_TESTS / "test_MessageMesh.py", "3 2 1 1 1 1", extra=["--stop-at", "40ns", "--output-json", str(gj)],
                  cwd=td)
        (HERE / "mesh_3x2_stats.json").write_text(gj.read_text())
        txt = O.run_ref(gj, cwd=td)  # the reference running its own JSON
        out["mesh_3x2_json_run"] = {
            "source": "sstsim.x test_MessageMesh.py '3 2 1 1 1 1' --stop-at 40ns --output-json; then sstsim.x g.json",
            "counts": mesh_counts(txt), "rows": (Path(td) / "StatisticOutput.csv").read_text().strip().splitlines()}
        # console statistic output
        scr = Path(td) / "console.py"
        scr.write_text((REF_TESTS / "test_MessageMesh.py").read_text().replace(
            'sst.setStatisticOutput("sst.statOutputCSV")', 'sst.setStatisticOutput("sst.statOutputConsole")'))
        txt = O.run_ref(scr, "2 2 1 1 1 1", extra=["--stop-at", "30ns"], cwd=td)
        out["mesh_2x2_console_stats"] = {
            "source": "sstsim.x test_MessageMesh.py (statOutputConsole) '2 2 1 1 1 1' --stop-at 30ns",
            "lines": [l for l in txt.splitlines() if " : Accumulator : " in l]}
    # handler-count profile tools: the reference's own goldens (4x4 mesh, 10 us) and runs of the reference binary
    for lvl in ("global", "type", "component", "subcomponent"):
        out[f"profiling_4x4_event_{lvl}"] = {
            "source": f"tests/refFiles/test_Profiling_event_{lvl}.out",
            "text": (REF_TESTS / f"refFiles/test_Profiling_event_{lvl}.out").read_text()}
    with tempfile.TemporaryDirectory() as td:
        spec = ("events:sst.profile.handler.event.count(level=component)[event];"
                "clocks:sst.profile.handler.clock.count(level=component)[clock]")
        O.run_ref(MODELS / "clockpop.py", "3 2 200ns 10", extra=["--profiling-output=p.txt", "--enable-profiling=" + spec],
                  cwd=td)
        out["profiling_clockpop_3x2_200ns"] = {"source": f"sstsim.x clockpop.py '3 2 200ns 10' --enable-profiling={spec}",
                                               "spec": spec, "text": (Path(td) / "p.txt").read_text()}
        spec = "ev:sst.profile.handler.event.count(level=subcomponent,track_ports=true)[event]"
        O.run_ref(REF_TESTS / "test_MessageMesh.py", "3 2", extra=["--stop-at", "50ns", "--profiling-output=q.txt",
                                                                  "--enable-profiling=" + spec], cwd=td)
        out["profiling_mesh_3x2_50ns_ports"] = {"source": f"sstsim.x test_MessageMesh.py '3 2' --stop-at 50ns --enable-profiling={spec}",
                                                "spec": spec, "text": (Path(td) / "q.txt").read_text()}
        spec = "t:sst.profile.handler.clock.count(level=type)[clock];g:sst.profile.handler.event.count(level=type)[event]"
        O.run_ref(MODELS / "phold.py", "4 3 60ns", extra=["--profiling-output=r.txt", "--enable-profiling=" + spec], cwd=td)
        out["profiling_phold_4x3_60ns_type"] = {"source": f"sstsim.x phold.py '4 3 60ns' --enable-profiling={spec}",
                                                "spec": spec, "text": (Path(td) / "r.txt").read_text()}
    # coreTestElement.coreTestComponent: the reference's own model script and golden file, plus handler counts and
    # statistics of the same model from the reference binary (the golden file itself holds no numbers)
    out["component_golden"] = {"source": "tests/refFiles/test_Component.out",
                               "lines": (REF_TESTS / "refFiles/test_Component.out").read_text().strip().splitlines()}
    with tempfile.TemporaryDirectory() as td:
        spec = ("events:sst.profile.handler.event.count(level=component)[event];"
                "clocks:sst.profile.handler.clock.count(level=component)[clock]")
        txt = O.run_ref(REF_TESTS / "test_Component.py", extra=["--profiling-output=p.txt", "--enable-profiling=" + spec], cwd=td)
        out["component_profile"] = {"source": f"sstsim.x tests/test_Component.py --enable-profiling={spec}", "spec": spec,
                                    "stdout": txt.strip().splitlines(), "text": (Path(td) / "p.txt").read_text()}
        O.run_ref(MODELS / "component.py", "10 10 37 3us 1", cwd=td)
        out["component_stats_csv"] = {"source": "sstsim.x tests/models/component.py '10 10 37 3us 1' (the test_Component.py "
                                                "graph with commFreq 37, 3 us, all statistics enabled)",
                                      "rows": (Path(td) / "StatisticOutput.csv").read_text().strip().splitlines()}
    # periodic statistic output ("rate"): the reference's own script option, and our probe elements where event times do
    # not line up with the lookahead windows
    with tempfile.TemporaryDirectory() as td:
        for key, script, opts, extra in [
                ("rate_mesh_2x2_600ns", REF_TESTS / "test_MessageMesh.py", "2 2 1 1 2 2", ["--stop-at", "600ns"]),
                ("rate_phold_3x3_100ns", MODELS / "phold.py", "3 3 100ns 16 8000 1 1 17ns", []),
                ("rate_clockpop_3x3_1us", MODELS / "clockpop.py", "3 3 1us 10 1 1 130ns", [])]:
            O.run_ref(script, opts, extra=extra, cwd=td)
            out[key] = {"source": f"sstsim.x {Path(script).name} '{opts}' {' '.join(extra)} StatisticOutput.csv",
                        "opts": opts, "rows": (Path(td) / "StatisticOutput.csv").read_text().strip().splitlines()}
    # larger parity sizes, kept as digests: the threaded reference on a 128 x 128 mesh
    # and on a 48 x 40 PHOLD torus
    import hashlib
    import numpy as np
    txt = O.run_ref(REF_TESTS / "test_MessageMesh.py", "128 128", extra=["--stop-at=150ns"], threads=8)
    counts = mesh_counts(txt)
    arr = np.array([counts[i] for i in range(128 * 128)], dtype="<i4")
    out["mesh_128x128_150ns_n8_digest"] = {"source": "sstsim.x -n 8 test_MessageMesh.py '128 128' --stop-at=150ns",
                                           "sum": int(arr.sum()), "sha256": hashlib.sha256(arr.tobytes()).hexdigest()}
    txt = O.run_ref(MODELS / "phold.py", "48 40 250ns", threads=4)
    rows = sorted((int(m.group(1)), int(m.group(2)), m.group(3)) for m in
                  re.finditer(r"^phold (\d+) recv (\d+) hash ([0-9a-f]+)", txt, re.M))
    blob = "\n".join(f"{i} {r} {h}" for i, r, h in rows).encode()
    out["phold_48x40_250ns_n4_digest"] = {"source": "sstsim.x -n 4 tests/models/phold.py '48 40 250ns'",
                                          "components": len(rows), "recv_sum": sum(r for _, r, _ in rows),
                                          "sha256": hashlib.sha256(blob).hexdigest()}
    # long streams of the reference's own generator classes (oracle/tools/rng_ref_driver.cc): the seeds and lengths
    # tests/test_gpu_rng.py draws on the device, and 2^24 draws per generator in blocks of 2^20
    import subprocess
    drv = ROOT / "oracle" / "_ref" / "libexec" / "rng_ref_driver"
    streams = []
    base = {0: (1447, 0), 1: (1053, 1447), 2: (1447, 0)}
    for kind, (s1, s2) in base.items():
        cases = [(s1 + 3 * j, s2 + 5 * j, 5000) for j in (0, 1, 17, 63)] + \
                [(s1 + j, s2 + 2 * j, 1 << 14) for j in (0, 65535, 12345)] + [(s1, s2, 1 << 24)]
        for a, b, n in cases:
            txt = subprocess.run([str(drv), str(kind), str(a), str(b), str(n)], capture_output=True, text=True, check=True).stdout
            tot = re.search(r"total (\d+) first (\d+) last (\d+)", txt)
            streams.append({"kind": kind, "s1": a, "s2": b, "n": n, "total": int(tot.group(1)), "first": int(tot.group(2)),
                            "last": int(tot.group(3)), "blocks": [int(x) for x in re.findall(r"block \d+ (\d+)", txt)]})
        txt = subprocess.run([str(drv), str(kind), str(s1), str(s2), str(100000), "u64"], capture_output=True, text=True, check=True).stdout
        tot = re.search(r"total (\d+) first (\d+) last (\d+)", txt)
        streams.append({"kind": kind, "s1": s1, "s2": s2, "n": 100000, "u64": True, "total": int(tot.group(1)),
                        "first": int(tot.group(2)), "last": int(tot.group(3)), "blocks": []})
    out["rng_reference_class_streams"] = {"source": "oracle/tools/rng_ref_driver.cc linked with the reference's rng/*.o: "
                                                    "sum_i (i+1)*draw_i mod 2^64 (total and per 2^20-draw block)",
                                          "streams": streams}
    # the reference's distribution classes over its generators: 400 values each, as bit patterns of the doubles
    dists = []
    for d, params in [(0, [0.1, 0.3, 0.35, 0.15, 0.1]), (1, [1.25]), (2, [10.0, 2.5]), (3, [3.5]), (4, [7]), (5, [42.0])]:
        for kind, (s1, s2) in base.items():
            txt = subprocess.run([str(drv), "dist", str(d), str(kind), str(s1), str(s2), "400", *map(str, params)],
                                 capture_output=True, text=True, check=True).stdout
            dists.append({"dist": d, "params": params, "kind": kind, "s1": s1, "s2": s2, "bits": [int(x) for x in txt.split()]})
    out["rng_reference_class_distributions"] = {"source": "oracle/tools/rng_ref_driver.cc dist ...: SST::RNG::*Distribution "
                                                          "(rng/discrete.h, expon.h, gaussian.h, poisson.h, uniform.h, constant.h)",
                                                "cases": dists}
    # statistics enabled on some components only (by object, by name, by type): only those are written, and field
    # columns of both value types share one CSV
    with tempfile.TemporaryDirectory() as td:
        txt = O.run_ref(MODELS / "partial_stats.py", cwd=td)
        out["partial_stats_csv"] = {"source": "sstsim.x tests/models/partial_stats.py StatisticOutput.csv",
                                    "rows": (Path(td) / "StatisticOutput.csv").read_text().strip().splitlines()}
    # irregular random models (mixed element types, 1..6 ports, different latencies at the two ends of a link): the
    # PHOLD delivery hash is order-sensitive, so these pin per-component event ORDER on irregular graphs
    with tempfile.TemporaryDirectory() as td:
        cases = []
        for opts in ("1 40 400ns", "2 40 400ns", "3 40 400ns", "6 17 900ns", "6 2 20ns", "4 25 333333ps", "9 33 100001ps"):
            txt = O.run_ref(MODELS / "random_graph.py", opts, cwd=td)
            cases.append({"opts": opts,
                          "lines": sorted(l for l in txt.splitlines() if l.startswith(("phold", "clocknode", "Simulation is"))),
                          "rows": (Path(td) / "StatisticOutput.csv").read_text().strip().splitlines()})
        out["random_graphs"] = {"source": "sstsim.x tests/models/random_graph.py '<seed> <components> <stop>'", "cases": cases}
    (HERE / "reference_golden.json").write_text(json.dumps(out, indent=1, sort_keys=True))
    print("wrote", HERE / "reference_golden.json", (HERE / "reference_golden.json").stat().st_size, "bytes")


if __name__ == "__main__":
    main()

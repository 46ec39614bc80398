#!/usr/bin/env python3
"""Golden files of the statistics row -> tests/golden/stats_golden.json: the reference's own
tests/refFiles/test_StatisticsComponent_basic.out and the two group files its test writes, plus a run of the reference
binary (oracle/_ref) on the same script as a cross-check.  Needs /root/reference; run from the repo root."""
import json
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import oracle_lib as O  # noqa: E402

REF = Path("/root/reference/tests")
keep = lambda text: [l for l in text.splitlines() if not l.startswith("#")]
out = {"basic_out": {"source": "tests/refFiles/test_StatisticsComponent_basic.out",
                     "lines": keep((REF / "refFiles/test_StatisticsComponent_basic.out").read_text())}}
for ext in ("csv", "txt"):
    f = f"test_StatisticsComponent_basic_group_stats.{ext}"
    out["basic_" + ext] = {"source": "tests/refFiles/" + f, "file": f, "lines": (REF / "refFiles" / f).read_text().splitlines()}
with tempfile.TemporaryDirectory() as td:
    txt = O.run_ref(REF / "test_StatisticsComponent_basic.py", cwd=td)
    out["basic_ref_run"] = {"source": "sstsim.x tests/test_StatisticsComponent_basic.py", "lines": keep(txt)}
    for ext in ("csv", "txt"):
        f = f"test_StatisticsComponent_basic_group_stats.{ext}"
        assert (Path(td) / f).read_text().splitlines() == out["basic_" + ext]["lines"], f
assert sorted(out["basic_ref_run"]["lines"]) == sorted(out["basic_out"]["lines"])
(ROOT / "tests" / "golden" / "stats_golden.json").write_text(json.dumps(out, indent=1) + "\n")
print({k: len(v["lines"]) for k, v in out.items()})

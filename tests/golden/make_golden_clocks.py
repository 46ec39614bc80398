#!/usr/bin/env python3
"""Golden vectors for the clock rows -> tests/golden/clocks_golden.json:
  * tests/refFiles/test_Clocks_basic.out, the reference's own golden file of tests/test_Clocks.py;
  * console output of the reference binary (oracle/_ref) on tests/models/clocks.py for other chain lengths and link
    latencies (lookahead windows that cut the 5 ns / 1 ns clock pattern differently).
Needs /root/reference and the hand-built reference (oracle/build_ref.py); run from the repo root."""
import json
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import oracle_lib as O  # noqa: E402

REF_TESTS = Path("/root/reference/tests")
MODELS = ROOT / "tests" / "models"


def lines(text):
    return [l for l in text.strip().splitlines() if not l.startswith("#") and not l.startswith("WARNING")]


out = {"clocks_basic": {"source": "tests/refFiles/test_Clocks_basic.out",
                        "lines": lines((REF_TESTS / "refFiles/test_Clocks_basic.out").read_text())}}
with tempfile.TemporaryDirectory() as td:
    txt = O.run_ref(REF_TESTS / "test_Clocks.py", cwd=td)
    out["clocks_ref_run"] = {"source": "sstsim.x tests/test_Clocks.py", "lines": lines(txt)}
    for opts in ("3 2ns", "5 13ns", "2 1ns", "6 7500ps", "1 6ns"):
        txt = O.run_ref(MODELS / "clocks.py", opts, cwd=td)
        out["clocks_" + opts.replace(" ", "_")] = {"source": f"sstsim.x tests/models/clocks.py '{opts}'", "opts": opts,
                                                   "lines": lines(txt)}
(ROOT / "tests" / "golden" / "clocks_golden.json").write_text(json.dumps(out, indent=1) + "\n")
print({k: len(v["lines"]) for k, v in out.items()})

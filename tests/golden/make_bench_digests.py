#!/usr/bin/env python3
"""Digests of the benchmarked MessageMesh configurations from the CPU golden model (oracle/mesh_golden.cc, pinned to
pdes_oracle.cc and through it to the reference in tests/test_oracle.py) -> tests/golden/bench_digests.json.
bench.py prints the digest of the result records it downloads as `parity.digest` and compares it with these; the
GPU tests do the same at full size.  Usage: python tests/golden/make_bench_digests.py   (about a minute on 8 cores)"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import oracle_lib as O  # noqa: E402
from sst_core_b200.digest import mesh_records_from_counts, records_digest  # noqa: E402

out = {"_doc": "key mesh_<X>x<Y>_w<windows>_stats<0|1>: records_digest(ids, result records[:, :6]) of MessageMesh X x Y "
               "(tests/models/mesh.py semantics, mod 1) after <windows> lookahead windows of 1 ns; events = events delivered",
       "digests": {}}
for size in (256, 1024, 2048, 4096):
    for windows in (8, 13, 25, 45, 100):
        counts, events = O.mesh_golden_counts(size, size, windows)
        ids = np.arange(size * size, dtype=np.uint64)
        for stats in (0, 1):
            words = 6 if stats else 1
            d = records_digest(ids, mesh_records_from_counts(counts, bool(stats))[:, :words])
            out["digests"][f"mesh_{size}x{size}_w{windows}_stats{stats}"] = {"digest": f"{d:016x}", "events": events}
        print(size, windows, events, flush=True)
# PHOLD at the bench's size: the oracle itself (no count-only shortcut exists for it); --phold, about four minutes
if "--phold" in sys.argv:
    from sst_core_b200 import models  # noqa: E402
    from sst_core_b200.run import result_words  # noqa: E402
    m = models.phold_model(1024, 1024, stop_at="25ns", verbose=0, stats_enabled=True)
    ref = O.run_model(m)
    d = records_digest(np.arange(m.ncomp, dtype=np.uint64), ref.results[:, :result_words(m)])
    out["digests"]["phold_1024x1024_w25"] = {"digest": f"{d:016x}", "events": int(ref.events)}
else:
    old = json.loads((ROOT / "tests" / "golden" / "bench_digests.json").read_text())["digests"]
    out["digests"].update({k: v for k, v in old.items() if k.startswith(("phold_", "clockpop_"))})
(ROOT / "tests" / "golden" / "bench_digests.json").write_text(json.dumps(out, indent=1) + "\n")

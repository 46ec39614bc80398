#!/usr/bin/env python3
"""First-contact check of the event-parallel window kernel: small meshes against the oracle (run under `timeout`)."""
import sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import oracle_lib as O
from sst_core_b200 import models
from sst_core_b200.run import simulate_model

for (x, y, ns) in [(16, 16, 20), (64, 64, 60), (37, 29, 200), (256, 128, 400), (512, 512, 30)]:
    m = models.mesh_model(x, y, stop_at=f"{ns}ns", stats_enabled=True)
    t0 = time.perf_counter()
    res, st = simulate_model(m, device=0)
    dt = time.perf_counter() - t0
    ref = O.run_model(m)
    ok = np.array_equal(res, ref.results) and st.events_delivered == ref.events
    print(f"mesh {x}x{y} {ns}ns: {'OK' if ok else 'MISMATCH'} events {st.events_delivered} vs {ref.events} "
          f"windows {st.windows} launches {st.launches} {dt*1e3:.0f} ms", flush=True)
    if not ok:
        bad = np.nonzero((res != ref.results).any(axis=1))[0]
        print("  first bad components", bad[:10], res[bad[:3], :6], ref.results[bad[:3], :6])
        sys.exit(1)

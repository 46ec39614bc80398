"""Small simulations of every element for compute-sanitizer (one tool per gpurun call)."""
import sys
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import numpy as np
import oracle_lib as O
from sst_core_b200 import models
from sst_core_b200.run import simulate_model
from random_models import random_model

cases = [(models.mesh_model(20, 13, stop_at="12ns", stats_enabled=True), {}),
         (models.phold_model(18, 18, stop_at="14ns", stats_enabled=True), {"calendar_slots": 16}),
         (models.clockpop_model(12, 12, stop_at="20ns", comm_freq=3, stats_enabled=True), {}),
         (random_model(5, ncomp=300, stop_at=12_000), {"calendar_slots": 4, "bin_capacity": 4096, "smem_events": 4096})]
for m, opts in cases:
    res, st = simulate_model(m, auto_tune=False, **opts)
    ref = O.run_model(m)
    assert np.array_equal(res, ref.results) and st.events_delivered == ref.events
    print("ok", st.events_delivered, "events", st.windows, "windows", flush=True)

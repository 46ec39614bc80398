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
# the three forms with their per-block fallbacks: a staging buffer shorter than the bins, bursts, busy clock populations
cases += [(models.phold_model(40, 30, stop_at="20ns", stats_enabled=True), {"calendar_slots": 16, "_env": {"SSTB200_PF_RCAP": "96"}}),
          (models.phold_model(8, 8, stop_at="20ns", init_events=50, delay_mod=1500, stats_enabled=True),
           {"calendar_slots": 16, "bin_capacity": 8192, "smem_events": 4096}),
          (models.clockpop_model(40, 30, stop_at="40ns", comm_freq=2, stats_enabled=True), {"calendar_slots": 4}),
          (models.clockpop_model(40, 30, stop_at="40500ps", comm_freq=40, stats_enabled=True), {})]
import os
for m, opts in cases:
    env = opts.pop("_env", {})
    os.environ.update(env)
    res, st = simulate_model(m, auto_tune=False, **opts)
    for k in env:
        del os.environ[k]
    ref = O.run_model(m)
    assert np.array_equal(res, ref.results) and st.events_delivered == ref.events
    print("ok", st.events_delivered, "events", st.windows, "windows", flush=True)

#!/usr/bin/env python3
"""Where the end-to-end time of a short run goes (MessageMesh 4096^2, 25 windows, one B200)."""
import sys
import time

import torch

sys.path.insert(0, ".")
from sst_core_b200 import models  # noqa: E402
from sst_core_b200.engine import Engine  # noqa: E402
from sst_core_b200.run import result_words  # noqa: E402

x = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 4096
t0 = time.perf_counter()
m = models.mesh_model(x, x, stop_at="100ns", verbose=0, stats_enabled=True)
t1 = time.perf_counter()
print(f"numpy model build        {t1 - t0:7.3f} s")
if "--pin" in sys.argv:
    from sst_core_b200.wireup import pinned_empty
    import numpy as np
    m.pin()
    buf = pinned_empty((m.ncomp, result_words(m)), np.uint64)
    print(f"pin model + result buffer {time.perf_counter() - t1:6.3f} s")
else:
    buf = None
for rep in range(2):
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    e = Engine(m, calendar_slots=2, bin_capacity=2048, smem_events=1280)
    e.sync(); t2 = time.perf_counter()
    e.setup(); e.sync(); t3 = time.perf_counter()
    e.run(max_windows=25, check_every=25); e.sync(); t4 = time.perf_counter()
    res = e.results(result_words(m), compact=True, out=buf); t5 = time.perf_counter()
    e.close(); t6 = time.perf_counter()
    print(f"rep {rep}: create {t2 - t1:6.3f}  setup {t3 - t2:6.3f}  25 windows {t4 - t3:6.3f}  results {t5 - t4:6.3f}  "
          f"destroy {t6 - t5:6.3f}  total {t5 - t1:6.3f} s")

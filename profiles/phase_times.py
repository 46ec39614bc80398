#!/usr/bin/env python3
"""Per-phase wall time inside dispatch_kernel CTAs (needs a build with -DSSTB200_TIMING).  Prints mean durations."""
import ctypes as C
import sys

import numpy as np

sys.path.insert(0, ".")
from sst_core_b200 import engine, models  # noqa: E402

x = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
m = models.mesh_model(x, x, stop_at="100ns", verbose=0, stats_enabled=True)
e = engine.Engine(m, calendar_slots=2, bin_capacity=2048, smem_events=1536)
e.setup()
e.run(max_windows=8, check_every=8)
e.status()
nb = min((x * x + 255) // 256, 1 << 17)
buf = np.zeros(nb * 8, dtype=np.uint64)
L = engine.lib()
L.sstb200_debug_phase_times.argtypes = [C.c_void_p, C.c_uint32]
assert L.sstb200_debug_phase_times(buf.ctypes.data_as(C.c_void_p), nb * 8) == 0
t = buf.reshape(nb, 8).astype(np.int64)
names = ["prologue (loads issued -> barrier)", "link cp.async issue + phase 1 (bin load, classify)", "phase 2 (scan, scatter, buckets)",
         "phase 3 (handlers) until the LAST warp is done", "phase 4 (flush)"]
d = np.stack([t[:, i + 1] - t[:, i] for i in range(5)], axis=1)
print("CTAs", nb, "kernel span %.1f us" % ((t[:, 5].max() - t[:, 0].min()) / 1e3))
for i, n in enumerate(names):
    print(f"{n:58s} mean {d[:, i].mean()/1e3:7.2f} us   p50 {np.median(d[:, i])/1e3:7.2f}   p95 {np.percentile(d[:, i],95)/1e3:7.2f}")
print(f"{'phase 3 until the LIGHTEST warp (thread 255) is done':58s} mean {(t[:,7]-t[:,3]).mean()/1e3:7.2f} us")
print(f"{'phase 3a (per component) until the barrier':58s} mean {(t[:,6]-t[:,3]).mean()/1e3:7.2f} us")
print(f"{'phase 3b (event-parallel handlers + deferred writes)':58s} mean {(t[:,4]-t[:,6]).mean()/1e3:7.2f} us")
print(f"{'CTA lifetime':58s} mean {(t[:,5]-t[:,0]).mean()/1e3:7.2f} us")
e.close()

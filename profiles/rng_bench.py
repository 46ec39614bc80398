#!/usr/bin/env python3
"""BASELINE config 2 throughput: 2^30 generateNextUInt32 draws per generator on one B200 (2^16 independent
generators x 2^14 draws, seeds s+j), checksum-only and with the draws written to HBM.  CUDA-event timed."""
import json
import sys

import torch

sys.path.insert(0, ".")
from sst_core_b200 import engine  # noqa: E402

dev = torch.device("cuda", 0)
out = []
for kind, name, s1, s2 in [(0, "mersenne", 1447, 0), (1, "marsaglia", 1053, 1447), (2, "xorshift", 1447, 0)]:
    # 2^30 draws as 2^16 generators x 2^14 draws (one lane / one warp per generator: 443 lanes per SM for the small
    # generators) and as 2^20 x 2^10 (every SM full)
    for n_streams, draws, write in [(1 << 16, 1 << 14, False), (1 << 16, 1 << 14, True), (1 << 20, 1 << 10, False),
                                    (1 << 20, 1 << 10, True)]:
        buf = torch.empty((n_streams, draws), dtype=torch.int32, device=dev) if write else None
        cs = torch.zeros(n_streams, dtype=torch.int64, device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        for it in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            engine._check(engine.lib().sstb200_rng_streams(0, st, kind, s1, s2, 1, 2, n_streams, draws,
                                                           buf.data_ptr() if write else None, cs.data_ptr()), "rng")
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b)
        out.append({"generator": name, "generators": n_streams, "draws": n_streams * draws, "written_to_hbm": write, "ms": ms,
                    "gdraws_per_s": n_streams * draws / ms / 1e6, "gb_per_s_written": (4 * n_streams * draws / ms / 1e6) if write else 0})
        del buf
print(json.dumps(out, indent=1))

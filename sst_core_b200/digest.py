"""Order-independent 64-bit digest of per-component result records: what bench.py prints as `parity.digest` and what
tests/golden/bench_digests.json holds for the benchmarked configurations (computed there from the CPU golden model).
The digest of a partitioned run is the wrap-around sum of the ranks' digests."""
from __future__ import annotations

import numpy as np

_K0 = np.uint64(0x9E3779B97F4A7C15)
_K1 = np.uint64(0xBF58476D1CE4E5B9)


def records_digest(ids: np.ndarray, records: np.ndarray) -> int:
    """sum over components of mix(id, words) mod 2^64.  ids: global component ids [n]; records: uint64 [n, w]."""
    rec = np.ascontiguousarray(records, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = (np.asarray(ids, dtype=np.uint64) + np.uint64(1)) * _K0
        for i in range(rec.shape[1]):
            x = (x ^ rec[:, i]) * _K1
            x ^= x >> np.uint64(31)
        return int(x.sum(dtype=np.uint64))


def mesh_records_from_counts(counts: np.ndarray, stats_enabled: bool = True) -> np.ndarray:
    """Result records [n, 6] of MessageMesh components that received `counts` events: message_count, then the
    msg_count Accumulator (Sum, SumSQ, Count, Min, Max) = (n, n, n, 1, 1) after n x addData(1)."""
    c = np.asarray(counts).astype(np.int64).astype(np.uint64)
    rec = np.zeros((c.size, 6), dtype=np.uint64)
    rec[:, 0] = c
    if stats_enabled:
        rec[:, 1] = rec[:, 2] = rec[:, 3] = c
        rec[:, 4] = rec[:, 5] = (c > 0).astype(np.uint64)
    return rec

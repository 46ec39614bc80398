"""BASELINE.json's full sizes on one B200, checked through size-independent properties (the oracle cannot run them
in seconds): conservation of events, checksum of checksums between differently-partitioned runs, determinism."""
import numpy as np
import pytest

from sst_core_b200 import models

pytestmark = pytest.mark.gpu


def _run(m, windows, **opts):
    from sst_core_b200.engine import Engine
    e = Engine(m, **opts)
    try:
        e.setup()
        e.run(max_windows=windows, check_every=windows)
        st = e.status()
        e.raise_on_error(st)
        return e.results(8), st
    finally:
        e.close()


def test_mesh_4096x4096_conservation():
    # 16.7 M components, 67.1 M events in flight: every delivered event is forwarded exactly once, so every window
    # after the first delivers exactly 4N events; counts and the msg_count statistic must add up to that
    n = 4096 * 4096
    m = models.mesh_model(4096, 4096, stop_at="1us", verbose=0, stats_enabled=True)
    res, st = _run(m, 5, calendar_slots=2, bin_capacity=2048, smem_events=1280)
    assert st.windows == 5 and st.events_delivered == 4 * 4 * n  # windows 1..4 deliver, window 0 has nothing due
    assert int(res[:, 0].sum()) == st.events_delivered
    assert int(res[:, 1].sum()) == st.events_delivered and int(res[:, 3].sum()) == st.events_delivered
    assert st.events_sent == st.events_delivered + 4 * n
    assert res[:, 0].max() < 64  # Poisson(4) per window x 4 windows: a runaway hot spot means a routing bug


def test_mesh_1024x1024_partitioning_invariance_checksum():
    # the same model as 1 partition and as 5 partitions on one GPU: identical per-component results
    from sst_core_b200.engine import LocalMultiRankRunner
    m = models.mesh_model(1024, 1024, stop_at="12ns", verbose=0, stats_enabled=True)
    res1, st1 = _run(m, 0)
    r = LocalMultiRankRunner(models.mesh_model(1024, 1024, stop_at="12ns", verbose=0, stats_enabled=True), 5)
    try:
        r.setup()
        assert r.run()
        res5 = np.zeros_like(res1)
        res5[r.pm.orig_id, :8] = r.results()[:, :8]
    finally:
        r.close()
    assert np.array_equal(res1[:, :6], res5[:, :6])
    assert st1.events_delivered == 11 * 4 * 1024 * 1024


def test_phold_1m_components_conservation_and_determinism():
    # BASELINE config 3: 1 M components x 16 events.  Events are never created or destroyed after setup, so
    # sent - delivered stays 16 N; two runs give bit-identical order-sensitive hashes.
    n = 1024 * 1024
    m = models.phold_model(1024, 1024, stop_at="1us", verbose=0, stats_enabled=True)
    a, sa = _run(m, 12, calendar_slots=16, bin_capacity=2560, smem_events=2560)
    b, sb = _run(m, 12, calendar_slots=16, bin_capacity=2560, smem_events=2560)
    assert sa.events_sent - sa.events_delivered == 16 * n
    assert int(a[:, 0].sum()) == sa.events_delivered and int(a[:, 4].sum()) == sa.events_delivered
    assert np.array_equal(a, b) and sa.events_delivered == sb.events_delivered


def test_phold_512x512_vs_oracle():
    import oracle_lib as O
    m = models.phold_model(512, 512, stop_at="20ns", verbose=0, stats_enabled=True)
    res, st = _run(m, 0, calendar_slots=16, bin_capacity=2560, smem_events=2560)
    ref = O.run_model(m)
    assert st.events_delivered == ref.events
    assert np.array_equal(res[:, :7], ref.results[:, :7])


# ---- the benchmarked configurations against the oracle, record for record (256^2 / 512^2 /
# 1024^2 meshes vs the oracle; bench.py's own engine options)
BENCH_MESH = dict(calendar_slots=2, bin_capacity=2048, smem_events=1280, outbox_capacity=16384)
BENCH_PHOLD = dict(calendar_slots=16, bin_capacity=2560, smem_events=1280, outbox_capacity=16384)


@pytest.mark.parametrize("size,ns", [(256, 40), (512, 16), (1024, 13)])
def test_mesh_at_bench_options_vs_oracle(size, ns):
    import oracle_lib as O
    m = models.mesh_model(size, size, stop_at=f"{ns}ns", verbose=0, stats_enabled=True)
    res, st = _run(m, 0, **BENCH_MESH)
    ref = O.run_model(m)
    assert st.events_delivered == ref.events == (ns - 1) * 4 * size * size and st.end_time == ref.end_time
    assert np.array_equal(res[:, :8], ref.results[:, :8])


def test_mesh_1024_long_run_crosses_mt19937_generations_vs_oracle():
    # 180 windows = ~1440 draws per component: every generator is rebuilt at least twice (624 words per generation), so
    # the generation maintenance of the event-parallel form (next generation, digest tail) is covered at scale
    import oracle_lib as O
    m = models.mesh_model(256, 512, stop_at="181ns", verbose=0, stats_enabled=True)
    res, st = _run(m, 0, **BENCH_MESH)
    ref = O.run_model(m)
    assert st.events_delivered == ref.events
    assert np.array_equal(res[:, :8], ref.results[:, :8])


def test_phold_1024x1024_at_bench_options_vs_oracle():
    import oracle_lib as O
    m = models.phold_model(1024, 1024, stop_at="12ns", verbose=0, stats_enabled=True)
    res, st = _run(m, 0, **BENCH_PHOLD)
    ref = O.run_model(m)
    assert st.events_delivered == ref.events
    assert np.array_equal(res[:, :7], ref.results[:, :7])


def test_mesh_4096x4096_25_windows_digest_equals_cpu_golden_model():
    # BASELINE config 4 exactly as bench.py runs it (25 windows): the digest of all 16.7 M result records against
    # tests/golden/bench_digests.json, computed on the CPU by oracle/mesh_golden.cc (tests/golden/make_bench_digests.py)
    import json
    from pathlib import Path
    from sst_core_b200.digest import records_digest
    fix = json.loads((Path(__file__).parent / "golden" / "bench_digests.json").read_text())["digests"]
    n = 4096 * 4096
    m = models.mesh_model(4096, 4096, stop_at="28ns", verbose=0, stats_enabled=True)
    res, st = _run(m, 25, **BENCH_MESH)
    want = fix["mesh_4096x4096_w25_stats1"]
    assert st.events_delivered == want["events"] == 24 * 4 * n
    assert f"{records_digest(np.arange(n), res[:, :6]):016x}" == want["digest"]


def test_clock_population_with_several_blocks_per_cta_equals_the_sequential_form(monkeypatch):
    # 2048 blocks on ~300 persistent CTAs: the warps of a CTA run their blocks without a barrier between them and drift
    # apart (clock_form.cuh) -- a bin cleared by one warp before another one had fetched its fill lost that bin's events
    # (seen at the bench's size only: every smaller test has at most one block per CTA).  The sequential form, which
    # the other tests pin to the oracle, is the reference here (the oracle needs minutes at this size).
    from sst_core_b200.run import simulate_model
    m = models.clockpop_model(1024, 512, stop_at="40ns", comm_freq=60, verbose=0, stats_enabled=True)
    opts = dict(calendar_slots=2, bin_capacity=512, smem_events=512)
    monkeypatch.setenv("SSTB200_NO_PARALLEL_FORM", "1")
    want, st0 = simulate_model(m, **opts)
    monkeypatch.delenv("SSTB200_NO_PARALLEL_FORM")
    for _ in range(3):
        res, st = simulate_model(m, **opts)
        assert np.array_equal(res, want)
        assert st.events_delivered == st0.events_delivered and st.clock_ticks == st0.clock_ticks

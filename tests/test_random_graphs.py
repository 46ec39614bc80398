"""Irregular random models: (CPU) the oracle gives the same result however the model is partitioned and renumbered;
(GPU) the CUDA engine equals the oracle on them -- mixed element types in one block, non-uniform port counts (binary
search path), loopback links, more distinct destination bins than the flush hash table holds, events far beyond the
calendar horizon."""
import numpy as np
import pytest

import oracle_lib as O
from random_models import random_model


@pytest.mark.parametrize("seed", [1, 2])
def test_oracle_partition_invariance_on_random_graphs(seed):
    from sst_core_b200 import engine
    m = random_model(seed, ncomp=120, stop_at=40_000)
    base = O.run_model(m).results
    pm = engine.partition_model(m, 4)
    got = O.run_model(pm).results
    assert np.array_equal(got, base[pm.orig_id])


def _gpu_vs_oracle(m, **opts):
    from sst_core_b200.run import simulate_model
    ref = O.run_model(m)
    res, st = simulate_model(m, **opts)
    bad = np.nonzero((res != ref.results).any(axis=1))[0]
    assert bad.size == 0, f"{bad.size} components differ, first {bad[:5]} type {m.comp_type[bad[:5]]}"
    assert st.events_delivered == ref.events and st.clock_ticks == ref.ticks and st.end_time == ref.end_time


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [11, 12, 13, 14])
def test_gpu_random_graph_vs_oracle(seed):
    _gpu_vs_oracle(random_model(seed, ncomp=700, stop_at=60_000), calendar_slots=4, bin_capacity=8192, smem_events=6144)


@pytest.mark.gpu
def test_gpu_many_destination_bins_per_block():
    # 20k components = 79 blocks, links pair far-apart components: a block sends to > 64 distinct bins per window
    _gpu_vs_oracle(random_model(21, ncomp=20_000, stop_at=25_000, far=True, p_clock=0.1), calendar_slots=8,
                   bin_capacity=4096, smem_events=4096)


@pytest.mark.gpu
def test_gpu_events_far_beyond_the_calendar_horizon():
    # delays up to 12 windows with a 2-slot calendar: most events are parked and re-binned several times
    _gpu_vs_oracle(random_model(31, ncomp=500, stop_at=80_000, latencies=(1000,)), calendar_slots=2,
                   bin_capacity=16384, smem_events=4096)


@pytest.mark.gpu
@pytest.mark.parametrize("n_ranks", [2, 5])
def test_gpu_random_graph_partitioned(n_ranks):
    from sst_core_b200.engine import LocalMultiRankRunner
    m = random_model(41, ncomp=900, stop_at=50_000, far=True)
    ref = O.run_model(m)
    r = LocalMultiRankRunner(m, n_ranks, calendar_slots=4, bin_capacity=8192, smem_events=6144)
    try:
        r.setup()
        assert r.run()
        res = r.results()
        sts = [e.status() for e in r.engines]
    finally:
        r.close()
    got = np.zeros_like(res)
    got[r.pm.orig_id] = res
    assert np.array_equal(got, ref.results)
    assert sum(s.events_delivered for s in sts) == ref.events and sum(s.clock_ticks for s in sts) == ref.ticks


@pytest.mark.gpu
def test_gpu_error_reports():
    from sst_core_b200.engine import EngineError
    from sst_core_b200.run import simulate_model
    from sst_core_b200 import models
    with pytest.raises(EngineError, match="staging overflow"):
        simulate_model(models.phold_model(16, 16, stop_at="50ns", init_events=40, delay_mod=1), calendar_slots=4,
                       bin_capacity=32768, smem_events=512, auto_tune=False)
    with pytest.raises(EngineError, match="window larger than a link latency|lookahead"):
        simulate_model(models.mesh_model(4, 4, stop_at="20ns"), window=5000)


@pytest.mark.gpu
def test_gpu_mesh_sequential_fallbacks_inside_the_event_parallel_kernel():
    # MessageMesh normally takes the event-parallel form.  (1) cap the events one component may have in that form at
    # 2: every block with a busier component falls back to the sequential form; (2) two links per port group: the
    # element refuses the parallel form (its second draw then selects a link).  Results must not change.
    from sst_core_b200 import models
    from sst_core_b200.run import simulate_model
    m = models.mesh_model(40, 30, stop_at="40ns", stats_enabled=True)
    ref = O.run_model(m)
    for opts in ({}, {"parallel_max_events": 2}, {"parallel_max_events": 5}):
        res, st = simulate_model(m, **opts)
        assert np.array_equal(res, ref.results), opts
        assert st.events_delivered == ref.events
    m2 = models.mesh_model(24, 24, stop_at="40ns", stats_enabled=True)
    m2.comp_param[:, 4] = 2       # two port groups ...
    m2.comp_param[:, 5] = 0x0202  # ... of two links each
    _gpu_vs_oracle(m2)


@pytest.mark.gpu
def test_gpu_mesh_trace_keeps_the_full_delivery_order():
    # with a delivery trace requested the (time, port, sequence) order is kept even for tie-interchangeable elements
    from sst_core_b200 import models
    from sst_core_b200.engine import Engine
    m = models.mesh_model(5, 4, stop_at="12ns")
    ref = O.run_model(m, trace=True)
    e = Engine(m, trace_capacity=1 << 16)
    e.setup()
    e.run()
    tr = e.trace()
    e.close()
    for c in range(20):
        sel, rsel = tr["comp"] == c, ref.trace["comp"] == c
        assert np.array_equal(tr["time"][sel], ref.trace["time"][rsel])
        assert np.array_equal(tr["port"][sel], ref.trace["port"][rsel]), c

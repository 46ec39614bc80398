"""Random irregular models (mixed element types, non-uniform port counts, mixed latencies, unconnected ports) as
WiredModel arrays -- exercised against the oracle on CPU (partition invariance) and by the CUDA engine on GPU."""
import numpy as np

from sst_core_b200.elements import NPARAM, TYPE_CLOCKNODE, TYPE_PHOLD
from sst_core_b200.wireup import WiredModel

NO_PEER = 0xFFFFFFFF


def random_model(seed, ncomp=300, stop_at=60_000, latencies=(1000, 1500, 2000, 5000), p_clock=0.3, far=False):
    rng = np.random.default_rng(seed)
    ctype = np.where(rng.random(ncomp) < p_clock, TYPE_CLOCKNODE, TYPE_PHOLD).astype(np.uint32)
    nports = np.where(ctype == TYPE_CLOCKNODE, 4, rng.integers(1, 7, ncomp)).astype(np.int64)
    port_off = np.zeros(ncomp + 1, dtype=np.int64)
    np.cumsum(nports, out=port_off[1:])
    nport = int(port_off[-1])
    peer = np.full(nport, NO_PEER, dtype=np.uint32)
    lat = np.zeros(nport, dtype=np.uint64)
    port_comp = np.repeat(np.arange(ncomp), nports)
    # phold needs port0.. contiguous-connected from 0 (it stops at the first unconnected port): connect phold ports
    # first, clocknode ports may stay open
    order = rng.permutation(nport)
    is_phold_port = ctype[port_comp] == TYPE_PHOLD
    must = [p for p in order if is_phold_port[p]]
    opt = [p for p in order if not is_phold_port[p]]
    pool = must + opt[: int(len(opt) * 0.8)]
    if len(pool) % 2:
        pool.append(opt[-1]) if len(opt) * 0.8 < len(opt) else pool.pop()
    if far:  # pair ports of distant components so that most links cross partitions and blocks
        pool = sorted(pool, key=lambda p: (port_comp[p] * 7919) % ncomp)
    for a, b in zip(pool[0::2], pool[1::2]):
        l = int(rng.choice(latencies))
        peer[a], peer[b] = b, a
        lat[a], lat[b] = l, int(rng.choice(latencies))
    # any phold port left unconnected (odd pool) -> make the component a loop on itself if it has a free port pair
    for c in range(ncomp):
        if ctype[c] != TYPE_PHOLD:
            continue
        ps = np.arange(port_off[c], port_off[c + 1])
        open_ports = [p for p in ps if peer[p] == NO_PEER]
        if open_ports:  # loop the port back onto itself (configuration loopback: same component, same port)
            for p in open_ports:
                peer[p] = p
                lat[p] = int(rng.choice(latencies))
    params = np.zeros((ncomp, NPARAM), dtype=np.int64)
    params[:, 0] = np.arange(ncomp)
    ph = ctype == TYPE_PHOLD
    params[ph, 1] = rng.integers(0, 9, ph.sum())         # init_events (some nodes inject nothing)
    params[ph, 2] = rng.choice([1, 700, 3000, 12000], ph.sum())  # delay_mod
    params[~ph, 1] = rng.choice([1, 3, 50], (~ph).sum())  # commFreq
    periods = np.array([500, 1000, 1300, 4000], dtype=np.uint64)
    clock = np.where(ph, 0, periods[rng.integers(0, 4, ncomp)]).astype(np.uint64)
    m = WiredModel(ctype, params, port_off.astype(np.uint32), peer, lat, clock, stop_at=stop_at, stats_enabled=True,
                   nocut_pairs=rng.integers(0, ncomp, (ncomp // 20, 2)).astype(np.uint32))
    m.validate()
    return m

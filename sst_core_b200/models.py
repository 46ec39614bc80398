"""Vectorised builders of WiredModel arrays for the large synthetic configurations.

A 4096x4096 torus has 16.7 M components and 67 M directed ports; building it object by object through the
``sst`` module surface (config.py) takes minutes in any interpreter (the reference needs ~90 us and ~20 KB per
component).  These builders produce, with numpy, exactly the arrays that
``wireup.wire(config.build_graph(script))`` produces for the corresponding scripts in tests/models/ (and for the
reference's tests/test_MessageMesh.py); tests/test_host.py checks that equality on small sizes.
"""
from __future__ import annotations

import numpy as np

from .elements import NPARAM, TYPE_CLOCKNODE, TYPE_MESH, TYPE_PHOLD
from .timelord import TimeLord
from .wireup import WiredModel


def _torus_ports(x: int, y: int, latency: int):
    """4 ports per component in the order x+, x-, y+, y- ; port 0 <-> neighbour's port 1, 2 <-> 3."""
    n = x * y
    i = np.arange(n, dtype=np.int64)
    mx, my = i % x, i // x
    xp = ((mx + 1) % x) + my * x
    xn = ((mx - 1) % x) + my * x
    yp = mx + ((my + 1) % y) * x
    yn = mx + ((my - 1) % y) * x
    peer = np.empty((n, 4), dtype=np.uint32)
    peer[:, 0] = 4 * xp + 1
    peer[:, 1] = 4 * xn + 0
    peer[:, 2] = 4 * yp + 3
    peer[:, 3] = 4 * yn + 2
    port_off = (4 * np.arange(n + 1, dtype=np.int64)).astype(np.uint32)
    lat = np.full(4 * n, latency, dtype=np.uint64)
    return port_off, peer.reshape(-1), lat


def mesh_model(x: int, y: int, stop_at="10us", mod: int = 1, verbose: int = 1, stats: int = 0,
               stats_enabled: bool = False, link_latency="1ns") -> WiredModel:
    """tests/test_MessageMesh.py (reference) with --model-options "X Y mod verbose stats stat_gen"."""
    tl = TimeLord()
    n = x * y
    port_off, peer, lat = _torus_ports(x, y, tl.cycles(link_latency))
    p = np.zeros((n, NPARAM), dtype=np.int64)
    p[:, 0] = np.arange(n)
    p[:, 1] = mod
    p[:, 2] = verbose
    p[:, 3] = stats
    p[:, 4] = 4
    p[:, 5] = 0x01010101
    i = np.arange(n, dtype=np.int64)
    mx, my = i % x, i // x
    even = (i % 2) == 0
    nocut = np.stack([i[even], (((mx + 1) % x) + my * x)[even]], axis=1).astype(np.uint32)
    m = WiredModel(np.full(n, TYPE_MESH, dtype=np.uint32), p, port_off, peer, lat,
                   np.zeros(n, dtype=np.uint64), stop_at=tl.cycles(stop_at), stats_enabled=stats_enabled,
                   nocut_pairs=nocut)
    return m


def phold_model(x: int, y: int, stop_at="1us", init_events: int = 16, delay_mod: int = 8000, verbose: int = 1,
                stats_enabled: bool = False, link_latency="1ns") -> WiredModel:
    """tests/models/phold.py."""
    tl = TimeLord()
    n = x * y
    port_off, peer, lat = _torus_ports(x, y, tl.cycles(link_latency))
    p = np.zeros((n, NPARAM), dtype=np.int64)
    p[:, 0] = np.arange(n)
    p[:, 1] = init_events
    p[:, 2] = delay_mod
    p[:, 3] = verbose
    return WiredModel(np.full(n, TYPE_PHOLD, dtype=np.uint32), p, port_off, peer, lat,
                      np.zeros(n, dtype=np.uint64), stop_at=tl.cycles(stop_at), stats_enabled=stats_enabled,
                      nocut_pairs=np.zeros((0, 2), dtype=np.uint32))


CLOCKS = ["1GHz", "500MHz", "250MHz", "2GHz"]


def clockpop_model(x: int, y: int, stop_at="1us", comm_freq: int = 1000, verbose: int = 1,
                   stats_enabled: bool = False, link_latency="1ns") -> WiredModel:
    """tests/models/clockpop.py."""
    tl = TimeLord()
    n = x * y
    port_off, peer, lat = _torus_ports(x, y, tl.cycles(link_latency))
    p = np.zeros((n, NPARAM), dtype=np.int64)
    p[:, 0] = np.arange(n)
    p[:, 1] = comm_freq
    p[:, 2] = verbose
    periods = np.array([tl.cycles(c) for c in CLOCKS], dtype=np.uint64)
    return WiredModel(np.full(n, TYPE_CLOCKNODE, dtype=np.uint32), p, port_off, peer, lat,
                      periods[np.arange(n) % 4], stop_at=tl.cycles(stop_at), stats_enabled=stats_enabled,
                      nocut_pairs=np.zeros((0, 2), dtype=np.uint32))

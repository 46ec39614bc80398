"""ctypes wrapper around oracle/liboracle.so (the CPU golden model) and runners for oracle/_ref (the reference
binary built from /root/reference).  TEST INFRASTRUCTURE: imported only from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / reference arm."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_SO = ROOT / "oracle" / "liboracle.so"
REF = ROOT / "oracle" / "_ref"
SSTSIM = REF / "libexec" / "sstsim.x"
REF_LIBDIR = REF / "lib" / "sst-elements-library"
NRESULT = 32


def build_oracle(force=False):
    src = ROOT / "oracle" / "pdes_oracle.cc"
    if force or not ORACLE_SO.exists() or ORACLE_SO.stat().st_mtime < src.stat().st_mtime:
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", str(ORACLE_SO), str(src)])
    return ORACLE_SO


MESH_GOLDEN_SO = ROOT / "oracle" / "libmeshgolden.so"
_mesh_lib = None


def mesh_golden_counts(x, y, windows, threads=0):
    """oracle/mesh_golden.cc: per-component received counts of an X x Y MessageMesh after `windows` lookahead windows
    (count-only golden model for sizes the event-by-event oracle cannot hold).  Returns (counts int32 [x*y], events)."""
    global _mesh_lib
    src = ROOT / "oracle" / "mesh_golden.cc"
    if _mesh_lib is None:
        if not MESH_GOLDEN_SO.exists() or MESH_GOLDEN_SO.stat().st_mtime < src.stat().st_mtime:
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-o", str(MESH_GOLDEN_SO),
                                   str(src)])
        _mesh_lib = C.CDLL(str(MESH_GOLDEN_SO))
        _mesh_lib.mesh_golden_counts.restype = C.c_uint64
        _mesh_lib.mesh_golden_counts.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint, C.c_void_p]
    counts = np.zeros(x * y, dtype=np.int32)
    ev = _mesh_lib.mesh_golden_counts(x, y, windows, threads or (os.cpu_count() or 1), _p(counts))
    if ev == 0xFFFFFFFFFFFFFFFF:
        raise RuntimeError("mesh_golden: a component drew more pairs than the route table holds (pairs_for in mesh_golden.cc)")
    return counts, int(ev)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build_oracle()
        L = C.CDLL(str(ORACLE_SO))
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [C.c_uint32] + [C.c_void_p] * 6 + [C.c_uint64, C.c_int, C.c_int]
        L.oracle_create2.restype = C.c_void_p
        L.oracle_create2.argtypes = [C.c_uint32] + [C.c_void_p] * 7 + [C.c_uint64, C.c_int, C.c_int]
        L.oracle_create3.restype = C.c_void_p
        L.oracle_create3.argtypes = [C.c_uint32] + [C.c_void_p] * 9 + [C.c_uint64, C.c_int, C.c_int]
        L.oracle_output_extra_count.restype = C.c_uint64
        L.oracle_output_extra_count.argtypes = [C.c_void_p]
        L.oracle_output_extra.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_output_count.restype = C.c_uint64
        L.oracle_output_count.argtypes = [C.c_void_p]
        L.oracle_output.argtypes = [C.c_void_p] * 6
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_run.argtypes = [C.c_void_p]
        L.oracle_results.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_counters.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_trace.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        L.oracle_rng_u32_stream.restype = C.c_uint64
        L.oracle_rng_u32_stream.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p]
        L.oracle_rng_ticks.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p]
        L.oracle_distribution.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_uint64,
                                          C.c_uint64, C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class OracleRun:
    def __init__(self, results, events, ticks, end_time, trace=None, output=None):
        self.results, self.events, self.ticks, self.end_time, self.trace = results, events, ticks, end_time, trace
        self.output = output  # console-output records in serial order: dict(time, comp, code, a, b)


def run_model(m, trace=False) -> OracleRun:
    """Serial golden-model run of a WiredModel."""
    L = lib()
    arrs = [np.ascontiguousarray(a) for a in
            (m.comp_type, m.comp_param, m.port_off, m.peer_port, m.latency, m.clock_period)]
    ps = getattr(m, "port_self", None)
    ps = np.ascontiguousarray(ps, dtype=np.uint8) if ps is not None else None
    xo = getattr(m, "xparam_off", None)
    xo = np.ascontiguousarray(xo, dtype=np.uint32) if xo is not None else None
    xp = np.ascontiguousarray(m.xparam, dtype=np.int64) if xo is not None else None
    h = L.oracle_create3(m.ncomp, *[_p(a) for a in arrs], _p(ps) if ps is not None else None,
                         _p(xo) if xo is not None else None, _p(xp) if xo is not None else None, int(m.stop_at),
                         int(bool(m.stats_enabled)), int(trace))
    if not h:
        raise RuntimeError("oracle_create failed (unknown component type)")
    try:
        L.oracle_run(h)
        res = np.zeros((m.ncomp, NRESULT), dtype=np.uint64)
        L.oracle_results(h, _p(res))
        cnt = np.zeros(4, dtype=np.uint64)
        L.oracle_counters(h, _p(cnt))
        tr = None
        if trace:
            n = int(cnt[3])
            tr = dict(time=np.zeros(n, np.uint64), comp=np.zeros(n, np.uint32), port=np.zeros(n, np.uint32),
                      payload=np.zeros(n, np.uint64))
            L.oracle_trace(h, _p(tr["time"]), _p(tr["comp"]), _p(tr["port"]), _p(tr["payload"]))
        no = int(L.oracle_output_count(h))
        out = dict(time=np.zeros(no, np.uint64), comp=np.zeros(no, np.uint32), code=np.zeros(no, np.uint32),
                   a=np.zeros(no, np.uint64), b=np.zeros(no, np.uint64))
        if no:
            L.oracle_output(h, *[_p(out[k]) for k in ("time", "comp", "code", "a", "b")])
        # statistic outputs carry five values: the last three come separately, in record order
        ne = int(L.oracle_output_extra_count(h))
        extra = np.zeros((ne, 3), dtype=np.uint64)
        if ne:
            L.oracle_output_extra(h, _p(extra))
        out["cde"] = np.zeros((no, 3), dtype=np.uint64)
        if ne:
            is_stat = np.isin(m.comp_type[out["comp"]], (7, 8))
            assert int(is_stat.sum()) == ne
            out["cde"][is_stat] = extra
        return OracleRun(res, int(cnt[0]), int(cnt[1]), int(cnt[2]), tr, out)
    finally:
        L.oracle_destroy(h)


def rng_u32_stream(kind, s1, s2, n, skip=0, want=True):
    out = np.zeros(n, dtype=np.uint32) if want else None
    cs = lib().oracle_rng_u32_stream(kind, s1, s2, skip, n, _p(out) if want else None)
    return out, cs


def rng_ticks(kind, s1, s2, first, n):
    out = np.zeros((n, 5), dtype=np.uint64)
    lib().oracle_rng_ticks(kind, s1, s2, first, n, _p(out))
    return out


def distribution(dist, params, rng_kind, s1, s2, n):
    p = np.asarray(params, dtype=np.float64)
    out = np.zeros(n, dtype=np.float64)
    lib().oracle_distribution(dist, _p(p), len(p), rng_kind, s1, s2, n, _p(out))
    return out


# ------------------------------------------------------------------------------------------------ reference binary
def have_ref() -> bool:
    return SSTSIM.exists() and (REF_LIBDIR / "libcoreTestElement.so").exists()


def run_ref(script, model_options="", extra=(), cwd=None, threads=1, timeout=3600) -> str:
    """Run the hand-built reference `sstsim.x` on a model script; returns stdout."""
    cmd = [str(SSTSIM), "--lib-path", str(REF_LIBDIR)]
    if threads > 1:
        cmd += ["-n", str(threads)]
    if model_options:
        cmd += ["--model-options", model_options]
    cmd += list(extra) + [str(script)]
    env = dict(os.environ)
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, cwd=cwd, env=env,
                       timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError(f"reference run failed ({r.returncode}): {' '.join(cmd)}\n{r.stdout[-2000:]}")
    return r.stdout

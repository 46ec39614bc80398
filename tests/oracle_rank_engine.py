"""CPU stand-in with the stepping interface of sst_core_b200.engine.Engine, backed by the oracle's partitioned
stepping (oracle_run_until / oracle_take_outbox / oracle_ingest).  Lets the N>1 host path -- partition_model, the
exchange wire format, DistributedRunner's step order and the termination protocol -- run over gloo on CPU.
Its advance() restates advance_kernel (sst_core_b200/csrc/engine.cu) in Python."""
import ctypes as C

import numpy as np
import torch

import oracle_lib as O
from sst_core_b200.engine import Status

CW = 8
XHDR = 10  # exchange slot header: count, 8 control words, reserved (include/sstb200.h)


class OracleRankEngine:
    def __init__(self, pm, rank, n_ranks, outbox_capacity=4096):
        L = O.lib()
        L.oracle_set_ownership.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        L.oracle_setup_owned.argtypes = [C.c_void_p]
        L.oracle_run_until.argtypes = [C.c_void_p, C.c_uint64]
        L.oracle_run_until.restype = C.c_uint64
        L.oracle_take_outbox.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_ingest.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        L.oracle_control_words.argtypes = [C.c_void_p, C.c_void_p]
        self.L, self.pm, self.rank, self.n_ranks = L, pm, rank, n_ranks
        counts = np.bincount(pm.rank, minlength=n_ranks)
        self.begins = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        self.c0, self.c1 = int(self.begins[rank]), int(self.begins[rank + 1])
        arrs = [np.ascontiguousarray(a) for a in
                (pm.comp_type, pm.comp_param, pm.port_off, pm.peer_port, pm.latency, pm.clock_period)]
        self.h = L.oracle_create(pm.ncomp, *[a.ctypes.data_as(C.c_void_p) for a in arrs], 0,
                                 int(bool(pm.stats_enabled)), 0)
        L.oracle_set_ownership(self.h, self.c0, self.c1)
        conn = pm.peer_port != 0xFFFFFFFF
        self.w = int(pm.latency[conn].min()) if conn.any() else 1_000_000
        self.cap = outbox_capacity
        self.words = XHDR + 4 * outbox_capacity
        self.outbox = torch.zeros((n_ranks, self.words), dtype=torch.int64)
        self.inbox = torch.zeros((n_ranks, self.words), dtype=torch.int64)
        self.local = torch.zeros(CW, dtype=torch.int64)
        self.gathered = torch.zeros(CW * n_ranks, dtype=torch.int64)
        self.k, self.done, self.end_time, self.windows = 0, False, 0, 0
        self.stop_at = int(pm.stop_at)
        self.model = pm
        self._report_at, self._snapshot = None, None

    # ---- Engine interface
    def setup(self):
        self.L.oracle_setup_owned(self.h)

    def exchange_tensors(self):
        return self.outbox, self.inbox

    def control_tensors(self):
        return self.local, self.gathered

    def _t_end(self):
        t = (self.k + 1) * self.w
        return min(t, self.stop_at) if self.stop_at else t

    def window(self):
        return self.w

    # periodic statistic output: bracket the window that contains the report time (Engine.report_begin/report_end)
    def report_begin(self, report_time):
        self._report_at = int(report_time)

    def report_end(self, words=0):
        self._report_at = None
        return self._snapshot

    def dispatch(self):
        if self.done:
            return
        n = 0
        if self._report_at is not None and self.k * self.w <= self._report_at < self._t_end():
            n = int(self.L.oracle_run_until(self.h, self._report_at + 1))  # everything with time <= the report time
            self._snapshot = self.results()
        n = int(self.L.oracle_run_until(self.h, self._t_end()))
        rec = np.zeros((n, 4), dtype=np.uint64)
        if n:
            self.L.oracle_take_outbox(self.h, rec.ctypes.data_as(C.c_void_p))
        self.outbox.zero_()
        ob = self.outbox.numpy().view(np.uint64)
        dst_rank = np.searchsorted(self.begins, rec[:, 3].astype(np.int64), side="right") - 1
        for r in range(self.n_ranks):
            sel = rec[dst_rank == r]
            assert sel.shape[0] <= self.cap, "outbox overflow"
            ob[r, 0] = sel.shape[0]
            ob[r, XHDR:XHDR + 4 * sel.shape[0]] = sel.reshape(-1)
        # control words ride in every outbox header (pending counts what is still queued here plus what was just sent)
        cw = np.zeros(CW, dtype=np.int64)
        self.L.oracle_control_words(self.h, cw.ctypes.data_as(C.c_void_p))
        cw[0] += n
        self.outbox[:, 1:1 + CW] = torch.from_numpy(cw)
        self.local.copy_(torch.from_numpy(cw))

    def ingest(self):
        if self.done:
            return
        ib = self.inbox.numpy().view(np.uint64)
        for r in range(self.n_ranks):
            if r == self.rank:
                continue
            n = int(ib[r, 0])
            rec = np.ascontiguousarray(ib[r, XHDR:XHDR + 4 * n])
            if n:
                self.L.oracle_ingest(self.h, rec.ctypes.data_as(C.c_void_p), n)
        g = self.gathered.view(self.n_ranks, CW)
        for r in range(self.n_ranks):
            g[r] = self.local if r == self.rank else self.inbox[r, 1:1 + CW]

    def advance(self):
        if self.done:
            return
        g = self.gathered.numpy().reshape(self.n_ranks, CW)
        pending, primary, clocks = int(g[:, 0].sum()), int(g[:, 1].sum()), int(g[:, 2].sum())
        exit_time = int(g[:, 4].max())
        self.windows += 1
        if primary <= 0:
            self.done, self.end_time = True, exit_time
            return
        next_t = (self.k + 1) * self.w
        if self.stop_at and next_t >= self.stop_at:
            self.done, self.end_time = True, self.stop_at
            return
        if pending <= 0 and clocks <= 0:
            self.done, self.end_time = True, self._t_end()
            return
        self.k += 1

    def status(self):
        cnt = np.zeros(4, dtype=np.uint64)
        self.L.oracle_counters(self.h, cnt.ctypes.data_as(C.c_void_p))
        return Status(self.k * self.w, self.end_time, self.windows, int(cnt[0]), int(cnt[1]), 0, 0, 0, self.done, 0,
                      (0, 0, 0, 0), 0)

    def raise_on_error(self, st=None):
        pass

    def results(self):
        res = np.zeros((self.pm.ncomp, O.NRESULT), dtype=np.uint64)
        self.L.oracle_results(self.h, res.ctypes.data_as(C.c_void_p))
        return res[self.c0:self.c1]

    def close(self):
        if self.h:
            self.L.oracle_destroy(self.h)
            self.h = None

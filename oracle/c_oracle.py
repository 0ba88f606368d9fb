"""ctypes wrapper of oracle/liboracle.so (window_oracle.c): the C restatement of the reference's
tumbling / sliding aggregate.  TEST INFRASTRUCTURE ONLY (checker + timed CPU baseline)."""
import ctypes as C
import os
import subprocess

import numpy as np

from . import arroyo_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")
_lib = None
I64P = C.POINTER(C.c_int64)


class _Out(C.Structure):
    _fields_ = [(n, I64P) for n in ("key", "wstart", "wend", "rows", "sum", "mn", "mx", "ts")] + [
        ("avg", C.POINTER(C.c_double)), ("n", C.c_int64), ("cap", C.c_int64)]


class RunResult(C.Structure):
    _fields_ = [("seconds", C.c_double), ("rows_in", C.c_uint64), ("rows_out", C.c_uint64),
                ("windows_out", C.c_uint64), ("late_rows", C.c_uint64), ("sum_of_sums", C.c_uint64),
                ("sum_of_rows", C.c_uint64), ("sum_of_avgs", C.c_double), ("threads", C.c_int)]


class WindowSum(C.Structure):
    """Checksums of one emitted window (OracleWindowSum)."""
    _fields_ = [("wstart", C.c_int64), ("wend", C.c_int64), ("rows_out", C.c_uint64), ("sum_of_rows", C.c_uint64),
                ("sum_of_sums", C.c_uint64), ("sum_of_avgs", C.c_double)]


def load():
    global _lib
    if _lib is not None:
        return _lib
    src = os.path.join(HERE, "window_oracle.c")
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE])
    lib = C.CDLL(LIB)
    lib.oracle_window_create.restype = C.c_void_p
    lib.oracle_window_create.argtypes = [C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int]
    lib.oracle_window_destroy.argtypes = [C.c_void_p]
    lib.oracle_window_process_batch.argtypes = [C.c_void_p, I64P, I64P, I64P, C.c_int64, C.c_int, C.c_int64]
    lib.oracle_window_handle_watermark.argtypes = [C.c_void_p, C.c_int64, C.POINTER(_Out)]
    lib.oracle_window_handle_checkpoint.argtypes = [C.c_void_p, C.c_int, C.c_int64]
    lib.oracle_window_late_rows.restype = C.c_uint64
    lib.oracle_window_late_rows.argtypes = [C.c_void_p]
    lib.oracle_out_create.restype = C.POINTER(_Out)
    lib.oracle_out_clear.argtypes = [C.POINTER(_Out)]
    lib.oracle_out_destroy.argtypes = [C.POINTER(_Out)]
    lib.oracle_max_threads.restype = C.c_int
    lib.oracle_run_windows.restype = C.c_int
    lib.oracle_run_windows.argtypes = [I64P, I64P, I64P, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                       C.c_int, C.c_int, C.POINTER(RunResult)]
    lib.oracle_runner_create.restype = C.c_void_p
    lib.oracle_runner_create.argtypes = [C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int64]
    lib.oracle_runner_feed.restype = C.c_double
    lib.oracle_runner_feed.argtypes = [C.c_void_p, I64P, I64P, I64P, C.c_int64]
    lib.oracle_runner_finish.restype = C.c_double
    lib.oracle_runner_finish.argtypes = [C.c_void_p]
    lib.oracle_runner_result.argtypes = [C.c_void_p, C.POINTER(RunResult)]
    lib.oracle_runner_destroy.argtypes = [C.c_void_p]
    lib.oracle_runner_windows.restype = C.c_int64
    lib.oracle_runner_windows.argtypes = [C.c_void_p, C.POINTER(WindowSum), C.c_int64]
    _lib = lib
    return lib


def _p(a):
    return a.ctypes.data_as(I64P)


class _WindowOp:
    """Same call surface as the numpy oracle's window operators (one value column at most)."""
    sliding = False

    def __init__(self, cfg: O.WindowAggConfig):
        self.cfg = cfg
        self.lib = load()
        cols = {a.col for a in cfg.aggs if a.col is not None}
        assert len(cols) <= 1 and len(cfg.key_names) <= 1
        self.val_col = next(iter(cols)) if cols else None
        self.key_col = cfg.key_names[0] if cfg.key_names else None
        want_minmax = any(a.kind in ("min", "max") for a in cfg.aggs)
        self.h = self.lib.oracle_window_create(cfg.width, cfg.slide if self.sliding else 0,
                                               1 if self.key_col else 0, 1 if want_minmax else 0,
                                               1 if cfg.final_projection else 0)
        self.out = self.lib.oracle_out_create()

    def __del__(self):
        try:
            self.lib.oracle_window_destroy(self.h)
            self.lib.oracle_out_destroy(self.out)
        except Exception:
            pass

    def process_batch(self, batch: O.Batch, ctx, collector):
        n = batch.num_rows
        if n == 0:
            return
        k = np.ascontiguousarray(batch[self.key_col], dtype=np.int64) if self.key_col else None
        v = np.ascontiguousarray(batch[self.val_col], dtype=np.int64) if self.val_col else None
        t = np.ascontiguousarray(batch[O.TIMESTAMP], dtype=np.int64)
        wm = ctx.last_present_watermark()
        self.lib.oracle_window_process_batch(self.h, _p(k) if k is not None else None, _p(v) if v is not None else None,
                                             _p(t), n, 0 if wm is None else 1,
                                             0 if wm is None else min(wm, (1 << 63) - 1))

    def handle_watermark(self, watermark, ctx, collector):
        wm = ctx.last_present_watermark()
        if wm is None:
            return watermark
        self.lib.oracle_window_handle_watermark(self.h, min(wm, (1 << 63) - 1), self.out)
        o = self.out.contents
        n = o.n
        if n:
            def arr(p, dt=np.int64):
                return np.ctypeslib.as_array(p, shape=(n,)).astype(dt, copy=True)
            ws = arr(o.wstart)
            cuts = np.flatnonzero(np.diff(ws)) + 1
            starts = np.concatenate([[0], cuts])
            ends = np.concatenate([cuts, [n]])
            full = {"key": arr(o.key), "window_start": ws, "window_end": arr(o.wend), "rows": arr(o.rows),
                    "sum": arr(o.sum), "avg": np.ctypeslib.as_array(o.avg, shape=(n,)).copy(), "mn": arr(o.mn),
                    "mx": arr(o.mx), O.TIMESTAMP: arr(o.ts)}
            for s, e in zip(starts, ends):
                items = []
                if self.key_col:
                    items.append((self.key_col, full["key"][s:e]))
                for a in self.cfg.aggs:
                    src = {"count": "rows", "sum": "sum", "avg": "avg", "min": "mn", "max": "mx"}[a.kind]
                    items.append((a.name, full[src][s:e]))
                if self.cfg.final_projection:
                    items[self.cfg.window_index:self.cfg.window_index] = [
                        ("window_start", full["window_start"][s:e]), ("window_end", full["window_end"][s:e])]
                cols = dict(items)
                cols[O.TIMESTAMP] = full[O.TIMESTAMP][s:e]
                collector.collect(O.Batch(cols))
            self.lib.oracle_out_clear(self.out)
        return watermark

    def handle_checkpoint(self, ctx):
        w = ctx.watermarks.cur_watermark
        has = w is not None and w != O.IDLE
        self.lib.oracle_window_handle_checkpoint(self.h, 1 if has else 0, min(w, (1 << 63) - 1) if has else 0)


class TumblingAggregatingWindowFunc(_WindowOp):
    sliding = False


class SlidingAggregatingWindowFunc(_WindowOp):
    sliding = True


run_single_input = O.run_single_input


# ---------------------------------------------------------------------------------------------------------------
# session windows (session_oracle.c)
# ---------------------------------------------------------------------------------------------------------------
SESSION_LIB = os.path.join(HERE, "liboracle_session.so")
_session_lib = None


class _SessionOut(C.Structure):
    _fields_ = [(n, I64P) for n in ("key", "start", "end", "rows", "sum", "mn", "mx", "ts")] + [
        ("avg", C.POINTER(C.c_double)), ("n", C.c_int64), ("cap", C.c_int64)]


def load_session():
    global _session_lib
    if _session_lib is not None:
        return _session_lib
    src = os.path.join(HERE, "session_oracle.c")
    if not os.path.exists(SESSION_LIB) or os.path.getmtime(SESSION_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE, "liboracle_session.so"])
    lib = C.CDLL(SESSION_LIB)
    lib.oracle_session_create.restype = C.c_void_p
    lib.oracle_session_create.argtypes = [C.c_int64, C.c_int]
    lib.oracle_session_destroy.argtypes = [C.c_void_p]
    lib.oracle_session_error.restype = C.c_int
    lib.oracle_session_error.argtypes = [C.c_void_p]
    lib.oracle_session_late_rows.restype = C.c_uint64
    lib.oracle_session_late_rows.argtypes = [C.c_void_p]
    lib.oracle_session_process_batch.argtypes = [C.c_void_p, I64P, I64P, I64P, C.c_int64, C.c_int, C.c_int64]
    lib.oracle_session_out_create.restype = C.POINTER(_SessionOut)
    lib.oracle_session_out_clear.argtypes = [C.POINTER(_SessionOut)]
    lib.oracle_session_out_destroy.argtypes = [C.POINTER(_SessionOut)]
    lib.oracle_session_handle_watermark.argtypes = [C.c_void_p, C.c_int64, C.POINTER(_SessionOut)]
    _session_lib = lib
    return lib


class SessionAggregatingWindowFunc:
    """The oracle's session operator interface over session_oracle.c (at most one Int64 key column and one Int64
    value column; COUNT(*) / SUM / AVG / MIN / MAX)."""

    def __init__(self, cfg: O.SessionConfig):
        self.cfg = cfg
        self.lib = load_session()
        cols = {a.col for a in cfg.aggs if a.col is not None}
        assert len(cols) <= 1 and len(cfg.key_names) <= 1
        self.val_col = next(iter(cols)) if cols else None
        self.key_col = cfg.key_names[0] if cfg.key_names else None
        self.h = self.lib.oracle_session_create(cfg.gap, 1 if self.key_col else 0)
        self.out = self.lib.oracle_session_out_create()

    def name(self):
        return "session_window"

    def __del__(self):
        try:
            self.lib.oracle_session_destroy(self.h)
            self.lib.oracle_session_out_destroy(self.out)
        except Exception:
            pass

    def _check(self):
        e = self.lib.oracle_session_error(self.h)
        if e & 1:
            raise RuntimeError("should not have flushed batches when adding a batch")
        if e & 2:
            raise RuntimeError("received a batch that starts before the current data_start - gap")

    def process_batch(self, batch: O.Batch, ctx, collector):
        n = batch.num_rows
        if n == 0:
            return
        k = np.ascontiguousarray(batch[self.key_col], dtype=np.int64) if self.key_col else None
        v = np.ascontiguousarray(batch[self.val_col], dtype=np.int64) if self.val_col else None
        t = np.ascontiguousarray(batch[O.TIMESTAMP], dtype=np.int64)
        wm = ctx.last_present_watermark()
        self.lib.oracle_session_process_batch(self.h, _p(k) if k is not None else None, _p(v) if v is not None else None,
                                              _p(t), n, 0 if wm is None else 1,
                                              0 if wm is None else min(wm, (1 << 63) - 1))
        self._check()

    def handle_watermark(self, watermark, ctx, collector):
        wm = ctx.last_present_watermark()
        if wm is None:
            return watermark
        self.lib.oracle_session_out_clear(self.out)
        self.lib.oracle_session_handle_watermark(self.h, min(wm, (1 << 63) - 1), self.out)
        self._check()
        o = self.out.contents
        n = int(o.n)
        if n:
            def arr(p):
                return np.ctypeslib.as_array(p, shape=(n,)).copy()
            full = {"rows": arr(o.rows), "sum": arr(o.sum), "avg": np.ctypeslib.as_array(o.avg, shape=(n,)).copy(),
                    "mn": arr(o.mn), "mx": arr(o.mx)}
            items = [(self.key_col, arr(o.key))] if self.key_col else []
            items[self.cfg.window_index:self.cfg.window_index] = [("window_start", arr(o.start)), ("window_end", arr(o.end))]
            for a in self.cfg.aggs:
                items.append((a.name, full[{"count": "rows", "sum": "sum", "avg": "avg", "min": "mn", "max": "mx"}[a.kind]]))
            cols = dict(items)
            cols[O.TIMESTAMP] = arr(o.ts)
            collector.collect(O.Batch(cols))
        return watermark


# ---------------------------------------------------------------------------------------------------------------
# instant join (join_oracle.c)
# ---------------------------------------------------------------------------------------------------------------
JOIN_LIB = os.path.join(HERE, "liboracle_join.so")
_join_lib = None
_JOIN_TYPES = {"inner": 0, "left": 1, "right": 2, "full": 3}


class _JoinOut(C.Structure):
    _fields_ = [("n", C.c_int64), ("cap", C.c_int64), ("n_cols", C.c_int), ("cols", C.POINTER(I64P)),
                ("valid", C.POINTER(C.POINTER(C.c_uint8)))]


def load_join():
    global _join_lib
    if _join_lib is not None:
        return _join_lib
    src = os.path.join(HERE, "join_oracle.c")
    if not os.path.exists(JOIN_LIB) or os.path.getmtime(JOIN_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE, "liboracle_join.so"])
    lib = C.CDLL(JOIN_LIB)
    lib.oracle_join_create.restype = C.c_void_p
    lib.oracle_join_create.argtypes = [C.c_int] * 7
    lib.oracle_join_destroy.argtypes = [C.c_void_p]
    lib.oracle_join_process.restype = C.c_int
    lib.oracle_join_process.argtypes = [C.c_void_p, C.c_int, C.POINTER(I64P), C.c_int64, C.c_int, C.c_int64]
    lib.oracle_join_out_create.restype = C.POINTER(_JoinOut)
    lib.oracle_join_out_create.argtypes = [C.c_void_p]
    lib.oracle_join_out_clear.argtypes = [C.POINTER(_JoinOut)]
    lib.oracle_join_out_destroy.argtypes = [C.POINTER(_JoinOut)]
    lib.oracle_join_handle_watermark.argtypes = [C.c_void_p, C.c_int64, C.POINTER(_JoinOut)]
    lib.oracle_join_buffered.restype = C.c_int64
    lib.oracle_join_buffered.argtypes = [C.c_void_p, C.c_int]
    _join_lib = lib
    return lib


class InstantJoin:
    """The oracle's InstantJoin interface over join_oracle.c (single-column Int64 equi-join, no routing-key copies).
    Both inputs' layouts must be known before the first watermark (true of every pipeline in tests/golden_cases.py)."""

    def __init__(self, cfg: O.JoinConfig):
        assert len(cfg.left_on) == 1 and len(cfg.right_on) == 1 and not cfg.left_routing_keys and not cfg.right_routing_keys
        self.cfg = cfg
        self.lib = load_join()
        self.names = [None, None]
        self.pending = []  # batches that arrived before both layouts were known
        self.h = None
        self.out = None

    def name(self):
        return "InstantJoin"

    def _create(self):
        ln, rn = self.names
        self.h = self.lib.oracle_join_create(_JOIN_TYPES[self.cfg.join_type], len(ln), ln.index(self.cfg.left_on[0]),
                                             ln.index(O.TIMESTAMP), len(rn), rn.index(self.cfg.right_on[0]),
                                             rn.index(O.TIMESTAMP))
        self.out = self.lib.oracle_join_out_create(self.h)
        out_names = [c for c in ln if c != O.TIMESTAMP]
        for c in rn:
            if c != O.TIMESTAMP:
                out_names.append(c if c not in out_names else c + "_right")
        self.out_names = out_names + [O.TIMESTAMP]
        for side, batch, wm in self.pending:
            self._send(side, batch, wm)
        self.pending = []

    def _send(self, side, batch, wm):
        cols = [np.ascontiguousarray(batch[c], dtype=np.int64) for c in self.names[side]]
        ptrs = (I64P * len(cols))(*[_p(c) for c in cols])
        rc = self.lib.oracle_join_process(self.h, side, ptrs, batch.num_rows, 0 if wm is None else 1,
                                          0 if wm is None else int(min(wm, (1 << 63) - 1)))
        if rc != 0:
            raise RuntimeError("shouldn't have a batch with timestamp before the watermark")  # instant_join.rs:129-139

    def process_batch_index(self, index, total_inputs, batch, ctx, collector):
        if batch.num_rows == 0:
            raise RuntimeError("should have max timestamp")  # instant_join.rs:123
        side = index // (total_inputs // 2)
        if self.names[side] is None:
            self.names[side] = batch.names()
        wm = ctx.last_present_watermark()
        if self.h is None:
            if self.names[0] is not None and self.names[1] is not None:
                self._create()
            else:
                self.pending.append((side, batch, wm))
                return
        self._send(side, batch, wm)

    def handle_watermark(self, watermark, ctx, collector):
        wm = ctx.last_present_watermark()
        if wm is None:
            return watermark
        if self.h is None:
            raise NotImplementedError("a watermark before both join inputs produced a batch")
        self.lib.oracle_join_out_clear(self.out)
        self.lib.oracle_join_handle_watermark(self.h, int(min(wm, (1 << 63) - 1)), self.out)
        o = self.out.contents
        n = int(o.n)
        if n:
            cols, valid = {}, {}
            for c, name in enumerate(self.out_names):
                cols[name] = np.ctypeslib.as_array(o.cols[c], shape=(n,)).copy()
                v = np.ctypeslib.as_array(o.valid[c], shape=(n,)).astype(bool)
                if not v.all():
                    valid[name] = v
            collector.collect(O.Batch(cols, valid))
        return wm

    def __del__(self):
        try:
            if self.out is not None:
                self.lib.oracle_join_out_destroy(self.out)
            if self.h is not None:
                self.lib.oracle_join_destroy(self.h)
        except Exception:
            pass


def run_windows(key, val, ts, batch_rows, width, slide, wm_delay, threads, flush_at_end=True) -> RunResult:
    lib = load()
    res = RunResult()
    key = np.ascontiguousarray(key, dtype=np.int64)
    val = np.ascontiguousarray(val, dtype=np.int64)
    ts = np.ascontiguousarray(ts, dtype=np.int64)
    rc = lib.oracle_run_windows(_p(key), _p(val), _p(ts), len(key), batch_rows, width, slide, wm_delay, threads,
                                1 if flush_at_end else 0, C.byref(res))
    assert rc == 0
    return res


class Runner:
    """Resumable parallel run (p key-partitioned single-threaded window subtasks)."""

    def __init__(self, threads, width, slide, wm_delay, batch_rows):
        self.lib = load()
        self.h = self.lib.oracle_runner_create(threads, width, slide, wm_delay, batch_rows)
        assert self.h

    def feed(self, key, val, ts) -> float:
        key = np.ascontiguousarray(key, dtype=np.int64)
        val = np.ascontiguousarray(val, dtype=np.int64)
        ts = np.ascontiguousarray(ts, dtype=np.int64)
        return self.lib.oracle_runner_feed(self.h, _p(key), _p(val), _p(ts), len(key))

    def finish(self) -> float:
        return self.lib.oracle_runner_finish(self.h)

    def result(self) -> RunResult:
        r = RunResult()
        self.lib.oracle_runner_result(self.h, C.byref(r))
        return r

    def windows(self):
        """[{wstart, wend, rows_out, sum_of_rows, sum_of_sums (wrapping u64), sum_of_avgs}] per emitted window."""
        n = self.lib.oracle_runner_windows(self.h, None, 0)
        buf = (WindowSum * max(n, 1))()
        self.lib.oracle_runner_windows(self.h, buf, n)
        return [{f: getattr(buf[i], f) for f, _ in WindowSum._fields_} for i in range(n)]

    def close(self):
        if self.h:
            self.lib.oracle_runner_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

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


SessionAggregatingWindowFunc = O.SessionAggregatingWindowFunc
InstantJoin = O.InstantJoin
run_single_input = O.run_single_input


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

    def close(self):
        if self.h:
            self.lib.oracle_runner_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

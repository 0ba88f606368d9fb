"""Adapter that lets the shared golden / parity cases (written against the oracle's operator
interface) drive the CUDA operators through arroyo_b200's host mirror and the C ABI.

Conversions only: oracle Batch (dict of numpy) <-> pyarrow RecordBatch; the `window` struct column is
flattened to window_start / window_end the way the oracle names them."""
import numpy as np
import pyarrow as pa

import arroyo_b200 as ab
from arroyo_b200 import operators as native
from oracle import arroyo_oracle as O

TIMESTAMP = O.TIMESTAMP
DEFAULT_KW = {}


def to_arrow(batch: O.Batch) -> pa.RecordBatch:
    arrays, names = [], []
    for k, v in batch.cols.items():
        if k == TIMESTAMP:
            arrays.append(pa.array(np.ascontiguousarray(v, dtype=np.int64), type=pa.int64()).cast(pa.timestamp("ns")))
        elif v.dtype == np.uint64:
            arrays.append(pa.array(np.ascontiguousarray(v), type=pa.uint64()))
        elif v.dtype.kind == "f":
            arrays.append(pa.array(np.ascontiguousarray(v, dtype=np.float64), type=pa.float64()))
        else:
            arrays.append(pa.array(np.ascontiguousarray(v, dtype=np.int64), type=pa.int64()))
        names.append(k)
    return pa.RecordBatch.from_arrays(arrays, names=names)


def _np(arr: pa.Array) -> np.ndarray:
    if pa.types.is_timestamp(arr.type):
        arr = arr.cast(pa.int64())
    return arr.to_numpy(zero_copy_only=False)


def from_arrow(rb: pa.RecordBatch) -> O.Batch:
    cols, valid = {}, {}
    for name, col in zip(rb.schema.names, rb.columns):
        if pa.types.is_struct(col.type):
            cols["window_start"] = _np(col.field(0))
            cols["window_end"] = _np(col.field(1))
        elif col.null_count:
            valid[name] = np.asarray(col.is_valid())
            cols[name] = _np(col.fill_null(0))
        else:
            cols[name] = _np(col)
    return O.Batch(cols, valid)


class _CollectAdapter(ab.Collector):
    def __init__(self, sink):
        super().__init__()
        self.sink = sink

    def collect(self, batch):
        self.sink.collect(from_arrow(batch))


class _WindowOp:
    native_cls = None

    def __init__(self, cfg, **kw):
        self.cfg = cfg
        opts = dict(DEFAULT_KW)
        opts.update(kw)
        self.op = self.native_cls(cfg, **opts)

    def process_batch(self, batch: O.Batch, ctx, collector):
        self.op.process_batch(to_arrow(batch), ctx, _CollectAdapter(collector))

    def handle_watermark(self, watermark, ctx, collector):
        return self.op.handle_watermark(watermark, ctx, _CollectAdapter(collector))

    def stats(self):
        return self.op.stats()

    def close(self):
        self.op.close()


class TumblingAggregatingWindowFunc(_WindowOp):
    native_cls = native.TumblingAggregatingWindowFunc


class SlidingAggregatingWindowFunc(_WindowOp):
    native_cls = native.SlidingAggregatingWindowFunc


class InstantJoin:
    def __init__(self, cfg, **kw):
        self.op = native.InstantJoin(cfg, **kw)

    def process_batch_index(self, index, total_inputs, batch: O.Batch, ctx, collector):
        self.op.process_batch_index(index, total_inputs, to_arrow(batch), ctx, _CollectAdapter(collector))

    def handle_watermark(self, watermark, ctx, collector):
        return self.op.handle_watermark(watermark, ctx, _CollectAdapter(collector))


class SessionAggregatingWindowFunc(_WindowOp):
    native_cls = native.SessionAggregatingWindowFunc


def run_single_input(op, batches, delay_ns: int = 1_000_000_000, ctx=None) -> O.Collector:
    """Same driver as oracle.run_single_input (operator.rs:932-1066 run loop for one input)."""
    ctx = ctx or O.OperatorContext(1)
    out = O.Collector()
    gen = O.WatermarkGenerator(delay_ns)
    for b in batches:
        op.process_batch(b, ctx, out)
        wm = gen.process_batch(b[TIMESTAMP])
        if wm is not None:
            ctx.watermarks.set(0, wm)
            op.handle_watermark(wm, ctx, out)
    ctx.watermarks.set(0, O.FINAL_WATERMARK)
    op.handle_watermark(O.FINAL_WATERMARK, ctx, out)
    return out

"""The reference's golden smoke-test queries (crates/arroyo-sql-testing/src/test/queries/*.sql)
expressed as operator pipelines.  Shared by the oracle tests (CPU) and the GPU parity tests:
`ops` is a factory namespace providing TumblingAggregatingWindowFunc, SlidingAggregatingWindowFunc,
SessionAggregatingWindowFunc and InstantJoin with the oracle's constructor signatures, plus the
`run_single_input` driver, so the same case runs against the oracle and the CUDA operators.

The smoke tests force `pipeline.source_batch_size = 32` (smoke_tests.rs:51-54); the default
watermark is `_timestamp - 1 s` (arroyo-planner/src/rewriters.rs:71-82).
"""
from collections import Counter

import numpy as np

from oracle import arroyo_oracle as O

S = 1_000_000_000
MIN = 60 * S
HOUR = 3600 * S
DAY = 24 * HOUR
BATCH = 32


def multiset(rows):
    return Counter(tuple(sorted(r.items())) for r in rows)


def _rows(batch, mapping):
    """Project an output batch into golden-row dicts: mapping = {golden_name: column}."""
    if batch is None:
        return []
    out = []
    for r in batch.rows():
        out.append({g: r[c] for g, c in mapping.items()})
    return out


def sliding_window_end(ops, inputs):
    """hop(2 s, 10 s) count(*), min(counter), max(counter) GROUP BY window (unkeyed)."""
    cfg = O.WindowAggConfig(width=10 * S, slide=2 * S, key_names=[],
                            aggs=[O.Agg("count", None, "count"), O.Agg("min", "counter", "min"),
                                  O.Agg("max", "counter", "max")], window_index=0)
    op = ops.SlidingAggregatingWindowFunc(cfg)
    b = O.source_batches({"counter": inputs["impulse_counter"], O.TIMESTAMP: inputs["impulse_ts"]}, BATCH)
    out = ops.run_single_input(op, b).all()
    return _rows(out, {"count": "count", "min": "min", "max": "max", "start": "window_start", "end": "window_end"})


def hourly_by_event_type(ops, inputs):
    """TUMBLE(1 h) COUNT(*) GROUP BY event_type (string key, dictionary-encoded by make_golden.py)."""
    cfg = O.WindowAggConfig(width=HOUR, key_names=["event_type"], aggs=[O.Agg("count", None, "count")],
                            window_index=1)
    op = ops.TumblingAggregatingWindowFunc(cfg)
    b = O.source_batches({"event_type": inputs["cars_event_type"], O.TIMESTAMP: inputs["cars_ts"]}, BATCH)
    out = ops.run_single_input(op, b).all()
    return _rows(out, {"event_type": "event_type", "hour": "window_start", "count": "count"})


def tight_watermark(ops, inputs):
    """WATERMARK FOR timestamp (no delay); TUMBLE(1 h) COUNT(*) unkeyed; emits window.end."""
    cfg = O.WindowAggConfig(width=HOUR, key_names=[], aggs=[O.Agg("count", None, "count")], window_index=0)
    op = ops.TumblingAggregatingWindowFunc(cfg)
    b = O.source_batches({O.TIMESTAMP: inputs["cars_ts"]}, BATCH)
    out = ops.run_single_input(op, b, delay_ns=0).all()
    return _rows(out, {"count": "count", "timestamp": "window_end"})


def month_loose_watermark(ops, inputs):
    """watermark = timestamp - 1 min; TUMBLE(30 days) COUNT(*) unkeyed."""
    cfg = O.WindowAggConfig(width=30 * DAY, key_names=[], aggs=[O.Agg("count", None, "count")], window_index=0)
    op = ops.TumblingAggregatingWindowFunc(cfg)
    b = O.source_batches({O.TIMESTAMP: inputs["cars_ts"]}, BATCH)
    out = ops.run_single_input(op, b, delay_ns=MIN).all()
    return _rows(out, {"month": "window_start", "count": "count"})


def most_active_driver_last_hour(ops, inputs):
    """hop(1 min, 1 h) count(*) GROUP BY driver_id, then ROW_NUMBER() OVER (PARTITION BY window
    ORDER BY count DESC, driver_id DESC) = 1 (the window function is outside the hot path and is
    applied here on the host)."""
    cfg = O.WindowAggConfig(width=HOUR, slide=MIN, key_names=["driver_id"],
                            aggs=[O.Agg("count", None, "count")], window_index=1)
    op = ops.SlidingAggregatingWindowFunc(cfg)
    b = O.source_batches({"driver_id": inputs["cars_driver_id"], O.TIMESTAMP: inputs["cars_ts"]}, BATCH)
    out = ops.run_single_input(op, b, delay_ns=HOUR).all()
    best = {}
    for r in out.rows():
        w = (r["window_start"], r["window_end"])
        cand = (r["count"], r["driver_id"])
        if w not in best or cand > best[w]:
            best[w] = cand
    return [{"start": w[0], "end": w[1], "driver_id": d, "count": c, "row_number": 1}
            for w, (c, d) in best.items()]


def session_window(ops, inputs):
    """SESSION(20 s) count(*) GROUP BY user_id = CASE WHEN counter % 10 = 0 THEN 0 ELSE counter END."""
    counter = inputs["impulse_counter"]
    user = np.where(counter % 10 == 0, 0, counter).astype(np.int64)
    cfg = O.SessionConfig(gap=20 * S, key_names=["user_id"], aggs=[O.Agg("count", None, "rows")], window_index=0)
    op = ops.SessionAggregatingWindowFunc(cfg)
    b = O.source_batches({"user_id": user, O.TIMESTAMP: inputs["impulse_ts"]}, BATCH)
    out = ops.run_single_input(op, b).all()
    return _rows(out, {"start": "window_start", "end": "window_end", "user_id": "user_id", "rows": "rows"})


def global_session_window(ops, inputs):
    cfg = O.SessionConfig(gap=20 * S, key_names=[], aggs=[O.Agg("count", None, "rows")], window_index=0)
    op = ops.SessionAggregatingWindowFunc(cfg)
    b = O.source_batches({O.TIMESTAMP: inputs["impulse_ts"]}, BATCH)
    out = ops.run_single_input(op, b, delay_ns=0).all()
    return _rows(out, {"start": "window_start", "end": "window_end", "rows": "rows"})


def _drive_join(ops, join, left_out, right_out):
    """Feed two already-windowed streams (lists of output batches in emission order) into an
    InstantJoin.  Upstream window operators stamp every row of a window with one `_timestamp`
    and forward their watermark after the rows it released; here each side delivers all its
    batches, then the final watermark (order across sides does not matter to the join)."""
    ctx = ops.join_ctx() if hasattr(ops, "join_ctx") else O.OperatorContext(2)
    out = O.Collector()
    for side, stream in ((0, left_out), (1, right_out)):
        for b in stream:
            join.process_batch_index(side, 2, b, ctx, out)
        if side == 0 and hasattr(ops, "restart_join"):
            join = ops.restart_join(join, ctx)  # checkpoint + restore between the two input streams
    for side in (0, 1):
        ctx.watermarks.set(side, O.FINAL_WATERMARK)
    join.handle_watermark(O.FINAL_WATERMARK, ctx, out)
    return out.all()


def _distinct_drivers_per_hour(ops, inputs, event_code):
    """COUNT(DISTINCT driver_id) per TUMBLE(1 h): distinct accumulators are outside the supported
    aggregate subset, so the two-level form is used: tumbling COUNT(*) GROUP BY driver_id (operator
    under test), then a host-side count of groups per window."""
    sel = inputs["cars_event_type"] == event_code
    cfg = O.WindowAggConfig(width=HOUR, key_names=["driver_id"], aggs=[O.Agg("count", None, "n")], window_index=1)
    op = ops.TumblingAggregatingWindowFunc(cfg)
    # the filter runs after the source batching (WHERE is evaluated downstream of the source)
    full = O.source_batches({"driver_id": inputs["cars_driver_id"], "ev": inputs["cars_event_type"],
                             O.TIMESTAMP: inputs["cars_ts"]}, BATCH)
    ctx = O.OperatorContext(1)
    out = O.Collector()
    gen = O.WatermarkGenerator()
    for fb in full:
        keep = fb["ev"] == event_code
        b = O.Batch({"driver_id": fb["driver_id"][keep], O.TIMESTAMP: fb[O.TIMESTAMP][keep]})
        if b.num_rows:
            op.process_batch(b, ctx, out)
        wm = gen.process_batch(fb[O.TIMESTAMP])
        if wm is not None:
            ctx.watermarks.set(0, wm)
            op.handle_watermark(wm, ctx, out)
    ctx.watermarks.set(0, O.FINAL_WATERMARK)
    op.handle_watermark(O.FINAL_WATERMARK, ctx, out)
    res = []
    for wb in out.batches:
        ws = int(wb["window_start"][0])
        res.append(O.Batch({"window_start": np.array([ws], dtype=np.int64),
                            "drivers": np.array([wb.num_rows], dtype=np.int64),
                            O.TIMESTAMP: wb[O.TIMESTAMP][:1].copy()}))
    return res


def _windowed_join(ops, inputs, join_type):
    dropoffs = _distinct_drivers_per_hour(ops, inputs, 1)
    pickups = _distinct_drivers_per_hour(ops, inputs, 0)
    pickups = [O.Batch({"window_start": b["window_start"], "pickups": b["drivers"], O.TIMESTAMP: b[O.TIMESTAMP]})
               for b in pickups]
    join = ops.InstantJoin(O.JoinConfig(left_on=["window_start"], right_on=["window_start"], join_type=join_type))
    out = _drive_join(ops, join, dropoffs, pickups)
    rows = []
    for r in out.rows():
        hour = r["window_start"]  # SELECT dropoffs.window ... window.start
        rows.append({"hour": hour, "drivers": r["drivers"], "pickups": r["pickups"]})
    return rows


def windowed_inner_join(ops, inputs):
    return _windowed_join(ops, inputs, "inner")


def windowed_outer_join(ops, inputs):
    return _windowed_join(ops, inputs, "full")


def offset_impulse_join(ops, inputs):
    """Two TUMBLE(1 s) count(*) GROUP BY counter streams (watermark delays 0 and 10 min) joined
    ON a.counter = b.counter per window instant."""
    def side(delay):
        cfg = O.WindowAggConfig(width=S, key_names=["counter"], aggs=[O.Agg("count", None, "n")], window_index=0)
        op = ops.TumblingAggregatingWindowFunc(cfg)
        b = O.source_batches({"counter": inputs["impulse_counter"], O.TIMESTAMP: inputs["impulse_ts"]}, BATCH)
        return ops.run_single_input(op, b, delay_ns=delay).batches
    a = side(0)
    bb = [O.Batch({"counter": x["counter"], "b_start": x["window_start"], O.TIMESTAMP: x[O.TIMESTAMP]})
          for x in side(10 * MIN)]
    join = ops.InstantJoin(O.JoinConfig(left_on=["counter"], right_on=["counter"], join_type="inner"))
    out = _drive_join(ops, join, a, bb)
    return _rows(out, {"start": "window_start", "counter": "counter"})


def nexmark_q5(ops, inputs):
    """hop(2 s, 10 s) count(*) GROUP BY auction; joined per window with max(count); keep num >= maxn.
    The per-window MAX (an instant-window aggregate keyed by the window) is computed on the host;
    the sliding aggregate and the instant join are the operators under test."""
    cfg = O.WindowAggConfig(width=10 * S, slide=2 * S, key_names=["auction"],
                            aggs=[O.Agg("count", None, "num")], window_index=1)
    op = ops.SlidingAggregatingWindowFunc(cfg)
    b = O.source_batches({"auction": inputs["bids_auction"], O.TIMESTAMP: inputs["bids_ts"]}, BATCH)
    windows = ops.run_single_input(op, b).batches
    left = [O.Batch({"auction": w["auction"], "num": w["num"], "window_start": w["window_start"],
                     O.TIMESTAMP: w[O.TIMESTAMP]}) for w in windows]
    right = [O.Batch({"w": w["window_start"][:1].copy(), "maxn": np.array([w["num"].max()], dtype=np.int64),
                      O.TIMESTAMP: w[O.TIMESTAMP][:1].copy()}) for w in windows]
    join = ops.InstantJoin(O.JoinConfig(left_on=["window_start"], right_on=["w"], join_type="inner"))
    out = _drive_join(ops, join, left, right)
    return [{"auction": r["auction"], "count": r["num"]} for r in out.rows() if r["num"] >= r["maxn"]]


def _accumulators(ops, inputs, keyed, window):
    """min / max / sum / count / avg over the whole impulse input, per `counter % 5` or globally: the final state the
    reference's updating aggregate reaches (queries grouped_aggregates.sql / aggregates.sql), computed here by the
    windowed operators over one window that contains every row -- same accumulators, pinned by the same vectors."""
    counter = inputs["impulse_counter"]
    cols = {"counter": counter, O.TIMESTAMP: inputs["impulse_ts"]}
    keys = []
    if keyed:
        cols = {"counter_mod": counter % 5, **cols}
        keys = ["counter_mod"]
    aggs = [O.Agg("min", "counter", "min"), O.Agg("max", "counter", "max"), O.Agg("sum", "counter", "sum"),
            O.Agg("count", None, "count"), O.Agg("avg", "counter", "avg")]
    if window == "tumbling":
        op = ops.TumblingAggregatingWindowFunc(O.WindowAggConfig(width=HOUR, key_names=keys, aggs=aggs,
                                                                 window_index=len(keys)))
    else:  # one hop whose window [start, start + 1 h) holds all rows: only the fullest window is compared
        op = ops.SlidingAggregatingWindowFunc(O.WindowAggConfig(width=HOUR, slide=HOUR // 2, key_names=keys, aggs=aggs,
                                                                window_index=len(keys)))
    out = ops.run_single_input(op, O.source_batches(cols, BATCH)).all()
    rows = _rows(out, {**({"counter_mod": "counter_mod"} if keyed else {}), "min": "min", "max": "max", "sum": "sum",
                       "count": "count", "avg": "avg", "_start": "window_start"})
    if window == "sliding":  # rows sit in two overlapping hops; keep the hop that starts on the hour
        rows = [r for r in rows if r["_start"] % HOUR == 0]
    for r in rows:
        del r["_start"]
    return rows


ACCUMULATOR_CASES = {
    ("grouped_aggregates", "tumbling"): lambda ops, inputs: _accumulators(ops, inputs, True, "tumbling"),
    ("grouped_aggregates", "sliding"): lambda ops, inputs: _accumulators(ops, inputs, True, "sliding"),
    ("aggregates", "tumbling"): lambda ops, inputs: _accumulators(ops, inputs, False, "tumbling"),
    ("aggregates", "sliding"): lambda ops, inputs: _accumulators(ops, inputs, False, "sliding"),
}


CASES = {
    "sliding_window_end": sliding_window_end,
    "hourly_by_event_type": hourly_by_event_type,
    "tight_watermark": tight_watermark,
    "month_loose_watermark": month_loose_watermark,
    "most_active_driver_last_hour": most_active_driver_last_hour,
    "session_window": session_window,
    "global_session_window": global_session_window,
    "windowed_inner_join": windowed_inner_join,
    "windowed_outer_join": windowed_outer_join,
    "offset_impulse_join": offset_impulse_join,
    "nexmark_q5": nexmark_q5,
}

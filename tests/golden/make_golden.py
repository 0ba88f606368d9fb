#!/usr/bin/env python
"""Generate the compact golden fixtures under tests/golden/ from the reference's
own smoke-test inputs and golden outputs.

Run in the build container only (the reference tree is not present on the GPU box):

    python tests/golden/make_golden.py [/root/reference]

Source of the vectors (read-only, never copied verbatim):
  crates/arroyo-sql-testing/inputs/{cars,impulse,nexmark_bids}.json     -> inputs.npz
  crates/arroyo-sql-testing/golden_outputs/<query>.json                  -> expected.json
  crates/arroyo-sql-testing/golden_outputs/{grouped_aggregates,aggregates,debezium_agg}.json -> accumulators.json
  crates/arroyo-sql-testing/inputs/aggregate_updates.json                -> updating_inputs.npz
The JSON lines are re-encoded: timestamps become int64 nanoseconds since the Unix
epoch, the `event_type` strings become the int64 codes in EVENT_TYPE_CODES, and the
columns the hot path never reads (location, url, extra, ...) are dropped.
"""
import json
import os
import sys
from datetime import datetime, timezone

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
BASE = os.path.join(REF, "crates", "arroyo-sql-testing")
OUT = os.path.dirname(os.path.abspath(__file__))

EVENT_TYPE_CODES = {"pickup": 0, "dropoff": 1}

EPOCH = datetime(1970, 1, 1, tzinfo=timezone.utc)


def ts_ns(s: str) -> int:
    """ISO-8601 (with or without offset / fractional seconds) -> ns since epoch."""
    s = s.replace("Z", "+00:00")
    if "+" not in s[10:] and "-" not in s[10:]:
        s = s + "+00:00"
    d = datetime.fromisoformat(s)
    delta = d - EPOCH
    return (delta.days * 86400 + delta.seconds) * 1_000_000_000 + delta.microseconds * 1000


def lines(path):
    with open(path) as f:
        return [json.loads(l) for l in f if l.strip()]


def main():
    cars = lines(os.path.join(BASE, "inputs", "cars.json"))
    impulse = lines(os.path.join(BASE, "inputs", "impulse.json"))
    bids = lines(os.path.join(BASE, "inputs", "nexmark_bids.json"))
    np.savez_compressed(
        os.path.join(OUT, "inputs.npz"),
        cars_ts=np.array([ts_ns(r["timestamp"]) for r in cars], dtype=np.int64),
        cars_driver_id=np.array([r["driver_id"] for r in cars], dtype=np.int64),
        cars_event_type=np.array([EVENT_TYPE_CODES[r["event_type"]] for r in cars], dtype=np.int64),
        impulse_ts=np.array([ts_ns(r["timestamp"]) for r in impulse], dtype=np.int64),
        impulse_counter=np.array([r["counter"] for r in impulse], dtype=np.int64),
        impulse_subtask_index=np.array([r["subtask_index"] for r in impulse], dtype=np.int64),
        bids_ts=np.array([ts_ns(r["datetime"]) for r in bids], dtype=np.int64),
        bids_auction=np.array([r["auction"] for r in bids], dtype=np.int64),
    )

    queries = [
        "sliding_window_end", "hourly_by_event_type", "tight_watermark", "month_loose_watermark",
        "most_active_driver_last_hour", "nexmark_q5", "session_window", "global_session_window",
        "windowed_inner_join", "windowed_outer_join", "offset_impulse_join",
    ]
    ts_fields = {"start", "end", "hour", "timestamp", "month"}
    expected = {}
    for q in queries:
        rows = []
        for r in lines(os.path.join(BASE, "golden_outputs", q + ".json")):
            o = {}
            for k, v in r.items():
                if k in ts_fields and v is not None:
                    o[k] = ts_ns(v)
                elif k == "event_type":
                    o[k] = EVENT_TYPE_CODES[v]
                else:
                    o[k] = v
            rows.append(o)
        expected[q] = rows
    with open(os.path.join(OUT, "expected.json"), "w") as f:
        json.dump(expected, f, separators=(",", ":"), sort_keys=True)
    print({q: len(v) for q, v in expected.items()})

    # Updating (non-windowed) aggregates: the sink is a Debezium change stream that the reference's test merges per
    # primary key before comparing (smoke_tests.rs:519-562).  Their merged final rows pin the SUM / AVG / MIN / MAX /
    # COUNT accumulators -- the same DataFusion accumulators the windowed operators use.
    acc = {}
    for q, pk in (("grouped_aggregates", "counter_mod"), ("aggregates", None)):
        state = {}
        for r in lines(os.path.join(BASE, "golden_outputs", q + ".json")):
            op, before, after = r["op"], r.get("before"), r.get("after")
            if op in ("u", "d") and before is not None:
                state.pop(before.get(pk) if pk else 0, None)
            if op in ("c", "u") and after is not None:
                state[after.get(pk) if pk else 0] = after
        acc[q] = [state[k] for k in sorted(state)]
    # debezium_agg.sql: a Debezium change stream (c / u / d with both row images) into an updating aggregate
    # GROUP BY product: count(*), count(distinct customer_name), sum(quantity + 5) + 10.  Strings become codes.
    upd = lines(os.path.join(BASE, "inputs", "aggregate_updates.json"))
    products, customers = {}, {}
    rows = []  # (is_retract, product code, customer code, quantity) in stream order: u = retract(before), append(after)
    for r in upd:
        imgs = []
        if r["op"] in ("u", "d"):
            imgs.append((1, r["before"]))
        if r["op"] in ("c", "u", "r"):
            imgs.append((0, r["after"]))
        for retract, img in imgs:
            rows.append((retract, products.setdefault(img["product_name"], len(products)),
                         customers.setdefault(img["customer_name"], len(customers)), img["quantity"]))
    arr = np.array(rows, dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, "updating_inputs.npz"), is_retract=arr[:, 0], product=arr[:, 1],
                        customer=arr[:, 2], quantity=arr[:, 3])
    state = {}
    for r in lines(os.path.join(BASE, "golden_outputs", "debezium_agg.json")):
        op, before, after = r["op"], r.get("before"), r.get("after")
        if op in ("u", "d") and before is not None:
            state.pop(before["id"], None)
        if op in ("c", "u") and after is not None:
            state[after["id"]] = after
    acc["debezium_agg"] = [{"product": products[v["id"][2:]], "c": v["c"], "d": v["d"], "q": v["q"]}
                           for _, v in sorted(state.items())]
    # updating_inner_join.sql: impulse A JOIN impulse_odd B ON A.counter = B.counter (append-only inputs, inner join)
    acc["updating_inner_join"] = [r["after"] for r in lines(os.path.join(BASE, "golden_outputs", "updating_inner_join.json"))
                                  if r["op"] == "c"]
    with open(os.path.join(OUT, "accumulators.json"), "w") as f:
        json.dump(acc, f, separators=(",", ":"), sort_keys=True)
    print({q: len(v) for q, v in acc.items()})


if __name__ == "__main__":
    main()

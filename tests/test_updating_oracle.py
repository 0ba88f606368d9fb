"""The oracle of the reference's updating (non-windowed) aggregate -- SURVEY.md 8(f) rank 2, no CUDA operator yet --
pinned by the reference's goldens `grouped_aggregates`, `aggregates` and `debezium_agg` on the merged change stream
(the way the reference's own harness compares Debezium sinks), for several flush cadences, and checked against a brute
force recomputation on random append / retract streams."""
import os

import numpy as np
import pytest

from oracle import arroyo_oracle as O
from oracle import updating_oracle as U

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BATCH = 32


def run(op, batches, flush_every):
    out = []
    for i, b in enumerate(batches):
        op.process_batch(b)
        if flush_every and (i + 1) % flush_every == 0:
            out.append(op.flush())
    out.append(op.flush())  # end of data
    return out


@pytest.mark.parametrize("flush_every", [1, 3, 0])
def test_grouped_aggregates_golden(golden, accumulator_golden, flush_every):
    counter, ts = golden[0]["impulse_counter"], golden[0]["impulse_ts"]
    cfg = U.UpdatingAggConfig(["counter_mod"], [O.Agg("min", "counter", "min"), O.Agg("max", "counter", "max"),
                                                O.Agg("sum", "counter", "sum"), O.Agg("count", None, "count"),
                                                O.Agg("avg", "counter", "avg")])
    b = O.source_batches({"counter_mod": counter % 5, "counter": counter, O.TIMESTAMP: ts}, BATCH)
    out = run(U.IncrementalAggregatingFunc(cfg), b, flush_every)
    assert U.merge_change_stream(out, ["counter_mod"]) == accumulator_golden["grouped_aggregates"]
    if flush_every == 1:  # every later flush retracts what the previous one said about the keys it touches
        assert sum(int(x[U.IS_RETRACT].sum()) for x in out if x is not None) > 0


@pytest.mark.parametrize("flush_every", [1, 2, 0])
def test_global_aggregates_golden(golden, accumulator_golden, flush_every):
    counter, ts = golden[0]["impulse_counter"], golden[0]["impulse_ts"]
    cfg = U.UpdatingAggConfig([], [O.Agg("min", "counter", "min"), O.Agg("max", "counter", "max"),
                                   O.Agg("sum", "counter", "sum"), O.Agg("count", None, "count"),
                                   O.Agg("avg", "counter", "avg")])
    out = run(U.IncrementalAggregatingFunc(cfg), O.source_batches({"counter": counter, O.TIMESTAMP: ts}, BATCH), flush_every)
    assert U.merge_change_stream(out, []) == accumulator_golden["aggregates"]


@pytest.mark.parametrize("flush_every", [1, 4, 0])
def test_debezium_agg_golden(accumulator_golden, flush_every):
    """A Debezium source: updates arrive as retract(before) + append(after), deletes as retract(before)."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "updating_inputs.npz"))
    n = len(z["product"])
    cols = {"product": z["product"], "customer": z["customer"], "qp5": z["quantity"] + 5,
            U.IS_RETRACT: z["is_retract"].astype(bool), O.TIMESTAMP: np.arange(n, dtype=np.int64)}
    cfg = U.UpdatingAggConfig(["product"], [O.Agg("count", None, "c"), O.Agg("count_distinct", "customer", "d"),
                                            O.Agg("sum", "qp5", "q")])
    out = run(U.IncrementalAggregatingFunc(cfg), O.source_batches(cols, BATCH), flush_every)
    got = U.merge_change_stream(out, ["product"])
    for r in got:
        r["q"] += 10  # the projection above the aggregate: sum(quantity + 5) + 10
    key = lambda r: r["product"]  # noqa: E731
    assert sorted(got, key=key) == sorted(accumulator_golden["debezium_agg"], key=key)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_random_retractions_against_brute_force(seed):
    rng = np.random.default_rng(seed)
    live = []  # rows currently in the table: (key, value, ts)
    batches = []
    t = 0
    for _ in range(40):
        rows = []
        for _ in range(int(rng.integers(1, 30))):
            if live and rng.random() < 0.35:
                k, v, ts = live.pop(int(rng.integers(0, len(live))))
                rows.append((k, v, ts, True))
            else:
                t += 1
                r = (int(rng.integers(0, 6)), int(rng.integers(-50, 50)), t)
                live.append(r)
                rows.append((*r, False))
        a = np.array([(k, v, ts, int(x)) for k, v, ts, x in rows], dtype=np.int64)
        batches.append(O.Batch({"k": a[:, 0], "v": a[:, 1], O.TIMESTAMP: a[:, 2], U.IS_RETRACT: a[:, 3].astype(bool)}))
    aggs = [O.Agg("count", None, "n"), O.Agg("sum", "v", "s"), O.Agg("avg", "v", "a"), O.Agg("min", "v", "lo"),
            O.Agg("max", "v", "hi"), O.Agg("count_distinct", "v", "d")]
    want = {}
    for k in {r[0] for r in live}:
        vs = [v for kk, v, _ in live if kk == k]
        want[k] = {"k": k, "n": len(vs), "s": sum(vs), "a": sum(vs) / len(vs), "lo": min(vs), "hi": max(vs), "d": len(set(vs))}
    for flush_every in (1, 5, 0):
        out = run(U.IncrementalAggregatingFunc(U.UpdatingAggConfig(["k"], aggs)), batches, flush_every)
        got = {r["k"]: r for r in U.merge_change_stream(out, ["k"])}
        assert got == want


@pytest.mark.parametrize("order", ["left_first", "right_first", "alternating"])
def test_updating_inner_join_golden(golden, accumulator_golden, order):
    """SURVEY.md 8(f) rank 3 (inner case): impulse A JOIN impulse_odd B ON A.counter = B.counter -- every matching
    pair leaves exactly once, whichever side's row arrives later."""
    from tests.golden_cases import multiset
    counter, ts = golden[0]["impulse_counter"], golden[0]["impulse_ts"]
    odd = counter % 2 == 1
    left = O.source_batches({"counter": counter, O.TIMESTAMP: ts}, BATCH)
    right = O.source_batches({"counter": counter[odd], O.TIMESTAMP: ts[odd]}, BATCH)
    join = U.JoinWithExpiration("counter", "counter")
    feed = {"left_first": [(0, b) for b in left] + [(1, b) for b in right],
            "right_first": [(1, b) for b in right] + [(0, b) for b in left],
            "alternating": [x for pair in zip([(0, b) for b in left], [(1, b) for b in right] + [None] * len(left)) for x in pair if x]}[order]
    out = []
    for side, b in feed:
        out += join.process_batch_index(side, 2, b)
    got = [{"left_count": r["counter"], "right_count": r["counter_right"]} for r in out]
    assert multiset(got) == multiset(accumulator_golden["updating_inner_join"])

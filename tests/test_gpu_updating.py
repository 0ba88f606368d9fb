"""The CUDA updating aggregate (csrc/updating_agg.cu = IncrementalAggregatingFunc, incremental_aggregator.rs) against
the reference's goldens `grouped_aggregates` / `aggregates` on the merged change stream (the way the reference's own
harness compares Debezium sinks, smoke_tests.rs:519-562) and, flush by flush, against the oracle's change rows
(oracle/updating_oracle.py): same retractions, same appends."""
import numpy as np
import pytest

from oracle import arroyo_oracle as O
from oracle import updating_oracle as U
from tests.golden_cases import multiset

pytestmark = pytest.mark.gpu
BATCH = 32
AGGS = [O.Agg("min", "counter", "min"), O.Agg("max", "counter", "max"), O.Agg("sum", "counter", "sum"),
        O.Agg("count", None, "count"), O.Agg("avg", "counter", "avg")]


class _GpuOp:
    def __init__(self, cfg, **kw):
        import arroyo_b200 as ab
        from arroyo_b200 import operators as native
        self.ab = ab
        self.op = native.UpdatingAggregatingFunc(cfg, **kw)
        self.cfg = cfg
        self.ctx = ab.OperatorContext(1)

    def process_batch(self, batch):
        from tests.gpu_ops import to_arrow
        self.op.process_batch(to_arrow(batch), self.ctx, None)

    def flush(self):
        from tests.gpu_ops import from_arrow
        col = self.ab.Collector()
        self.op.handle_tick(0, self.ctx, col)
        if not col.batches:
            return None
        assert len(col.batches) == 1
        b = from_arrow(col.batches[0])
        cols = dict(b.cols)
        cols[U.IS_RETRACT] = cols.pop("_is_retract").astype(bool)
        return O.Batch(cols)


def run(op, batches, flush_every):
    out = []
    for i, b in enumerate(batches):
        op.process_batch(b)
        if flush_every and (i + 1) % flush_every == 0:
            out.append(op.flush())
    out.append(op.flush())
    return out


def _rows(b):
    if b is None:
        return []
    out = []
    for r in b.rows():
        out.append({k: (round(float(v), 9) if isinstance(v, (float, np.floating)) else (bool(v) if isinstance(v, (bool, np.bool_)) else int(v)))
                    for k, v in r.items()})
    return out


@pytest.mark.parametrize("flush_every", [1, 3, 0])
@pytest.mark.parametrize("keyed", [True, False])
def test_updating_aggregate_goldens_and_change_rows(golden, accumulator_golden, keyed, flush_every):
    counter, ts = golden[0]["impulse_counter"], golden[0]["impulse_ts"]
    keys = ["counter_mod"] if keyed else []
    cols = {"counter": counter, O.TIMESTAMP: ts}
    if keyed:
        cols = {"counter_mod": counter % 5, **cols}
    batches = O.source_batches(cols, BATCH)
    cfg = U.UpdatingAggConfig(keys, AGGS)
    want = run(U.IncrementalAggregatingFunc(cfg), batches, flush_every)
    got = run(_GpuOp(cfg), batches, flush_every)
    assert U.merge_change_stream(got, keys) == accumulator_golden["grouped_aggregates" if keyed else "aggregates"]
    assert len(got) == len(want)
    for g, w in zip(got, want):  # flush by flush: the same retractions and the same appends
        assert multiset(_rows(g)) == multiset(_rows(w))
        if g is not None:  # a key's retraction precedes its append
            r = g[U.IS_RETRACT]
            assert not r[int(r.sum()):].any()


def test_updating_aggregate_random_streams_growth_and_idle_keys():
    """60 000 keys through a dictionary sized for 256 (it has to grow, ids are permuted, the touched list follows);
    keys whose aggregates do not change between two flushes (only max(_timestamp) moves: SUM of zeros, MIN / MAX
    unchanged) must not be re-emitted (incremental_aggregator.rs:655-664)."""
    rng = np.random.default_rng(5)
    cfg = U.UpdatingAggConfig(["k"], [O.Agg("count", None, "n"), O.Agg("sum", "v", "s"), O.Agg("avg", "v", "a"),
                                      O.Agg("min", "v", "lo"), O.Agg("max", "v", "hi")])
    quiet = U.UpdatingAggConfig(["k"], [O.Agg("sum", "v", "s"), O.Agg("min", "v", "lo"), O.Agg("max", "v", "hi")])
    batches, t = [], 0
    for i in range(30):
        n = int(rng.integers(1000, 8000))
        k = rng.integers(0, 60_000, n, dtype=np.int64) * 7919 - 3
        v = rng.integers(-10**6, 10**6, n, dtype=np.int64)
        if i % 3 == 2:
            v[:] = 0  # sums, minima and maxima of the touched keys mostly stay what they were
        batches.append(O.Batch({"k": k, "v": v, O.TIMESTAMP: np.arange(t, t + n, dtype=np.int64)}))
        t += n
    for c in (cfg, quiet):
        want = run(U.IncrementalAggregatingFunc(c), batches, 2)
        gop = _GpuOp(c, expected_keys=256)
        got = run(gop, batches, 2)
        for g, w in zip(got, want):
            assert multiset(_rows(g)) == multiset(_rows(w))
        assert gop.op.stats()["n_keys"] == len({int(x) for b in batches for x in b["k"]})


def test_updating_input_is_refused():
    from arroyo_b200 import ffi
    import pyarrow as pa
    from arroyo_b200 import operators as native
    cfg = U.UpdatingAggConfig(["k"], [O.Agg("count", None, "n")])
    with pytest.raises(ffi.UnsupportedPlan):
        native.UpdatingAggregatingFunc(cfg, input_schema=pa.schema([("k", pa.int64()), ("_timestamp", pa.timestamp("ns"))]),
                                       updating_input=True)

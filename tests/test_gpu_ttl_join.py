"""The CUDA join with expiration (csrc/ttl_join.cu = JoinWithExpiration, join_with_expiration.rs) against the
reference's golden `updating_inner_join` (three arrival orders: every pair leaves exactly once, whichever side's row
arrives later) and against the oracle (oracle/updating_oracle.py::JoinWithExpiration) on random streams with duplicate
keys on both sides, batch by batch."""
import numpy as np
import pytest

from oracle import arroyo_oracle as O
from oracle import updating_oracle as U
from tests.golden_cases import multiset

pytestmark = pytest.mark.gpu
BATCH = 32


class _GpuJoin:
    def __init__(self, left_on, right_on):
        import arroyo_b200 as ab
        from arroyo_b200 import operators as native
        self.ab = ab
        self.op = native.JoinWithExpiration(O.JoinConfig(left_on=[left_on], right_on=[right_on], join_type="inner"))
        self.ctx = ab.OperatorContext(2)

    def process_batch_index(self, index, total, batch):
        from tests.gpu_ops import from_arrow, to_arrow
        col = self.ab.Collector()
        self.op.process_batch_index(index, total, to_arrow(batch), self.ctx, col)
        return [r for b in col.batches for r in from_arrow(b).rows()]


@pytest.mark.parametrize("order", ["left_first", "right_first", "alternating"])
def test_updating_inner_join_golden(golden, accumulator_golden, order):
    counter, ts = golden[0]["impulse_counter"], golden[0]["impulse_ts"]
    odd = counter % 2 == 1
    left = O.source_batches({"counter": counter, O.TIMESTAMP: ts}, BATCH)
    right = O.source_batches({"counter": counter[odd], O.TIMESTAMP: ts[odd]}, BATCH)
    feed = {"left_first": [(0, b) for b in left] + [(1, b) for b in right],
            "right_first": [(1, b) for b in right] + [(0, b) for b in left],
            "alternating": [x for pair in zip([(0, b) for b in left], [(1, b) for b in right] + [None] * len(left)) for x in pair if x]}[order]
    join = _GpuJoin("counter", "counter")
    out = []
    for side, b in feed:
        out += join.process_batch_index(side, 2, b)
    got = [{"left_count": r["counter"], "right_count": r["counter_right"]} for r in out]
    assert multiset(got) == multiset(accumulator_golden["updating_inner_join"])


def test_random_streams_match_the_oracle_batch_by_batch():
    rng = np.random.default_rng(3)
    oracle, gpu = U.JoinWithExpiration("k", "k2"), _GpuJoin("k", "k2")
    t = 0
    total = 0
    for step in range(60):
        side = int(rng.integers(0, 2))
        n = int(rng.integers(1, 4000))
        keys = rng.integers(0, 500, n, dtype=np.int64) * 104729 - 7  # duplicates on both sides, inside and across batches
        ts = np.arange(t, t + n, dtype=np.int64)
        t += n
        if side == 0:
            b = O.Batch({"k": keys, "a": rng.integers(-10**9, 10**9, n, dtype=np.int64), O.TIMESTAMP: ts})
        else:
            b = O.Batch({"k2": keys, "b": rng.integers(-10**9, 10**9, n, dtype=np.int64), "c": rng.integers(0, 9, n, dtype=np.int64),
                         O.TIMESTAMP: ts})
        want = oracle.process_batch_index(side, 2, b)
        got = gpu.process_batch_index(side, 2, b)
        norm = lambda rows: multiset([{k: int(v) for k, v in r.items()} for r in rows])  # noqa: E731
        assert norm(got) == norm(want), step
        total += len(want)
    assert total > 100_000
    st = gpu.op.stats()
    assert st["rows_out"] == total and st["kernel_launches"] > 0


def test_outer_join_is_refused():
    from arroyo_b200 import ffi
    from arroyo_b200 import operators as native
    import pyarrow as pa
    sch = pa.schema([("k", pa.int64()), ("_timestamp", pa.timestamp("ns"))])
    with pytest.raises(ffi.UnsupportedPlan):
        native.JoinWithExpiration(O.JoinConfig(left_on=["k"], right_on=["k"], join_type="left"), left_schema=sch, right_schema=sch)

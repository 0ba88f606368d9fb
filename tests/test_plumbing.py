"""BASELINE configs[0]: the stateless CPU plumbing (Nexmark q1 / q2 over 64 Ki-row Arrow batches) and the run loop's
barrier alignment -- correctness against plain numpy / the reference's own unit expectations."""
import numpy as np
import pyarrow as pa

from arroyo_b200 import plumbing as P
from arroyo_b200.context import Collector

S = 1_000_000_000


def bids(n, seed=42):
    rng = np.random.default_rng(seed)
    price = np.floor(np.power(10.0, rng.random(n) * 6.0) * 100.0).astype(np.int64)  # nexmark/operator.rs:643-645
    return pa.RecordBatch.from_arrays(
        [pa.array(rng.integers(1000, 2000, n)), pa.array(rng.integers(1000, 50_000, n)), pa.array(price),
         pa.array(1_700_000_000 * S + np.arange(n, dtype=np.int64) * 1000, type=pa.int64()).cast(pa.timestamp("ns"))],
        names=["auction", "bidder", "price", "_timestamp"])


def test_nexmark_q1_projection_matches_numpy():
    b, out = bids(65_536), Collector()
    P.nexmark_q1().process_batch(b, None, out)
    (o,) = out.batches
    assert o.schema.names == ["auction", "bidder", "price", "_timestamp"] and o.num_rows == b.num_rows
    np.testing.assert_array_equal(o.column(0).to_numpy(), b.column(0).to_numpy())
    np.testing.assert_array_equal(o.column(1).to_numpy(), b.column(1).to_numpy())
    np.testing.assert_array_equal(o.column(2).to_numpy(), b.column(2).to_numpy().astype(np.float64) * 0.908)
    assert o.column(3).equals(b.column(3))


def test_nexmark_q2_filter_matches_numpy_and_keeps_empty_batches():
    b, out = bids(65_536, seed=7), Collector()
    op = P.nexmark_q2(123)
    op.process_batch(b, None, out)
    keep = b.column(0).to_numpy() % 123 == 0
    assert out.batches[0].num_rows == int(keep.sum()) > 0
    np.testing.assert_array_equal(out.batches[0].column(1).to_numpy(), b.column(2).to_numpy()[keep])
    none = pa.RecordBatch.from_arrays([pa.array([1, 2]), pa.array([1, 1]), pa.array([5, 6]),
                                       pa.array([0, 1], type=pa.int64()).cast(pa.timestamp("ns"))], names=b.schema.names)
    op.process_batch(none, None, out)
    assert out.batches[1].num_rows == 0 and out.batches[1].schema.names == ["auction", "price", "_timestamp"]


def test_key_projection_prepends_routing_copies():
    b, out = bids(1000), Collector()
    P.KeyExecutionOperator("key", [P.col("bidder")], ["_key_bidder"]).process_batch(b, None, out)
    o = out.batches[0]
    assert o.schema.names == ["_key_bidder", "auction", "bidder", "price", "_timestamp"]
    assert o.column(0).equals(o.column(2))


def test_checkpoint_counter_aligns_barriers_like_the_reference():
    c = P.CheckpointCounter(3)
    assert c.all_clear() and not c.is_blocked(0)
    assert c.mark(1, 7) is False and c.is_blocked(1) and not c.is_blocked(0)
    assert c.mark(0, 7) is False and c.is_blocked(0)
    assert c.mark(2, 7) is True and c.all_clear()          # the last input completes the barrier and unblocks all
    assert c.mark(2, 8) is False and c.is_blocked(2)       # the next epoch starts over
    single = P.CheckpointCounter(1)
    assert single.mark(0, 1) is True and single.all_clear()  # one input: never blocked (lib.rs:94-96)

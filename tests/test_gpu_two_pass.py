"""The two-pass ingest (radix partition by dictionary bucket -> per-bucket aggregation in shared memory,
csrc/ingest_two_pass.cuh) against the oracle, forced on for every launch (FLAG_TWO_PASS_ALWAYS) on streams small enough
for the numpy oracle and shaped to hit every hand-off between the fast path and the one-pass path:

  uniform / one hot key (75 % of the rows) / ten keys (several blocks per bucket, flush through REDs)
  tiles that straddle pane boundaries, late rows, rows far ahead of the pane ring
  a dictionary sized far too small (buckets run out of ids -> rows deferred -> bucket count doubles, ids permuted)
  the key that equals the dictionary's empty sentinel, negative values, a value beyond the exact-AVG guard
  COUNT(*)-only (no value column) and SUM-only plans, tumbling and sliding
and FLAG_NO_TWO_PASS (the one-pass kernel on the same bucketed dictionary)."""
import numpy as np
import pytest

from oracle import arroyo_oracle as O
from tests.test_gpu_parity import S, SUM_AVG, T0, assert_same, gen_stream, run_both

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G():
    from tests import gpu_ops
    return gpu_ops


def _flags(mode):
    from arroyo_b200 import ffi
    return {"two_pass": ffi.FLAG_TWO_PASS_ALWAYS, "one_pass": ffi.FLAG_NO_TWO_PASS,
            "two_pass_no_combine": ffi.FLAG_TWO_PASS_ALWAYS | ffi.FLAG_NO_COMBINE}[mode]


CFG = O.WindowAggConfig(width=4 * S, slide=S, key_names=["key"], aggs=SUM_AVG, window_index=1)


@pytest.mark.parametrize("mode", ["two_pass", "one_pass", "two_pass_no_combine"])
@pytest.mark.parametrize("dist,n_keys", [("uniform", 40_000), ("hot", 40_000), ("uniform", 10), ("uniform", 3_000)])
def test_sliding_sum_avg_count(G, mode, dist, n_keys):
    rng = np.random.default_rng(11)
    # 20 000 rows per second of event time, batches of 32 768: every launch spans more than one pane
    batches = gen_stream(rng, 400_000, n_keys, rate_per_s=20_000, key_dist=dist, batch=32_768)
    want, got, gop = run_both(G, lambda: O.SlidingAggregatingWindowFunc(CFG),
                              lambda: G.SlidingAggregatingWindowFunc(CFG, flags=_flags(mode), expected_keys=n_keys), batches)
    assert_same(want, got, float_cols=("avg",))
    st = gop.stats()
    assert st["rows_in"] == 400_000 and st["rows_late"] == 0


@pytest.mark.parametrize("mode", ["two_pass", "one_pass"])
def test_tumbling_count_only_and_sum_only(G, mode):
    rng = np.random.default_rng(12)
    batches = gen_stream(rng, 300_000, 10_000, rate_per_s=100_000, batch=65_536)
    for aggs in ([O.Agg("count", None, "n")], [O.Agg("sum", "value", "s")]):
        cfg = O.WindowAggConfig(width=S, key_names=["key"], aggs=aggs, window_index=1)
        want, got, _ = run_both(G, lambda: O.TumblingAggregatingWindowFunc(cfg),
                                lambda: G.TumblingAggregatingWindowFunc(cfg, flags=_flags(mode), expected_keys=10_000), batches)
        assert_same(want, got)


@pytest.mark.parametrize("mode", ["two_pass", "one_pass"])
def test_late_rows_far_future_rows_and_a_dictionary_that_has_to_grow(G, mode):
    rng = np.random.default_rng(13)
    batches = gen_stream(rng, 300_000, 60_000, rate_per_s=50_000, batch=16_384)
    # late rows: 500 rows two panes behind the watermark inside batch 9; far future: 100 rows 500 panes ahead in batch 3
    b9 = batches[9]
    old = np.full(500, int(b9[O.TIMESTAMP].min()) - 4 * S, dtype=np.int64)
    batches[9] = O.Batch({"key": np.concatenate([b9["key"], np.arange(500, dtype=np.int64)]),
                          "value": np.concatenate([b9["value"], np.ones(500, dtype=np.int64)]),
                          O.TIMESTAMP: np.concatenate([b9[O.TIMESTAMP], old])})
    b3 = batches[3]
    batches[3] = O.Batch({"key": np.concatenate([b3["key"], np.arange(100, dtype=np.int64)]),
                          "value": np.concatenate([b3["value"], np.ones(100, dtype=np.int64)]),
                          O.TIMESTAMP: np.concatenate([b3[O.TIMESTAMP], np.full(100, T0 + 500 * S, dtype=np.int64)])})
    want, got, gop = run_both(G, lambda: O.SlidingAggregatingWindowFunc(CFG),
                              lambda: G.SlidingAggregatingWindowFunc(CFG, flags=_flags(mode), expected_keys=256), batches)
    assert_same(want, got, float_cols=("avg",))
    st = gop.stats()
    assert st["rows_deferred"] > 0 and st["n_keys"] >= 50_000


@pytest.mark.parametrize("mode", ["two_pass", "one_pass"])
def test_sentinel_key_negative_values_and_the_avg_guard(G, mode):
    rng = np.random.default_rng(14)
    batches = gen_stream(rng, 200_000, 5_000, rate_per_s=40_000, batch=32_768)
    out = []
    for i, b in enumerate(batches):
        key, val = b["key"].copy(), b["value"].copy()
        key[::101] = -2**63            # the dictionary's empty sentinel is a legal key
        key[1::103] = 2**63 - 1
        if i == 3:
            val[7] = 2**40             # beyond the exact-AVG guard: the operator promotes itself to f64 AVG accumulators
        out.append(O.Batch({"key": key, "value": val, O.TIMESTAMP: b[O.TIMESTAMP]}))
    want, got, _ = run_both(G, lambda: O.SlidingAggregatingWindowFunc(CFG),
                            lambda: G.SlidingAggregatingWindowFunc(CFG, flags=_flags(mode), expected_keys=5_000), out)
    assert_same(want, got, float_cols=("avg",))
    assert any((b["key"] == -2**63).any() for b in got)


def test_two_pass_survives_a_checkpoint_and_restore(G):
    import arroyo_b200 as ab
    from arroyo_b200 import operators as native
    from tests.gpu_ops import from_arrow, to_arrow
    rng = np.random.default_rng(15)
    batches = gen_stream(rng, 300_000, 20_000, rate_per_s=30_000, batch=32_768)
    want = O.run_single_input(O.SlidingAggregatingWindowFunc(CFG), batches, S).batches
    schema = to_arrow(batches[0]).schema
    ctx, out, gen = ab.OperatorContext(1), ab.Collector(), ab.WatermarkGenerator(S)
    mk = lambda: native.SlidingAggregatingWindowFunc(CFG, input_schema=schema, flags=_flags("two_pass"), expected_keys=20_000)  # noqa: E731
    op = mk()
    for i, b in enumerate(batches):
        if i == len(batches) // 2:
            op.handle_checkpoint(None, ctx, out)
            op.close()
            op = mk()
            op.on_start(ctx)
        op.process_batch(to_arrow(b), ctx, out)
        wm = gen.on_batch(int(b[O.TIMESTAMP].min()), int(b[O.TIMESTAMP].max()))
        if wm is not None:
            ctx.watermarks.set(0, wm)
            op.handle_watermark(wm, ctx, out)
    ctx.watermarks.set(0, ab.FINAL_WATERMARK)
    op.handle_watermark(ab.FINAL_WATERMARK, ctx, out)
    assert_same(want, [from_arrow(b) for b in out.batches], float_cols=("avg",))

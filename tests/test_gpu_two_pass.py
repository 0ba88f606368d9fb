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


def test_one_large_launch_with_wide_and_negative_values_against_a_group_by(G):
    """One 2^24-row launch over 100 000 keys, SUM + COUNT (no AVG, so no value guard): 128 buckets x 3 blocks, i.e.
    ~44 000 rows per aggregation block -- more than the packed row-count / carry word holds between two flushes -- and
    values that are full-range, negative, or just around the 32-bit boundaries (the wide-value path, carries, borrows).
    Checked row by row against an independent group-by (torch.unique + index_add, which wraps like i64 SUM)."""
    import pyarrow as pa
    import torch

    import arroyo_b200 as ab
    from arroyo_b200 import ffi, operators as native
    from arroyo_b200.multi_gpu import _Ptr

    device = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    torch.cuda.set_stream(torch.cuda.Stream(device=device))
    g = torch.Generator(device=device)
    g.manual_seed(5)
    n, n_keys = 1 << 24, 100_000
    ids = torch.randint(0, n_keys, (n,), generator=g, device=device, dtype=torch.int64)
    key = ids * 0x1E3779B97F4A7C15 % (1 << 62) - (1 << 61)  # scattered (the product wraps), some negative
    val = torch.randint(-(1 << 62), 1 << 62, (n,), generator=g, device=device, dtype=torch.int64) * 2
    kind = torch.randint(0, 4, (n,), generator=g, device=device)
    small = torch.randint(-(1 << 31), 1 << 31, (n,), generator=g, device=device, dtype=torch.int64)
    edge = torch.tensor([-1, -(1 << 31), (1 << 31) - 1, 1 << 31, -(1 << 31) - 1, (1 << 32) - 1, -(1 << 32), 0],
                        device=device)[torch.randint(0, 8, (n,), generator=g, device=device)]
    val = torch.where(kind == 0, val, torch.where(kind == 1, edge, small))
    ts = T0 + torch.randint(0, S, (n,), generator=g, device=device, dtype=torch.int64)
    cfg = ab.WindowAggConfig(width=S, key_names=["key"], aggs=[ab.Agg("sum", "value", "sum"), ab.Agg("count", None, "n")],
                             window_index=1)
    schema = pa.schema([("key", pa.int64()), ("value", pa.int64()), ("_timestamp", pa.timestamp("ns"))])
    op = native.TumblingAggregatingWindowFunc(cfg, input_schema=schema, device=0,
                                              stream=torch.cuda.current_stream().cuda_stream,
                                              flags=ffi.FLAG_TWO_PASS_ALWAYS, expected_keys=n_keys)
    # a first small batch tells the operator where the stream is (the two passes need a known newest pane)
    op.process_device_batch([key.data_ptr(), val.data_ptr(), ts.data_ptr()], 4096)
    op.flush()
    op.process_device_batch([key.data_ptr() + 8 * 4096, val.data_ptr() + 8 * 4096, ts.data_ptr() + 8 * 4096], n - 4096)
    got = None
    for rows, ptrs in op.handle_watermark_device(T0 + 2 * S):
        assert got is None
        got = [torch.as_tensor(_Ptr(c, rows), device=device).clone() for c in ptrs]
    st = op.stats()
    op.close()
    assert got is not None
    uk, inv = torch.unique(key, return_inverse=True)
    want_sum = torch.zeros(uk.numel(), dtype=torch.int64, device=device).index_add_(0, inv, val)
    want_cnt = torch.zeros(uk.numel(), dtype=torch.int64, device=device).index_add_(0, inv, torch.ones_like(val))
    k_out, ws, we, s_out, n_out = got[0], got[1], got[2], got[3], got[4]  # device layout: window = two i64 columns
    order = torch.argsort(k_out)
    assert torch.equal(k_out[order], uk)
    assert torch.equal(n_out[order], want_cnt)
    assert torch.equal(s_out[order], want_sum)
    assert bool((ws == T0).all()) and bool((we == T0 + S).all())
    assert st["rows_deferred"] <= 4096  # only the first small batch (no pane was resident yet)

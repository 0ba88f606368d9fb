"""Pin the oracle against the reference's own golden vectors.

The reference compares outputs as a sorted multiset of JSON lines
(crates/arroyo-sql-testing/src/smoke_tests.rs:619-692); so do we."""
import pytest

from oracle import arroyo_oracle as O
from tests.golden_cases import CASES, multiset


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_reference_golden(golden, name):
    inputs, expected = golden
    got = CASES[name](O, inputs)
    assert multiset(got) == multiset(expected[name]), name


@pytest.mark.parametrize("frac", [0.3, 0.6])
@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_reference_golden_across_a_checkpoint_and_restore(golden, name, frac):
    """The reference's smoke tests checkpoint every query mid-stream, stop it and resume it from the checkpoint;
    the output must still be the golden file (smoke_tests.rs).  Same here: pins handle_checkpoint / on_start of
    the restated operators, state tables "t" (tumbling / sliding), "s" + "e" (session) and "left" / "right" (join)."""
    from tests.restart_ops import OracleKit, RestartOps
    inputs, expected = golden
    got = CASES[name](RestartOps(O, OracleKit, frac), inputs)
    assert multiset(got) == multiset(expected[name]), name


@pytest.mark.parametrize("impl", ["numpy", "c"])
@pytest.mark.parametrize("case", sorted(__import__("tests.golden_cases", fromlist=["x"]).ACCUMULATOR_CASES))
def test_accumulators_match_reference_updating_aggregate_goldens(golden, accumulator_golden, case, impl):
    """SUM / AVG / MIN / MAX / COUNT over Int64 are pinned by the reference's grouped_aggregates / aggregates
    goldens (the updating aggregate reaches these final rows with the same DataFusion accumulators)."""
    from tests.golden_cases import ACCUMULATOR_CASES
    ops = O
    if impl == "c":
        from oracle import c_oracle
        ops = c_oracle
    got = ACCUMULATOR_CASES[case](ops, golden[0])
    assert multiset(got) == multiset(accumulator_golden[case[0]])


C_CASES = ["sliding_window_end", "hourly_by_event_type", "tight_watermark", "month_loose_watermark",
           "most_active_driver_last_hour", "offset_impulse_join", "nexmark_q5", "windowed_inner_join",
           "windowed_outer_join", "session_window", "global_session_window"]


@pytest.mark.parametrize("name", C_CASES)
def test_c_oracle_matches_reference_golden(golden, name):
    """The C restatement (timed CPU baseline) is pinned by the same vectors: its tumbling / sliding
    operators replace the numpy ones in every golden pipeline that contains a window aggregate."""
    from oracle import c_oracle
    inputs, expected = golden
    got = CASES[name](c_oracle, inputs)
    assert multiset(got) == multiset(expected[name]), name


def test_c_oracle_parallel_driver_matches_numpy_oracle():
    """oracle_run_windows (p key-partitioned subtasks) against the numpy operator: same checksums."""
    import numpy as np
    from oracle import c_oracle
    rng = np.random.default_rng(5)
    S = 1_000_000_000
    n = 120_000
    ts = 1_700_000_000 * S + np.arange(n, dtype=np.int64) * (S // 20_000)
    key = rng.integers(0, 3000, n, dtype=np.int64)
    val = rng.integers(0, 10**8, n, dtype=np.int64)
    cfg = O.WindowAggConfig(width=4 * S, slide=S, key_names=["key"],
                            aggs=[O.Agg("sum", "value", "sum"), O.Agg("avg", "value", "avg"),
                                  O.Agg("count", None, "n")], window_index=1)
    out = O.run_single_input(O.SlidingAggregatingWindowFunc(cfg),
                             O.source_batches({"key": key, "value": val, O.TIMESTAMP: ts}, 4096), S).all()
    for p in (1, 3):
        r = c_oracle.run_windows(key, val, ts, 4096, 4 * S, S, S, p)
        assert r.rows_out == out.num_rows
        assert r.sum_of_rows == int(out["n"].sum())
        assert r.sum_of_sums == int(out["sum"].astype(np.uint64).sum())
        assert abs(r.sum_of_avgs - float(out["avg"].sum())) <= 1e-9 * abs(float(out["avg"].sum()))


@pytest.mark.parametrize("kind", ["tumbling", "sliding"])
def test_windowed_sum_avg_min_max_against_an_independent_engine(kind):
    """No reference golden vector uses SUM or AVG inside a *windowed* aggregate; this pins the
    oracle's accumulators for them against an independent columnar engine: Arrow C++ / Acero group-by
    (hash_sum, hash_mean, hash_min, hash_max, hash_count) over (key, window) computed from first principles
    (every row belongs to the windows [b - width + slide .. b] of its bin b)."""
    import numpy as np
    import pyarrow as pa
    S = 1_000_000_000
    T0 = 1_700_000_000 * S
    rng = np.random.default_rng(12)
    n = 60_000
    ts = T0 + np.sort(rng.integers(0, 20 * S, n)).astype(np.int64)
    key = rng.integers(0, 500, n, dtype=np.int64) * 7919 - 3
    val = rng.integers(-10**8, 10**8, n, dtype=np.int64)
    width, slide = (2 * S, 0) if kind == "tumbling" else (5 * S, S)
    aggs = [O.Agg("sum", "value", "sum"), O.Agg("avg", "value", "avg"), O.Agg("min", "value", "mn"),
            O.Agg("max", "value", "mx"), O.Agg("count", None, "n")]
    cfg = O.WindowAggConfig(width=width, slide=slide, key_names=["key"], aggs=aggs, window_index=1)
    op = O.TumblingAggregatingWindowFunc(cfg) if kind == "tumbling" else O.SlidingAggregatingWindowFunc(cfg)
    # in-order stream, watermark far behind: nothing is late, every window is emitted by the final watermark
    got = O.run_single_input(op, O.source_batches({"key": key, "value": val, O.TIMESTAMP: ts}, 4096), 30 * S).all()

    step = width if kind == "tumbling" else slide
    bins = ts - ts % step
    ks, ws, vs = [], [], []
    for j in range(width // step):  # the window starting at bin - j * step contains the row
        ks.append(key)
        ws.append(bins - j * step)
        vs.append(val)
    t = pa.table({"key": np.concatenate(ks), "window_start": np.concatenate(ws), "value": np.concatenate(vs)})
    ref = t.group_by(["key", "window_start"]).aggregate([("value", "sum"), ("value", "mean"), ("value", "min"),
                                                         ("value", "max"), ("value", "count")])
    want = {(k, w): (s, m, lo, hi, c) for k, w, s, m, lo, hi, c in zip(
        ref["key"].to_pylist(), ref["window_start"].to_pylist(), ref["value_sum"].to_pylist(),
        ref["value_mean"].to_pylist(), ref["value_min"].to_pylist(), ref["value_max"].to_pylist(),
        ref["value_count"].to_pylist())}
    assert got.num_rows == len(want)
    for i in range(got.num_rows):
        s, m, lo, hi, c = want[(int(got["key"][i]), int(got["window_start"][i]))]
        assert int(got["window_end"][i]) == int(got["window_start"][i]) + width
        assert (int(got["sum"][i]), int(got["mn"][i]), int(got["mx"][i]), int(got["n"][i])) == (s, lo, hi, c)
        assert abs(float(got["avg"][i]) - m) <= 1e-9 * max(1.0, abs(m))


@pytest.mark.parametrize("join_type", ["inner", "left", "right", "full"])
def test_c_join_matches_numpy_oracle_on_random_streams(join_type):
    """join_oracle.c against the numpy restatement: duplicates on both sides, unmatched rows on both sides, several
    instants per watermark, rows that stay buffered across watermarks, and the "batch older than the watermark" panic."""
    import numpy as np
    from oracle import c_oracle
    from tests.golden_cases import multiset as ms
    S = 1_000_000_000
    T0 = 1_700_000_000 * S
    rng = np.random.default_rng(3)

    def drive(join):
        ctx, out = O.OperatorContext(2), O.Collector()
        t = T0
        for step in range(12):
            for _ in range(int(rng.integers(1, 4))):  # several instants per step
                t += int(rng.integers(1, 3)) * S
                n_l, n_r = int(rng.integers(1, 40)), int(rng.integers(1, 60))
                join.process_batch_index(0, 2, O.Batch({"id": rng.integers(0, 25, n_l), "a": rng.integers(0, 10**6, n_l),
                                                        O.TIMESTAMP: np.full(n_l, t, dtype=np.int64)}), ctx, out)
                join.process_batch_index(1, 2, O.Batch({"seller": rng.integers(10, 40, n_r), "id": rng.integers(0, 10**6, n_r),
                                                        "b": rng.integers(0, 9, n_r), O.TIMESTAMP: np.full(n_r, t, dtype=np.int64)}),
                                         ctx, out)
            wm = t - int(rng.integers(0, 3)) * S  # sometimes leaves the newest instants buffered
            for side in (0, 1):
                ctx.watermarks.set(side, wm)
            join.handle_watermark(wm, ctx, out)
        with pytest.raises(RuntimeError):  # instant_join.rs:129-139
            join.process_batch_index(0, 2, O.Batch({"id": np.array([1]), "a": np.array([1]),
                                                    O.TIMESTAMP: np.array([T0], dtype=np.int64)}), ctx, out)
        for side in (0, 1):
            ctx.watermarks.set(side, O.FINAL_WATERMARK)
        join.handle_watermark(O.FINAL_WATERMARK, ctx, out)
        return [r for b in out.batches for r in b.rows()]

    cfg = O.JoinConfig(left_on=["id"], right_on=["seller"], join_type=join_type)
    rng = np.random.default_rng(3)
    want = drive(O.InstantJoin(cfg))
    rng = np.random.default_rng(3)
    got = drive(c_oracle.InstantJoin(cfg))
    assert ms(got) == ms(want)
    assert len(want) > 100


@pytest.mark.parametrize("shape", ["bursts", "multi_row_runs", "disorder", "unkeyed"])
def test_c_session_matches_numpy_oracle_on_random_streams(shape):
    """session_oracle.c against the numpy restatement on streams that exercise the state machine's corners: bursts
    separated by more / exactly / less than the gap, several rows of a key inside one batch (the scan's off-by-one),
    bounded disorder with late rows, and the single global key."""
    import numpy as np
    from oracle import c_oracle
    from tests.golden_cases import multiset as ms
    S = 1_000_000_000
    T0 = 1_700_000_000 * S
    rng = np.random.default_rng({"bursts": 1, "multi_row_runs": 2, "disorder": 3, "unkeyed": 4}[shape])
    gap = 2 * S
    n_keys = 1 if shape == "unkeyed" else (12 if shape == "multi_row_runs" else 150)
    rows = []
    for k in range(n_keys):
        t = T0 + int(rng.integers(0, 3 * gap))
        for _ in range(25 if shape in ("multi_row_runs", "unkeyed") else 6):
            for _ in range(int(rng.integers(1, 7))):
                rows.append((t, k * 31 + 5, int(rng.integers(-1000, 1000))))
                t += int(rng.integers(1, gap)) if rng.random() > 0.05 else gap
            t += gap + int(rng.integers(1, 4 * gap))
    jitter = S if shape == "disorder" else 0
    rows.sort(key=lambda r: r[0] + (int(rng.integers(-jitter, jitter + 1)) if jitter else 0))
    a = np.array(rows, dtype=np.int64)
    batch = 400 if shape in ("multi_row_runs", "unkeyed") else 300
    cols = {"key": a[:, 1].copy(), "value": a[:, 2].copy(), O.TIMESTAMP: a[:, 0].copy()}
    keys = [] if shape == "unkeyed" else ["key"]
    if shape == "unkeyed":
        del cols["key"]
    batches = O.source_batches(cols, batch)
    cfg = O.SessionConfig(gap=gap, key_names=keys, window_index=len(keys),
                          aggs=[O.Agg("sum", "value", "sum"), O.Agg("avg", "value", "avg"), O.Agg("min", "value", "mn"),
                                O.Agg("max", "value", "mx"), O.Agg("count", None, "n")])
    delay = S // 2 if shape == "disorder" else S
    want = O.run_single_input(O.SessionAggregatingWindowFunc(cfg), batches, delay).all()
    got = O.run_single_input(c_oracle.SessionAggregatingWindowFunc(cfg), batches, delay).all()
    assert ms(got.rows()) == ms(want.rows())
    assert want.num_rows > 15

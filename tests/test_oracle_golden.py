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


C_CASES = ["sliding_window_end", "hourly_by_event_type", "tight_watermark", "month_loose_watermark",
           "most_active_driver_last_hour", "offset_impulse_join", "nexmark_q5", "windowed_inner_join",
           "windowed_outer_join"]


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

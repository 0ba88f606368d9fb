"""GPU parity against the reference's own golden vectors, through the C ABI."""
import pytest

from tests.golden_cases import CASES, multiset

pytestmark = pytest.mark.gpu

WINDOW_ONLY = ["sliding_window_end", "hourly_by_event_type", "tight_watermark", "month_loose_watermark",
               "most_active_driver_last_hour",
               # window aggregates + instant join, both on the GPU
               "windowed_inner_join", "windowed_outer_join", "offset_impulse_join", "nexmark_q5",
               # session windows (per-key state machines on the GPU)
               "session_window", "global_session_window"]


@pytest.fixture(scope="module")
def gpu_ops():
    from tests import gpu_ops as g
    return g


@pytest.mark.parametrize("name", WINDOW_ONLY)
def test_cuda_operators_match_reference_golden(golden, gpu_ops, name):
    inputs, expected = golden
    got = CASES[name](gpu_ops, inputs)
    assert multiset(got) == multiset(expected[name]), name


@pytest.mark.parametrize("name", ["sliding_window_end", "most_active_driver_last_hour"])
def test_remerge_mode_matches_reference_golden(golden, gpu_ops, name):
    """FLAG_REMERGE_ONLY = the reference's own algorithm (re-merge width/slide panes per slide)."""
    from arroyo_b200 import ffi
    inputs, expected = golden
    gpu_ops.DEFAULT_KW["flags"] = ffi.FLAG_REMERGE_ONLY
    try:
        got = CASES[name](gpu_ops, inputs)
    finally:
        gpu_ops.DEFAULT_KW.clear()
    assert multiset(got) == multiset(expected[name]), name


@pytest.mark.parametrize("case", sorted(__import__("tests.golden_cases", fromlist=["x"]).ACCUMULATOR_CASES))
def test_cuda_accumulators_match_reference_updating_aggregate_goldens(golden, accumulator_golden, gpu_ops, case):
    """SUM / AVG / MIN / MAX / COUNT on the GPU against the final rows of the reference's grouped_aggregates /
    aggregates goldens."""
    from tests.golden_cases import ACCUMULATOR_CASES
    got = ACCUMULATOR_CASES[case](gpu_ops, golden[0])
    assert multiset(got) == multiset(accumulator_golden[case[0]])


@pytest.mark.parametrize("frac", [0.3, 0.6])
@pytest.mark.parametrize("name", WINDOW_ONLY)
def test_cuda_operators_match_reference_golden_across_a_checkpoint_and_restore(golden, gpu_ops, name, frac):
    """As the reference's smoke tests run every query (smoke_tests.rs): checkpoint mid-stream, throw the operator away,
    restore a fresh one from the state tables ("t"; "s" + "e"; "left" / "right") and continue: still the golden."""
    from tests.restart_ops import GpuKit, RestartOps
    inputs, expected = golden
    got = CASES[name](RestartOps(gpu_ops, GpuKit, frac), inputs)
    assert multiset(got) == multiset(expected[name]), name

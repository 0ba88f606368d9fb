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

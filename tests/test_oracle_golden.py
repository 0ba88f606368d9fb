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

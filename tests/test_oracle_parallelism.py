"""The reference's smoke tests re-run every golden query at parallelism 2 and 3 and expect the same output
(crates/arroyo-sql-testing/src/smoke_tests.rs: the pipeline is rescaled 1 -> 2 -> 3 between checkpoints).  The same
here for the oracle: the keyed golden pipelines behind a p x p shuffle topology -- p source subtasks with their own
watermark generators, key-hash repartition, p window subtasks that min-merge the senders' watermarks -- must still
reproduce the reference's golden vectors."""
import pytest

from oracle import arroyo_oracle as O
from tests.golden_cases import CASES, multiset

S = 1_000_000_000


class _Spec:
    def __init__(self, cls, cfg):
        self.cls, self.cfg = cls, cfg


class _All:
    def __init__(self, batches):
        self.batches = batches

    def all(self):
        return O.Batch.concat(self.batches) if self.batches else None


class ParallelOps:
    """Factory namespace for tests.golden_cases: every window operator becomes p subtasks behind a shuffle."""

    def __init__(self, p):
        self.p = p

    def TumblingAggregatingWindowFunc(self, cfg):
        return _Spec(O.TumblingAggregatingWindowFunc, cfg)

    def SlidingAggregatingWindowFunc(self, cfg):
        return _Spec(O.SlidingAggregatingWindowFunc, cfg)

    def SessionAggregatingWindowFunc(self, cfg):
        return _Spec(O.SessionAggregatingWindowFunc, cfg)

    def run_single_input(self, spec, batches, delay_ns=S):
        p = self.p
        keys = list(spec.cfg.key_names)
        assert keys, "unkeyed aggregates are planned with parallelism 1"
        ops = [spec.cls(spec.cfg) for _ in range(p)]
        ctxs = [O.OperatorContext(p) for _ in range(p)]
        outs = [O.Collector() for _ in range(p)]
        shards = [batches[s::p] for s in range(p)]  # source subtask s reads every p-th batch
        gens = [O.WatermarkGenerator(delay_ns) for _ in range(p)]
        n_rounds = max(len(s) for s in shards)
        for i in range(n_rounds + 1):
            wms = []
            for s in range(p):
                if i < len(shards[s]):
                    b = shards[s][i]
                    for d, sb in O.repartition(b, keys, p):
                        ops[d].process_batch(sb, ctxs[d], outs[d])
                    wms.append(gens[s].process_batch(b[O.TIMESTAMP]))
                else:
                    wms.append(O.FINAL_WATERMARK if i >= len(shards[s]) else None)
            for d in range(p):
                before = ctxs[d].last_present_watermark()
                for s in range(p):
                    if wms[s] is not None:
                        ctxs[d].watermarks.set(s, wms[s])
                after = ctxs[d].last_present_watermark()
                if after is not None and after != before:
                    ops[d].handle_watermark(after, ctxs[d], outs[d])
        return _All([b for o in outs for b in o.batches])


@pytest.mark.parametrize("p", [2, 3])
@pytest.mark.parametrize("name", ["hourly_by_event_type", "most_active_driver_last_hour", "session_window"])
def test_oracle_matches_golden_at_parallelism(golden, name, p):
    inputs, expected = golden
    got = CASES[name](ParallelOps(p), inputs)
    assert multiset(got) == multiset(expected[name]), (name, p)

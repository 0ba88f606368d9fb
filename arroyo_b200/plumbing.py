"""The stateless CPU plumbing around the hot path (BASELINE configs[0]: "Nexmark q1 stateless map/filter, 1 CPU worker,
64 Ki-row Arrow batches (plumbing, no GPU)") and the run loop's barrier alignment.  Nothing here touches the GPU: these
operators evaluate expressions over Arrow batches on the host with Arrow C++ kernels (pyarrow.compute), the way the
reference evaluates DataFusion physical expressions:

  ProjectionOperator      arroyo-worker/src/arrow/mod.rs:133-177   one expression per output column
  ValueExecutionOperator  arroyo-worker/src/arrow/mod.rs:53-96      a stateless plan (here: filter + projection)
  KeyExecutionOperator    arroyo-worker/src/arrow/mod.rs:179-243    key projection: routing-key copies prepended
                                                                    (arroyo-planner/src/plan/aggregate.rs:228-260)
  CheckpointCounter       arroyo-operator/src/lib.rs:71-113         barrier alignment over an operator's inputs
"""
from typing import Callable, List, Optional, Sequence

import pyarrow as pa
import pyarrow.compute as pc

from .context import Collector, OperatorContext

Expr = Callable[[pa.RecordBatch], pa.Array]


def col(name: str) -> Expr:
    return lambda b: b.column(b.schema.get_field_index(name))


class ProjectionOperator:
    """`exprs[i](batch)` produces output column i (mod.rs:156-177)."""

    def __init__(self, name: str, exprs: Sequence[Expr], output_names: Sequence[str]):
        self._name, self.exprs, self.output_names = name, list(exprs), list(output_names)

    def name(self):
        return self._name

    def process_batch(self, batch: pa.RecordBatch, ctx: Optional[OperatorContext], collector: Collector):
        collector.collect(pa.RecordBatch.from_arrays([e(batch) for e in self.exprs], names=self.output_names))


class ValueExecutionOperator:
    """A stateless plan over each batch (mod.rs:84-96): an optional filter predicate, then a projection.  Batches the
    filter empties are still forwarded with zero rows, like DataFusion's FilterExec stream yields them."""

    def __init__(self, name: str, exprs: Sequence[Expr], output_names: Sequence[str], predicate: Optional[Expr] = None):
        self._name, self.predicate = name, predicate
        self.project = ProjectionOperator(name, exprs, output_names)

    def name(self):
        return self._name

    def process_batch(self, batch: pa.RecordBatch, ctx: Optional[OperatorContext], collector: Collector):
        if self.predicate is not None:
            batch = batch.filter(self.predicate(batch))
        self.project.process_batch(batch, ctx, collector)


class KeyExecutionOperator:
    """Key projection in front of a Shuffle edge: `[_key_<g>..., <all input columns>]`, the key columns being copies
    used for routing only (mod.rs:179-243; plan/aggregate.rs:228-260)."""

    def __init__(self, name: str, key_exprs: Sequence[Expr], key_names: Sequence[str]):
        self._name, self.key_exprs, self.key_names = name, list(key_exprs), list(key_names)

    def name(self):
        return self._name

    def process_batch(self, batch: pa.RecordBatch, ctx: Optional[OperatorContext], collector: Collector):
        arrays = [e(batch) for e in self.key_exprs] + list(batch.columns)
        collector.collect(pa.RecordBatch.from_arrays(arrays, names=self.key_names + batch.schema.names))


def nexmark_q1() -> ValueExecutionOperator:
    """Nexmark q1 (currency conversion): SELECT auction, bidder, 0.908 * price AS price, datetime FROM bid."""
    return ValueExecutionOperator(
        "q1", [col("auction"), col("bidder"), lambda b: pc.multiply(pc.cast(col("price")(b), pa.float64()), 0.908),
               col("_timestamp")], ["auction", "bidder", "price", "_timestamp"])


def nexmark_q2(modulus: int = 123) -> ValueExecutionOperator:
    """Nexmark q2 (selection): SELECT auction, price FROM bid WHERE auction % 123 = 0."""
    def pred(b):
        a = col("auction")(b)
        return pc.equal(pc.subtract(a, pc.multiply(pc.divide(a, modulus), modulus)), 0)
    return ValueExecutionOperator("q2", [col("auction"), col("price"), col("_timestamp")],
                                  ["auction", "price", "_timestamp"], predicate=pred)


class CheckpointCounter:
    """Barrier alignment (lib.rs:71-113): an input that delivered the barrier is blocked until every input has;
    `mark` returns True when the barrier is complete (and unblocks everything)."""

    def __init__(self, size: int):
        self.inputs: List[Optional[int]] = [None] * size
        self.counter: Optional[int] = None

    def is_blocked(self, idx: int) -> bool:
        return self.inputs[idx] is not None

    def all_clear(self) -> bool:
        return all(x is None for x in self.inputs)

    def mark(self, idx: int, epoch: int) -> bool:
        assert self.inputs[idx] is None
        if len(self.inputs) == 1:
            return True
        self.inputs[idx] = epoch
        if self.counter is None:
            self.counter = len(self.inputs) - 1
        elif self.counter == 1:
            self.inputs = [None] * len(self.inputs)
            self.counter = None
        else:
            self.counter -= 1
        return self.counter is None

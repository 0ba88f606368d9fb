"""Operator configurations: what the reference's protobuf operator configs carry
(arroyo-rpc/proto/api.proto:39-80), reduced to the supported plan subset.  Durations are int ns."""
from dataclasses import dataclass, field
from typing import List, Optional


@dataclass
class Agg:
    kind: str            # count | sum | avg | min | max
    col: Optional[str]   # input column (None for count(*))
    name: str            # output column name


@dataclass
class WindowAggConfig:
    """TumblingWindowAggregateOperator / SlidingWindowAggregateOperator.
    Input  [key cols..., value cols..., _timestamp]; output = aggregate output [keys, aggs] with the
    window struct {start, end} inserted at `window_index`, then `_timestamp = bin + width - 1 ns`
    (arroyo-planner/src/extension/aggregate.rs:292-390).  `final_projection=False` (tumbling only)
    gives [keys, aggs, _timestamp = bin]."""
    width: int
    slide: int = 0
    key_names: List[str] = field(default_factory=list)
    aggs: List[Agg] = field(default_factory=list)
    final_projection: bool = True
    window_index: int = 0
    # final stage of a partial -> shuffle -> final plan: name of the input column that carries how many
    # original rows each (partial-aggregate) input row stands for; None = inputs are raw rows
    partial_count_col: Optional[str] = None


@dataclass
class SessionConfig:
    """SessionWindowAggregateOperator (planner extension/aggregate.rs:170-231)."""
    gap: int
    key_names: List[str] = field(default_factory=list)
    aggs: List[Agg] = field(default_factory=list)
    window_index: int = 0


@dataclass
class JoinConfig:
    """JoinOperator for the instant (windowed) join (arroyo-worker/src/arrow/instant_join.rs)."""
    left_on: List[str]
    right_on: List[str]
    join_type: str = "inner"
    left_routing_keys: List[str] = field(default_factory=list)
    right_routing_keys: List[str] = field(default_factory=list)

"""Host logic: the C++ window state machines (arroyo_b200/csrc/planner.h) against the oracle's
restatement of the reference operators, on random event sequences.  No GPU involved."""
import ctypes as C

import numpy as np
import pytest

from arroyo_b200 import ffi
from oracle import arroyo_oracle as O

S = 1_000_000_000
NO_TIME = -(1 << 63)


def plan(kind, width, slide, events):
    lib = ffi.load()
    ev = np.array(events, dtype=np.int64).reshape(-1)
    out = np.zeros(4 * (len(events) * 64 + 1024), dtype=np.int64)
    evp = ev.ctypes.data_as(C.POINTER(C.c_int64))
    outp = out.ctypes.data_as(C.POINTER(C.c_int64))
    if kind == "sliding":
        n = lib.arroyo_b200_plan_sliding(width, slide, evp, len(events), outp, len(out))
    else:
        n = lib.arroyo_b200_plan_tumbling(width, evp, len(events), outp, len(out))
    assert n >= 0
    return out[:4 * n].reshape(n, 4).tolist()


def windows_from_plan(steps, rows_per_bin_at_close):
    """Interpret the plan the way the device executor does: JOIN/LEAVE maintain the window store;
    EMIT with no member pane emits nothing."""
    store = set()
    res = []
    for kind, a, b, c in steps:
        if kind == 2:
            store.add(a)
        elif kind == 3:
            store.discard(a)
        elif kind == 1:
            members = [m for m in store if a <= m < b] if rows_per_bin_at_close["sliding"] else [c]
            total = sum(rows_per_bin_at_close["rows"].get(m, 0) for m in members)
            if total:
                res.append((a, b, total))
    return res


def random_events(rng, slide, n_steps, gap_prob, checkpoint_prob):
    """A sequence of (touch bin-with-n-rows | watermark | checkpoint) events with increasing-ish time."""
    t = 1_696_871_600 * S
    t -= t % slide
    events = []
    wm = None
    for _ in range(n_steps):
        r = rng.random()
        if r < 0.55:
            # data for a bin near `t` (some disorder, some late)
            b = t + int(rng.integers(-3, 4)) * slide
            events.append(("data", b, int(rng.integers(1, 5))))
        elif r < 0.55 + checkpoint_prob:
            events.append(("checkpoint",))
        else:
            cand = t - int(rng.integers(0, 3)) * slide + int(rng.integers(0, slide))
            if wm is None or cand > wm:
                wm = cand
                events.append(("watermark", wm))
        if rng.random() < gap_prob:
            t += int(rng.integers(5, 40)) * slide  # long silence: the window store runs empty
        else:
            t += int(rng.integers(0, 2)) * slide
    events.append(("watermark", O.FINAL_WATERMARK))
    return events


@pytest.mark.parametrize("seed", range(12))
@pytest.mark.parametrize("width_panes", [1, 3, 5])
def test_sliding_planner_matches_oracle(seed, width_panes):
    rng = np.random.default_rng(seed)
    slide = 2 * S
    width = width_panes * slide
    if width_panes == 1:
        width = slide * 2  # slide == width degrades to tumbling in the planner; keep hop shape
    evs = random_events(rng, slide, 120, gap_prob=0.08, checkpoint_prob=0.06 if seed % 2 else 0.0)

    # oracle run (unkeyed count(*)), one row-batch per data event
    cfg = O.WindowAggConfig(width=width, slide=slide, key_names=[], aggs=[O.Agg("count", None, "count")])
    op = O.SlidingAggregatingWindowFunc(cfg)
    ctx, out = O.OperatorContext(1), O.Collector()
    plan_events = []
    ontime_rows = {}
    for e in evs:
        if e[0] == "data":
            _, b, n = e
            ts = np.full(n, b + 1, dtype=np.int64)
            op.process_batch(O.Batch({O.TIMESTAMP: ts}), ctx, out)
            wm = ctx.last_present_watermark()
            if wm is None or b >= O.bin_start(wm, slide):
                ontime_rows[b] = ontime_rows.get(b, 0) + n
            plan_events.append((0, b))
        elif e[0] == "watermark":
            ctx.watermarks.set(0, e[1])
            op.handle_watermark(e[1], ctx, out)
            plan_events.append((1, min(e[1], (1 << 63) - 1)))
        else:
            op.handle_checkpoint(ctx)
            plan_events.append((2, 0))
    want = [(int(b["window_start"][0]), int(b["window_end"][0]), int(b["count"][0])) for b in out.batches]

    steps = plan("sliding", width, slide, plan_events)
    got = windows_from_plan(steps, {"sliding": True, "rows": ontime_rows})
    assert got == want


@pytest.mark.parametrize("seed", range(6))
def test_tumbling_planner_matches_oracle(seed):
    rng = np.random.default_rng(100 + seed)
    width = 5 * S
    evs = random_events(rng, width, 100, gap_prob=0.05, checkpoint_prob=0.05)
    cfg = O.WindowAggConfig(width=width, key_names=[], aggs=[O.Agg("count", None, "count")])
    op = O.TumblingAggregatingWindowFunc(cfg)
    ctx, out = O.OperatorContext(1), O.Collector()
    plan_events, ontime_rows = [], {}
    for e in evs:
        if e[0] == "data":
            _, b, n = e
            op.process_batch(O.Batch({O.TIMESTAMP: np.full(n, b + 7, dtype=np.int64)}), ctx, out)
            wm = ctx.last_present_watermark()
            if wm is None or b >= O.bin_start(wm, width):
                ontime_rows[b] = ontime_rows.get(b, 0) + n
            plan_events.append((0, b))
        elif e[0] == "watermark":
            ctx.watermarks.set(0, e[1])
            op.handle_watermark(e[1], ctx, out)
            plan_events.append((1, min(e[1], (1 << 63) - 1)))
        else:
            op.handle_checkpoint(ctx)
            plan_events.append((2, 0))
    want = [(int(b["window_start"][0]), int(b["window_end"][0]), int(b["count"][0])) for b in out.batches]
    steps = plan("tumbling", width, 0, plan_events)
    got = windows_from_plan(steps, {"sliding": False, "rows": ontime_rows})
    assert got == want

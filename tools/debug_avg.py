import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import arroyo_oracle as O
from tests import gpu_ops as G
from tests.test_gpu_parity import gen_stream, SUM_AVG, S, rows_of
from arroyo_b200 import ffi

rng = np.random.default_rng(33)
batches = gen_stream(rng, 120_000, 700, rate_per_s=10_000, batch=5000)
for b in batches[len(batches) // 2:]:
    v = b["value"].copy()
    v[::211] = (1 << 40) + 12345
    v[5::499] = -(1 << 61)
    b.cols["value"] = v
cfg = O.WindowAggConfig(width=5 * S, slide=S, key_names=["key"], aggs=SUM_AVG, window_index=1)
want = rows_of(O.run_single_input(O.SlidingAggregatingWindowFunc(cfg), batches, S).batches)
wd = {(r["key"], r["window_start"]): r for r in want}
for name, flags in [("default", 0), ("nocombine", ffi.FLAG_NO_COMBINE), ("avgf64", ffi.FLAG_AVG_F64),
                    ("avgf64+nocombine", ffi.FLAG_AVG_F64 | ffi.FLAG_NO_COMBINE), ("remerge", ffi.FLAG_REMERGE_ONLY)]:
    gop = G.SlidingAggregatingWindowFunc(cfg, flags=flags)
    got = rows_of(G.run_single_input(gop, batches, S).batches)
    bad = 0
    shown = 0
    for r in got:
        w = wd[(r["key"], r["window_start"])]
        ok = r["sum"] == w["sum"] and r["count"] == w["count"] and abs(r["avg"] - w["avg"]) <= 1e-6 * abs(w["avg"])
        if not ok:
            bad += 1
            if shown < 4:
                shown += 1
                print("   ", name, r, "WANT", w)
    print(name, "rows", len(got), "bad", bad, gop.stats()["rows_deferred"])

"""Runs one session parity case (argv[1]) and reports the first differences."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import arroyo_oracle as O
from tests import gpu_ops as G
from tests.test_gpu_parity import S, SESSION_AGGS, rows_of, session_stream

case = sys.argv[1]
limit = int(sys.argv[2]) if len(sys.argv) > 2 else 10**9
seeds = {"in_order": 1, "small_batches": 2, "multi_row_runs": 3, "disorder": 4, "no_watermark_until_end": 5}
rng = np.random.default_rng(seeds[case])
gap = 5 * S
kw = dict(n_keys=300, n_bursts=6, gap=gap, batch=500)
delay = S
if case == "small_batches":
    kw.update(batch=37)
if case == "multi_row_runs":
    kw.update(n_keys=12, n_bursts=25, batch=400)
if case == "disorder":
    kw.update(jitter=2 * S)
    delay = 3 * S
batches = session_stream(rng, **kw)[:limit]
cfg = O.SessionConfig(gap=gap, key_names=["key"], aggs=SESSION_AGGS, window_index=0)
want = rows_of(O.run_single_input(O.SessionAggregatingWindowFunc(cfg), batches, delay).batches)
print(case, "batches", len(batches), "oracle rows", len(want), flush=True)
got = rows_of(G.run_single_input(G.SessionAggregatingWindowFunc(cfg), batches, delay).batches)
print("gpu rows", len(got), flush=True)
key = lambda r: (r["key"], r["window_start"])
wd, gd = {key(r): r for r in want}, {key(r): r for r in got}
bad = 0
for k in sorted(set(wd) | set(gd)):
    w, g = wd.get(k), gd.get(k)
    same = w is not None and g is not None and all(w[c] == g[c] for c in w if c != "avg") and abs(w["avg"] - g["avg"]) <= 1e-6 * abs(w["avg"]) + 1e-9
    if not same:
        bad += 1
        if bad <= 6:
            print("DIFF", k, "\n  want", w, "\n  got ", g)
print("bad", bad)

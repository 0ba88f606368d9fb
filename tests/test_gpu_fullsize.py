"""The benchmarked configuration itself, checked: BASELINE configs[2] at full width -- 2^20 scattered i64 keys,
2^24 rows per 1-s pane in 256 batches of 65 536 rows, hop(1 s, 10 s) SUM / AVG / COUNT -- 14 panes through the CUDA
operator (device-resident entry points, exactly as bench.py drives them) against

  * oracle/window_oracle.c on the same panes, window by window: rows out, sum of COUNT(*), wrapping sum of SUM(value)
    bit-exact, sum of AVG(value) to 1e-6 relative (the reference's own tests compare sorted outputs:
    arroyo-sql-testing/src/smoke_tests.rs:619-692; at 10^6 rows per window a checksum per column stands in for that);
  * a full row-by-row comparison of one window against an independent group-by (torch.unique + index_add on the
    window's raw rows): keys, counts and sums bit-exact, averages to 1e-6 relative.

Uniform keys and the Nexmark hot-key skew (75 % of the rows on one key)."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def _args(dist):
    import bench as B
    argv = sys.argv
    sys.argv = ["bench.py", "--dist", dist]
    try:
        return B.parse()
    finally:
        sys.argv = argv


@pytest.mark.parametrize("dist", ["uniform", "hot"])
def test_headline_config_full_size_matches_the_c_oracle(dist):
    import torch

    import bench as B
    from arroyo_b200 import ffi, operators as native
    from arroyo_b200.multi_gpu import _Ptr

    args = _args(dist)
    device = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    torch.cuda.set_stream(torch.cuda.Stream(device=device))
    rows, n_panes = args.rows_per_pane, 14
    assert rows == 1 << 24 and args.keys == 1 << 20

    # the checker first: it decides whether this host can afford full panes (it can on the GPU boxes)
    cpu = B.run_cpu(torch, args, device, budget_s=240.0, warm_panes=n_panes - 2, timed_panes=2)
    assert cpu["rows_per_step"] == rows, "host too slow for full-size panes: the test would not check the benchmarked size"
    gen = B.make_generator(torch, device, rows, args.keys, args.dist, 42, args.keyspace)
    panes = [gen(p) for p in range(n_panes)]

    _, d, _, _, sums = B.device_resident(args, torch, native, ffi, 0, panes, n_panes, 0, rows, collect=True)
    res = B.compare_windows(sums, cpu["windows"], min_windows=n_panes - 4)
    assert res["verified"], res
    # every window of the steady state holds every key
    full = [ws for ws, v in sums.items() if v[2] == 10 * rows]
    assert full and all(sums[ws][1] == args.keys for ws in full)

    # one window row by row against an independent group-by over its raw rows
    import pyarrow as pa
    schema = pa.schema([("key", pa.int64()), ("value", pa.int64()), ("_timestamp", pa.timestamp("ns"))])
    op = native.SlidingAggregatingWindowFunc(B.window_config(), input_schema=schema, device=0,
                                             stream=torch.cuda.current_stream().cuda_stream, flags=B.op_flags(args),
                                             expected_keys=args.keys)
    plans = B.build_batch_lists(torch, panes, rows)
    wend = B.T0 + 3 * B.S  # the window [T0 - 7 s, T0 + 3 s): panes 0, 1, 2
    wstart = wend - B.WIDTH
    got = None
    for p in range(6):
        for cols, nrows, wm in plans[p]:
            op.process_device_batches(cols, nrows, 3)
            if wm is not None:
                for n, ptrs in op.handle_watermark_device(wm):
                    ws = int(torch.as_tensor(_Ptr(ptrs[1], n), device=device)[0].item())
                    if ws == wstart and got is None:
                        got = [torch.as_tensor(_Ptr(c, n), device=device).clone() for c in ptrs]
    op.close()
    assert got is not None, "the sampled window was not emitted"
    key = torch.cat([panes[p][0] for p in range(3)])
    val = torch.cat([panes[p][1] for p in range(3)])
    uk, inv = torch.unique(key, return_inverse=True)
    want_sum = torch.zeros(uk.numel(), dtype=torch.int64, device=device).index_add_(0, inv, val)
    want_cnt = torch.zeros(uk.numel(), dtype=torch.int64, device=device).index_add_(0, inv, torch.ones_like(val))
    order = torch.argsort(got[0])
    assert torch.equal(got[0][order], uk)
    assert torch.equal(got[3][order], want_sum)
    assert torch.equal(got[5][order], want_cnt)
    avg = got[4][order].view(torch.float64)
    want_avg = want_sum.to(torch.float64) / want_cnt.to(torch.float64)
    assert torch.allclose(avg, want_avg, rtol=1e-6, atol=0)
    assert bool((got[1] == wstart).all()) and bool((got[2] == wend).all())
    assert bool((got[6] == wend - 1).all())

"""bench.py's N > 1 verification sums the owners' per-window checksums with all-reduces (multi_gpu._gather_window_sums):
world 2 over gloo on CPU.  Keys are disjoint across owners, so a window's global checksum is the sum of the owners';
the SUM checksum must stay a wrapping 64-bit sum through the int64 all-reduce, and ranks that emitted different windows
must be reported, not summed."""
import json
import os
import sys
import tempfile

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

WORLD = 2
S = 1_000_000_000
M64 = (1 << 64) - 1


def sums_of(rank, mismatch):
    # ws -> (we, rows_out, sum COUNT, wrapping sum SUM (unsigned), sum AVG)
    out = {}
    for w in range(5):
        ws = 1_700_000_000 * S + w * S
        out[ws] = (ws + 10 * S, 1000 + rank, 16_000 + 7 * rank + w, (0xF000000000000000 + 12345 * (rank + 1) + w) & M64,
                   1.5 * (rank + 1) + w)
    if mismatch and rank == 1:
        out.pop(max(out))
        out[1] = (2, 1, 1, 1, 1.0)
    return out


def worker(rank, port, outdir, mismatch):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(WORLD))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    from arroyo_b200.multi_gpu import _gather_window_sums
    merged, err = _gather_window_sums(torch, dist, torch.device("cpu"), sums_of(rank, mismatch))
    with open(os.path.join(outdir, f"rank{rank}.json"), "w") as f:
        json.dump({"err": err, "merged": None if merged is None else {str(k): list(v) for k, v in merged.items()}}, f)
    dist.barrier()
    dist.destroy_process_group()


def _run(mismatch):
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(worker, args=(36533 + os.getpid() % 1000 + (500 if mismatch else 0), d, mismatch), nprocs=WORLD, join=True)
        return [json.load(open(os.path.join(d, f"rank{r}.json"))) for r in range(WORLD)]


def test_owner_checksums_are_summed_with_a_wrapping_sum_checksum():
    got = _run(False)
    a, b = sums_of(0, False), sums_of(1, False)
    for res in got:
        assert res["err"] is None
        for ws, (we, n, cnt, sm, av) in a.items():
            m = res["merged"][str(ws)]
            assert m[0] == we and m[1] == n + b[ws][1] and m[2] == cnt + b[ws][2]
            assert m[3] == (sm + b[ws][3]) & M64  # 0xF... + 0xF... wraps past 2^64
            assert abs(m[4] - (av + b[ws][4])) < 1e-9


def test_ranks_that_emitted_different_windows_are_reported():
    for res in _run(True):
        assert res["merged"] is None and "different windows" in res["err"]

"""Session windows and the instant join behind the key-hash shuffle (BASELINE configs[3] / configs[4] plan shapes), world
size 2 over gloo on CPU: ShuffleExchange.round_packed carries the raw rows of each edge, every sender's block reaches the
owner's operator as its own batch (session results depend on what shares a batch), the owner min-merges the senders'
watermarks per edge and, for the join, across its two inputs.  Expected = a single-process simulation of the same
2 x 2 topology with the oracle's operators."""
import ctypes
import json
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import arroyo_oracle as O  # noqa: E402
from tests.golden_cases import multiset  # noqa: E402

S = 1_000_000_000
T0 = 1_700_000_000 * S
WORLD = 2


def np_pack(batch, names, key):
    counts = np.zeros(WORLD, dtype=np.int64)
    parts = {d: sb for d, sb in O.repartition(batch, [key], WORLD)}
    blocks = []
    for d in range(WORLD):
        if d in parts:
            counts[d] = parts[d].num_rows
            blocks += [np.ascontiguousarray(parts[d][c]).astype(np.int64) for c in names]
    flat = np.concatenate(blocks) if blocks else np.empty(0, dtype=np.int64)
    return torch.from_numpy(flat), torch.from_numpy(counts)


def read_ptr(ptr, n):
    return np.ctypeslib.as_array((ctypes.c_int64 * n).from_address(ptr)).copy() if n else np.empty(0, dtype=np.int64)


# ---------------------------------------------------------------------------------------------------------------
# session windows
# ---------------------------------------------------------------------------------------------------------------
SESS_NAMES = ("key", "value", O.TIMESTAMP)


def session_shard(rank):
    rng = np.random.default_rng(300 + rank)
    rows = []
    for k in range(60):
        t = T0 + int(rng.integers(0, 4 * S))
        for _ in range(8):
            for _ in range(int(rng.integers(1, 5))):
                rows.append((t, 1000 + 7 * k + rank, int(rng.integers(-50, 50))))
                t += int(rng.integers(1, 2 * S))
            t += 2 * S + int(rng.integers(1, 6 * S))
    rows.sort()
    a = np.array(rows, dtype=np.int64)
    # both ranks hold rows of every key parity: keys 1000 + 7k + rank hash to either owner
    return O.source_batches({"key": a[:, 1].copy(), "value": a[:, 2].copy(), O.TIMESTAMP: a[:, 0].copy()}, 37)


def session_cfg():
    return O.SessionConfig(gap=2 * S, key_names=["key"], aggs=[O.Agg("sum", "value", "sum"), O.Agg("count", None, "n")],
                           window_index=1)


def session_expected():
    shards = [session_shard(r) for r in range(WORLD)]
    gens = [O.WatermarkGenerator() for _ in range(WORLD)]
    ops = [O.SessionAggregatingWindowFunc(session_cfg()) for _ in range(WORLD)]
    ctxs = [O.OperatorContext(WORLD) for _ in range(WORLD)]
    outs = [O.Collector() for _ in range(WORLD)]
    n_rounds = max(len(s) for s in shards)
    for i in range(n_rounds + 1):
        wms = []
        for s in range(WORLD):
            if i < len(shards[s]):
                b = shards[s][i]
                for d, sb in O.repartition(b, ["key"], WORLD):
                    ops[d].process_batch(sb, ctxs[d], outs[d])
                wms.append(gens[s].process_batch(b[O.TIMESTAMP]))
            else:
                wms.append(O.FINAL_WATERMARK if i == n_rounds else None)
        for d in range(WORLD):
            before = ctxs[d].last_present_watermark()
            for s in range(WORLD):
                if wms[s] is not None:
                    ctxs[d].watermarks.set(s, wms[s])
            after = ctxs[d].last_present_watermark()
            if after is not None and after != before:
                ops[d].handle_watermark(after, ctxs[d], outs[d])
    return [r for o in outs for b in o.batches for r in b.rows()]


def session_worker(rank, port, outdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(WORLD))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    from arroyo_b200.multi_gpu import ShuffleExchange
    ex = ShuffleExchange(torch, dist, rank, WORLD, None, torch.device("cpu"), max_recv_rows=1 << 14, n_cols=3)
    batches = session_shard(rank)
    gen = O.WatermarkGenerator()
    op = O.SessionAggregatingWindowFunc(session_cfg())
    ctx, out = O.OperatorContext(1), O.Collector()
    n_rounds = torch.tensor([len(batches)])
    dist.all_reduce(n_rounds, op=dist.ReduceOp.MAX)
    n_rounds = int(n_rounds)
    for i in range(n_rounds + 1):
        if i < len(batches):
            b = batches[i]
            packed, counts = np_pack(b, SESS_NAMES, "key")
            m, wm = b.num_rows, gen.process_batch(b[O.TIMESTAMP])
        else:
            packed, counts, m = None, None, 0
            wm = O.FINAL_WATERMARK if i == n_rounds else None
        got, eff, _ = ex.round_packed(packed, counts, m, wm)
        for cols, r in got:  # one batch per sender, in sender order
            op.process_batch(O.Batch({c: read_ptr(p, r) for c, p in zip(SESS_NAMES, cols)}), ctx, out)
        if eff is not None:
            ctx.watermarks.set(0, eff)
            op.handle_watermark(eff, ctx, out)
    rows = [r for b in out.batches for r in b.rows()]
    with open(os.path.join(outdir, f"rank{rank}.json"), "w") as f:
        json.dump(rows, f)
    dist.barrier()
    dist.destroy_process_group()


def test_session_windows_behind_the_shuffle_world2_gloo():
    want = session_expected()
    assert len(want) > 300
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(session_worker, args=(32533 + os.getpid() % 1000, d), nprocs=WORLD, join=True)
        got = [r for k in range(WORLD) for r in json.load(open(os.path.join(d, f"rank{k}.json")))]
    assert multiset(got) == multiset(want)


# ---------------------------------------------------------------------------------------------------------------
# instant join: two shuffled inputs
# ---------------------------------------------------------------------------------------------------------------
L_NAMES = ("id", "name_code", O.TIMESTAMP)
R_NAMES = ("seller", "auction", "reserve", O.TIMESTAMP)
N_WINDOWS = 7
W30 = 30 * S


def join_shard(rank):
    """Per window w: this rank's share of the persons and auctions of that window, stamped with the window's
    _timestamp (what the upstream window operators emit), and the watermark that follows them."""
    rng = np.random.default_rng(500 + rank)
    lefts, rights, wms = [], [], []
    for w in range(N_WINDOWS):
        ts = T0 + (w + 1) * W30 - 1
        n_p, n_a = 40 + 5 * w + rank, 120 + 11 * w + 3 * rank
        lefts.append(O.Batch({"id": rng.integers(0, 80, n_p, dtype=np.int64), "name_code": rng.integers(0, 10**6, n_p, dtype=np.int64),
                              O.TIMESTAMP: np.full(n_p, ts, dtype=np.int64)}))
        rights.append(O.Batch({"seller": rng.integers(40, 120, n_a, dtype=np.int64),
                               "auction": np.arange(n_a, dtype=np.int64) + 1000 * w + 100_000 * rank,
                               "reserve": rng.integers(1, 10**5, n_a, dtype=np.int64),
                               O.TIMESTAMP: np.full(n_a, ts, dtype=np.int64)}))
        wms.append(ts + 1 if w % 2 == 1 or rank == 0 else None)  # rank 1 reports only every other watermark
    return lefts, rights, wms


def join_cfg(join_type):
    return O.JoinConfig(left_on=["id"], right_on=["seller"], join_type=join_type)


def join_expected(join_type):
    shards = [join_shard(r) for r in range(WORLD)]
    joins = [O.InstantJoin(join_cfg(join_type)) for _ in range(WORLD)]
    ctxs = [O.OperatorContext(2 * WORLD) for _ in range(WORLD)]  # inputs: left senders, then right senders
    outs = [O.Collector() for _ in range(WORLD)]
    for w in range(N_WINDOWS + 1):
        for s in range(WORLD):
            if w < N_WINDOWS:
                for d, sb in O.repartition(shards[s][0][w], ["id"], WORLD):
                    joins[d].process_batch_index(s, 2 * WORLD, sb, ctxs[d], outs[d])
                for d, sb in O.repartition(shards[s][1][w], ["seller"], WORLD):
                    joins[d].process_batch_index(WORLD + s, 2 * WORLD, sb, ctxs[d], outs[d])
        for d in range(WORLD):
            before = ctxs[d].last_present_watermark()
            for s in range(WORLD):
                wm = shards[s][2][w] if w < N_WINDOWS else O.FINAL_WATERMARK
                if wm is not None:
                    ctxs[d].watermarks.set(s, wm)
                    ctxs[d].watermarks.set(WORLD + s, wm)
            after = ctxs[d].last_present_watermark()
            if after is not None and after != before:
                joins[d].handle_watermark(after, ctxs[d], outs[d])
    return [r for o in outs for b in o.batches for r in b.rows()]


def join_worker(rank, port, outdir, join_type):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(WORLD))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    from arroyo_b200.multi_gpu import ShuffleExchange
    cpu = torch.device("cpu")
    ex_l = ShuffleExchange(torch, dist, rank, WORLD, None, cpu, max_recv_rows=1 << 12, n_cols=3)
    ex_r = ShuffleExchange(torch, dist, rank, WORLD, None, cpu, max_recv_rows=1 << 12, n_cols=4)
    lefts, rights, wms = join_shard(rank)
    join = O.InstantJoin(join_cfg(join_type))
    ctx, out = O.OperatorContext(2), O.Collector()  # each edge already min-merged its senders
    for w in range(N_WINDOWS + 1):
        wm = wms[w] if w < N_WINDOWS else O.FINAL_WATERMARK
        for side, (ex, names, key, src) in enumerate(((ex_l, L_NAMES, "id", lefts), (ex_r, R_NAMES, "seller", rights))):
            if w < N_WINDOWS:
                packed, counts = np_pack(src[w], names, key)
                m = src[w].num_rows
            else:
                packed, counts, m = None, None, 0
            got, eff, _ = ex.round_packed(packed, counts, m, wm)
            for cols, r in got:
                join.process_batch_index(side, 2, O.Batch({c: read_ptr(p, r) for c, p in zip(names, cols)}), ctx, out)
            if eff is not None:
                ctx.watermarks.set(side, eff)
        # both edges delivered this round's rows: apply whatever the two inputs' watermarks now allow
        cur = ctx.last_present_watermark()
        if cur is not None and cur != getattr(join, "_applied", None):
            join._applied = cur
            join.handle_watermark(cur, ctx, out)
    rows = [r for b in out.batches for r in b.rows()]
    with open(os.path.join(outdir, f"rank{rank}.json"), "w") as f:
        json.dump(rows, f)
    dist.barrier()
    dist.destroy_process_group()


def _run_join(join_type, port_base):
    want = join_expected(join_type)
    assert len(want) > 200
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(join_worker, args=(port_base + os.getpid() % 1000, d, join_type), nprocs=WORLD, join=True)
        got = [r for k in range(WORLD) for r in json.load(open(os.path.join(d, f"rank{k}.json")))]

    def norm(rows):  # JSON turns the NaN / None of outer joins' missing sides into None
        return [{k: (None if v is None or (isinstance(v, float) and v != v) else v) for k, v in r.items()} for r in rows]
    assert multiset(norm(got)) == multiset(norm(want))


def test_instant_join_behind_two_shuffles_world2_gloo_inner():
    _run_join("inner", 33533)


def test_instant_join_behind_two_shuffles_world2_gloo_full():
    _run_join("full", 34533)

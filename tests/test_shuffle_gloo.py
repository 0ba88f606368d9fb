"""The N>1 path on CPU: ShuffleExchange (control all-gather, variable all-to-all, watermark min-merge)
with world_size 2 over gloo, driven with the oracle's partition function and window operator, against a
single-process simulation of the same two-subtask topology."""
import json
import os
import sys
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import arroyo_oracle as O  # noqa: E402
from tests.golden_cases import multiset  # noqa: E402

S = 1_000_000_000
T0 = 1_700_000_000 * S
WORLD = int(os.environ.get("SHUFFLE_TEST_WORLD", "2"))  # CI runs 2; 4 and 8 were run by hand
BATCH = 500


def shard(rank):
    rng = np.random.default_rng(100 + rank)
    n = 12_000
    ts = T0 + np.arange(n, dtype=np.int64) * (S // 2_000) + rng.integers(0, 40_000_000, n)
    key = rng.integers(0, 400, n, dtype=np.int64) * 104729
    val = rng.integers(-1000, 1000, n, dtype=np.int64)
    return O.source_batches({"key": key, "value": val, O.TIMESTAMP: ts}, BATCH)


def cfg():
    return O.WindowAggConfig(width=3 * S, slide=S, key_names=["key"],
                             aggs=[O.Agg("sum", "value", "sum"), O.Agg("count", None, "n")], window_index=1)


def np_partition(cols, n_rows):
    """ArrowCollector::repartition restated on torch CPU tensors (oracle hash + range formula)."""
    b = O.Batch({"key": cols[0].numpy(), "value": cols[1].numpy(), O.TIMESTAMP: cols[2].numpy()})
    counts = np.zeros(WORLD, dtype=np.int64)
    parts = {d: sb for d, sb in O.repartition(b, ["key"], WORLD)}
    out = [[], [], []]
    for d in range(WORLD):
        if d in parts:
            counts[d] = parts[d].num_rows
            for i, c in enumerate(("key", "value", O.TIMESTAMP)):
                out[i].append(parts[d][c])
    return [torch.from_numpy(np.concatenate(x)) for x in out], torch.from_numpy(counts)


def expected():
    """Sequential simulation: 2 senders, 2 receivers, one batch per round."""
    shards = [shard(r) for r in range(WORLD)]
    gens = [O.WatermarkGenerator() for _ in range(WORLD)]
    ops = [O.SlidingAggregatingWindowFunc(cfg()) for _ in range(WORLD)]
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
    rows = []
    for o in outs:
        for b in o.batches:
            rows += b.rows()
    return rows


def worker(rank, port, outdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(WORLD))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    from arroyo_b200.multi_gpu import ShuffleExchange
    ex = ShuffleExchange(torch, dist, rank, WORLD, np_partition, torch.device("cpu"), max_recv_rows=4 * BATCH, n_cols=3)
    batches = shard(rank)
    gen = O.WatermarkGenerator()
    op = O.SlidingAggregatingWindowFunc(cfg())
    ctx = O.OperatorContext(1)  # the exchange already min-merged the senders' watermarks
    out = O.Collector()
    n_rounds = torch.tensor([len(batches)])
    dist.all_reduce(n_rounds, op=dist.ReduceOp.MAX)
    for i in range(int(n_rounds) + 1):
        if i < len(batches):
            b = batches[i]
            cols = [torch.from_numpy(np.ascontiguousarray(b[c])) for c in ("key", "value", O.TIMESTAMP)]
            wm = gen.process_batch(b[O.TIMESTAMP])
            n = b.num_rows
        else:
            cols = [torch.empty(0, dtype=torch.int64)] * 3
            wm = O.FINAL_WATERMARK if i == int(n_rounds) else None
            n = 0
        rc, n_recv, eff = ex.round(cols, n, wm)
        if n_recv:
            op.process_batch(O.Batch({"key": rc[0].numpy().copy(), "value": rc[1].numpy().copy(),
                                      O.TIMESTAMP: rc[2].numpy().copy()}), ctx, out)
        if eff is not None:
            ctx.watermarks.set(0, eff)
            op.handle_watermark(eff, ctx, out)
    rows = []
    for b in out.batches:
        rows += b.rows()
    with open(os.path.join(outdir, f"rank{rank}.json"), "w") as f:
        json.dump(rows, f)
    dist.barrier()
    dist.destroy_process_group()


def np_pack(batch, names):
    """arroyo_b200_partition_packed restated with the oracle's partition function: destination d's block holds
    its columns back to back."""
    counts = np.zeros(WORLD, dtype=np.int64)
    parts = {d: sb for d, sb in O.repartition(batch, ["key"], WORLD)}
    blocks = []
    for d in range(WORLD):
        if d in parts:
            counts[d] = parts[d].num_rows
            blocks += [np.ascontiguousarray(parts[d][c]).astype(np.int64) for c in names]
    flat = np.concatenate(blocks) if blocks else np.empty(0, dtype=np.int64)
    return torch.from_numpy(flat), torch.from_numpy(counts)


def read_ptr(ptr, n):
    import ctypes
    return np.ctypeslib.as_array((ctypes.c_int64 * n).from_address(ptr)).copy() if n else np.empty(0, dtype=np.int64)


def combiner_worker(rank, port, outdir):
    """partial -> shuffle -> final (SURVEY 8(e) combiner) over ShuffleExchange.round_packed: the local stage is a
    tumbling pre-aggregate of width = slide whose late filter follows the min-merged watermark, the owner merges
    partial rows (SUM of sums, SUM of counts)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(WORLD))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    from arroyo_b200.multi_gpu import ShuffleExchange
    names = ("key", "sum", "n", O.TIMESTAMP)
    ex = ShuffleExchange(torch, dist, rank, WORLD, None, torch.device("cpu"), max_recv_rows=1 << 16, n_cols=4)
    batches = shard(rank)
    gen = O.WatermarkGenerator()
    local = O.TumblingAggregatingWindowFunc(O.WindowAggConfig(width=S, key_names=["key"], aggs=[O.Agg("sum", "value", "sum"), O.Agg("count", None, "n")], final_projection=False))
    owner = O.SlidingAggregatingWindowFunc(O.WindowAggConfig(width=3 * S, slide=S, key_names=["key"], aggs=[O.Agg("sum", "sum", "sum"), O.Agg("sum", "n", "n")], window_index=1))
    lctx, octx = O.OperatorContext(1), O.OperatorContext(1)
    lout, out = O.Collector(), O.Collector()
    n_rounds = torch.tensor([len(batches)])
    dist.all_reduce(n_rounds, op=dist.ReduceOp.MAX)
    multi_round_seen = 0
    for i in range(int(n_rounds) + 1):
        wm = None
        if i < len(batches):
            local.process_batch(batches[i], lctx, lout)
            wm = gen.process_batch(batches[i][O.TIMESTAMP])
        elif i == int(n_rounds):
            wm = O.FINAL_WATERMARK
        eff = ex.exchange_watermark(wm)
        if eff is None:
            continue
        lctx.watermarks.set(0, eff)
        local.handle_watermark(eff, lctx, lout)
        chunks, lout.batches = list(lout.batches), []
        multi_round_seen += len(chunks) > 1
        j = 0
        while True:
            if j < len(chunks):
                packed, counts = np_pack(chunks[j], names)
                m = chunks[j].num_rows
            else:
                packed, counts, m = None, None, 0
            j += 1
            got, _, any_more = ex.round_packed(packed, counts, m, None, more=j < len(chunks))
            for cols, r in got:
                owner.process_batch(O.Batch({c: read_ptr(ptr, r) for c, ptr in zip(names, cols)}), octx, out)
            if not any_more:
                break
        octx.watermarks.set(0, eff)
        owner.handle_watermark(eff, octx, out)
    rows = []
    for b in out.batches:
        rows += b.rows()
    with open(os.path.join(outdir, f"rank{rank}.json"), "w") as f:
        json.dump({"rows": rows, "multi": int(multi_round_seen)}, f)
    dist.barrier()
    dist.destroy_process_group()


def lagged_worker(rank, port, outdir):
    """The same plan through LaggedCombiner: local stage on the caller's thread, shuffle edge + owner stage on a second
    thread, watermarks riding on the data rounds, panes closed two rounds behind."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(WORLD))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    from arroyo_b200.multi_gpu import LaggedCombiner, ShuffleExchange
    names = ("key", "sum", "n", O.TIMESTAMP)
    ex = ShuffleExchange(torch, dist, rank, WORLD, None, torch.device("cpu"), max_recv_rows=1 << 16, n_cols=4)
    batches = shard(rank)
    gen = O.WatermarkGenerator()
    local = O.TumblingAggregatingWindowFunc(O.WindowAggConfig(width=S, key_names=["key"], aggs=[O.Agg("sum", "value", "sum"), O.Agg("count", None, "n")], final_projection=False))
    owner = O.SlidingAggregatingWindowFunc(O.WindowAggConfig(width=3 * S, slide=S, key_names=["key"], aggs=[O.Agg("sum", "sum", "sum"), O.Agg("sum", "n", "n")], window_index=1))
    lctx, octx = O.OperatorContext(1), O.OperatorContext(1)
    lout, out = O.Collector(), O.Collector()

    def local_close(eff):
        lctx.watermarks.set(0, eff)
        local.handle_watermark(eff, lctx, lout)
        chunks, lout.batches = list(lout.batches), []
        return chunks

    def pack(chunk):
        packed, counts = np_pack(chunk, names)
        return packed, counts, chunk.num_rows

    def owner_ingest(got, consume_now):
        for cols, r in got:
            owner.process_batch(O.Batch({c: read_ptr(ptr, r) for c, ptr in zip(names, cols)}), octx, out)

    def owner_watermark(eff):
        octx.watermarks.set(0, eff)
        owner.handle_watermark(eff, octx, out)

    pipe = LaggedCombiner(ex, local_close, pack, owner_ingest, owner_watermark, lag=2)
    n_rounds = torch.tensor([len(batches)])
    dist.all_reduce(n_rounds, op=dist.ReduceOp.MAX)
    n_rounds = int(n_rounds)
    # the final watermark needs `lag` more steps to come back to the local stage, and one more to reach the owner
    for i in range(n_rounds + 4):
        wm = None
        b = batches[i] if i < len(batches) else None
        if b is not None:
            wm = gen.process_batch(b[O.TIMESTAMP])
        elif i == n_rounds:
            wm = O.FINAL_WATERMARK
        pipe.local_step((lambda: local.process_batch(b, lctx, lout)) if b is not None else (lambda: None), wm)
    pipe.drain()
    pipe.close()
    rows = []
    for b in out.batches:
        rows += b.rows()
    with open(os.path.join(outdir, f"rank{rank}.json"), "w") as f:
        json.dump({"rows": rows, "multi": 0}, f)
    dist.barrier()
    dist.destroy_process_group()


def test_lagged_combiner_pipeline_world2_gloo_matches_direct_topology():
    want = expected()
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(lagged_worker, args=(31533 + os.getpid() % 1000, d), nprocs=WORLD, join=True)
        got = []
        for r in range(WORLD):
            got += json.load(open(os.path.join(d, f"rank{r}.json")))["rows"]
    assert multiset(got) == multiset(want)


def test_combiner_plan_world2_gloo_matches_direct_topology():
    want = expected()
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(combiner_worker, args=(30533 + os.getpid() % 1000, d), nprocs=WORLD, join=True)
        got, multi = [], 0
        for r in range(WORLD):
            o = json.load(open(os.path.join(d, f"rank{r}.json")))
            got += o["rows"]
            multi += o["multi"]
    assert multiset(got) == multiset(want)
    assert multi > 0  # the final watermark closes several panes at once: the multi-round path ran


def test_shuffle_exchange_world2_gloo_matches_topology_simulation():
    want = expected()
    assert len(want) > 1000
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(worker, args=(29533 + os.getpid() % 1000, d), nprocs=WORLD, join=True)
        got = []
        owners = []
        for r in range(WORLD):
            rows = json.load(open(os.path.join(d, f"rank{r}.json")))
            got += rows
            owners.append({row["key"] for row in rows})
    assert multiset(got) == multiset(want)
    assert owners[0] and owners[1] and not (owners[0] & owners[1])  # each key lives on exactly one subtask

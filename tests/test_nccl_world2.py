"""The N > 1 path on hardware: two ranks on two GPUs over NCCL, the CUDA operators behind the device partitioner and
`ShuffleExchange` -- the plan shapes of BASELINE configs[2] (raw-row shuffle and partial -> shuffle -> final),
configs[3] (instant join behind two shuffles) and configs[4] (session windows behind one).  Expected = the
single-process simulation of the same 2 x 2 topology with the oracle's operators that the gloo tests use
(tests/test_shuffle_gloo*.py); here every piece on the data path is the product: `arroyo_b200_partition_packed`,
NCCL all-gather / all-to-all, the CUDA operators through the C ABI.  (The reference re-runs every smoke-test query at
parallelism 2 and 3: arroyo-sql-testing/src/smoke_tests.rs.)

Needs two GPUs (`gpurun --gpus 2`); skipped on a one-GPU box."""
import json
import os
import sys
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu
WORLD = 2


def _need_two_gpus():
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < WORLD:
        pytest.skip("needs 2 GPUs (NCCL refuses two ranks on one device)")


def _setup(rank, port):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(WORLD))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=WORLD, device_id=dev)
    torch.cuda.set_stream(torch.cuda.Stream(device=dev))
    return torch, dist, dev


class _Edge:
    """One Shuffle edge: device partitioner + NCCL exchange.  `send(batch or None, wm)` -> ([O.Batch per sender], eff)."""

    def __init__(self, torch, dist, rank, dev, names, key, max_rows=1 << 16):
        from arroyo_b200.multi_gpu import DevicePartitioner, ShuffleExchange
        self.torch, self.dev, self.names, self.key = torch, dev, names, key
        stream = torch.cuda.current_stream().cuda_stream
        self.part = DevicePartitioner(torch, WORLD, len(names), names.index(key), max_rows, rank, stream)
        self.ex = ShuffleExchange(torch, dist, rank, WORLD, None, dev, max_recv_rows=2 * max_rows, n_cols=len(names))

    def send(self, batch, wm, more=False):
        from arroyo_b200.multi_gpu import _Ptr
        from oracle import arroyo_oracle as O
        torch = self.torch
        if batch is not None and batch.num_rows:
            cols = [torch.from_numpy(np.ascontiguousarray(batch[c]).astype(np.int64)).to(self.dev) for c in self.names]
            packed, counts = self.part.pack([c.data_ptr() for c in cols], batch.num_rows)
            m = batch.num_rows
        else:
            packed, counts, m = None, None, 0
        got, eff, any_more = self.ex.round_packed(packed, counts, m, wm, more=more)
        out = []
        for ptrs, r in got:
            out.append(O.Batch({c: torch.as_tensor(_Ptr(p, r), device=self.dev).cpu().numpy().copy()
                                for c, p in zip(self.names, ptrs)}))
        return out, eff, any_more


def _dump(outdir, rank, out):
    rows = [r for b in out.batches for r in b.rows()]
    with open(os.path.join(outdir, f"rank{rank}.json"), "w") as f:
        json.dump(rows, f)


# ---------------------------------------------------------------------------------------------------------------
def raw_rows_worker(rank, port, outdir):
    torch, dist, dev = _setup(rank, port)
    from oracle import arroyo_oracle as O
    from tests import gpu_ops as G
    from tests import test_shuffle_gloo as T
    edge = _Edge(torch, dist, rank, dev, ("key", "value", O.TIMESTAMP), "key")
    batches = T.shard(rank)
    gen, op = O.WatermarkGenerator(), G.SlidingAggregatingWindowFunc(T.cfg())
    ctx, out = O.OperatorContext(1), O.Collector()
    n_rounds = torch.tensor([len(batches)], device=dev)
    dist.all_reduce(n_rounds, op=dist.ReduceOp.MAX)
    n_rounds = int(n_rounds)
    for i in range(n_rounds + 1):
        b = batches[i] if i < len(batches) else None
        wm = gen.process_batch(b[O.TIMESTAMP]) if b is not None else (O.FINAL_WATERMARK if i == n_rounds else None)
        got, eff, _ = edge.send(b, wm)
        for rb in got:
            # the reference's receiver sees one batch per sender
            op.process_batch(rb, ctx, out)
        if eff is not None:
            ctx.watermarks.set(0, eff)
            op.handle_watermark(eff, ctx, out)
    _dump(outdir, rank, out)
    dist.barrier()
    dist.destroy_process_group()


def combiner_worker(rank, port, outdir):
    torch, dist, dev = _setup(rank, port)
    from oracle import arroyo_oracle as O
    from tests import gpu_ops as G
    from tests import test_shuffle_gloo as T
    S = T.S
    names = ("key", "sum", "n", O.TIMESTAMP)
    edge = _Edge(torch, dist, rank, dev, names, "key")
    batches = T.shard(rank)
    gen = O.WatermarkGenerator()
    local = G.TumblingAggregatingWindowFunc(O.WindowAggConfig(
        width=S, key_names=["key"], aggs=[O.Agg("sum", "value", "sum"), O.Agg("count", None, "n")], final_projection=False))
    owner = G.SlidingAggregatingWindowFunc(O.WindowAggConfig(
        width=3 * S, slide=S, key_names=["key"], aggs=[O.Agg("sum", "sum", "sum"), O.Agg("sum", "n", "n")], window_index=1))
    lctx, octx, lout, out = O.OperatorContext(1), O.OperatorContext(1), O.Collector(), O.Collector()
    n_rounds = torch.tensor([len(batches)], device=dev)
    dist.all_reduce(n_rounds, op=dist.ReduceOp.MAX)
    n_rounds = int(n_rounds)
    for i in range(n_rounds + 1):
        wm = None
        if i < len(batches):
            local.process_batch(batches[i], lctx, lout)
            wm = gen.process_batch(batches[i][O.TIMESTAMP])
        elif i == n_rounds:
            wm = O.FINAL_WATERMARK
        eff = edge.ex.exchange_watermark(wm)
        if eff is None:
            continue
        lctx.watermarks.set(0, eff)
        local.handle_watermark(eff, lctx, lout)
        chunks, lout.batches = list(lout.batches), []
        j = 0
        while True:
            chunk = chunks[j] if j < len(chunks) else None
            if chunk is not None:
                chunk = O.Batch({c: chunk[c] for c in names})
            j += 1
            got, _, any_more = edge.send(chunk, None, more=j < len(chunks))
            for rb in got:
                owner.process_batch(rb, octx, out)
            if not any_more:
                break
        octx.watermarks.set(0, eff)
        owner.handle_watermark(eff, octx, out)
    _dump(outdir, rank, out)
    dist.barrier()
    dist.destroy_process_group()


def session_worker(rank, port, outdir):
    torch, dist, dev = _setup(rank, port)
    from oracle import arroyo_oracle as O
    from tests import gpu_ops as G
    from tests import test_shuffle_gloo_ops as T
    edge = _Edge(torch, dist, rank, dev, T.SESS_NAMES, "key", max_rows=1 << 14)
    batches = T.session_shard(rank)
    gen, op = O.WatermarkGenerator(), G.SessionAggregatingWindowFunc(T.session_cfg())
    ctx, out = O.OperatorContext(1), O.Collector()
    n_rounds = torch.tensor([len(batches)], device=dev)
    dist.all_reduce(n_rounds, op=dist.ReduceOp.MAX)
    n_rounds = int(n_rounds)
    for i in range(n_rounds + 1):
        b = batches[i] if i < len(batches) else None
        wm = gen.process_batch(b[O.TIMESTAMP]) if b is not None else (O.FINAL_WATERMARK if i == n_rounds else None)
        got, eff, _ = edge.send(b, wm)
        for rb in got:  # one batch per sender, in sender order: session results depend on what shares a batch
            op.process_batch(rb, ctx, out)
        if eff is not None:
            ctx.watermarks.set(0, eff)
            op.handle_watermark(eff, ctx, out)
    _dump(outdir, rank, out)
    dist.barrier()
    dist.destroy_process_group()


def join_worker(rank, port, outdir, join_type):
    torch, dist, dev = _setup(rank, port)
    from oracle import arroyo_oracle as O
    from tests import gpu_ops as G
    from tests import test_shuffle_gloo_ops as T
    ex_l = _Edge(torch, dist, rank, dev, T.L_NAMES, "id", max_rows=1 << 12)
    ex_r = _Edge(torch, dist, rank, dev, T.R_NAMES, "seller", max_rows=1 << 12)
    lefts, rights, wms = T.join_shard(rank)
    join = G.InstantJoin(T.join_cfg(join_type))
    ctx, out = O.OperatorContext(2), O.Collector()
    applied = None
    for w in range(T.N_WINDOWS + 1):
        wm = wms[w] if w < T.N_WINDOWS else O.FINAL_WATERMARK
        for side, (edge, src) in enumerate(((ex_l, lefts), (ex_r, rights))):
            got, eff, _ = edge.send(src[w] if w < T.N_WINDOWS else None, wm)
            for rb in got:
                join.process_batch_index(side, 2, rb, ctx, out)
            if eff is not None:
                ctx.watermarks.set(side, eff)
        cur = ctx.last_present_watermark()
        if cur is not None and cur != applied:
            applied = cur
            join.handle_watermark(cur, ctx, out)
    _dump(outdir, rank, out)
    dist.barrier()
    dist.destroy_process_group()


def _spawn(fn, port_base, *args):
    import torch.multiprocessing as mp
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(fn, args=(port_base + os.getpid() % 1000, d) + args, nprocs=WORLD, join=True)
        return [r for k in range(WORLD) for r in json.load(open(os.path.join(d, f"rank{k}.json")))]


def _norm(rows):  # JSON turns the NaN / None of outer joins' missing sides into None
    return [{k: (None if v is None or (isinstance(v, float) and v != v) else v) for k, v in r.items()} for r in rows]


def test_raw_row_shuffle_into_the_cuda_sliding_operator_world2_nccl():
    _need_two_gpus()
    from tests import test_shuffle_gloo as T
    from tests.golden_cases import multiset
    assert multiset(_spawn(raw_rows_worker, 41533)) == multiset(T.expected())


def test_combiner_plan_with_cuda_operators_world2_nccl():
    _need_two_gpus()
    from tests import test_shuffle_gloo as T
    from tests.golden_cases import multiset
    assert multiset(_spawn(combiner_worker, 42533)) == multiset(T.expected())


def test_session_windows_behind_the_shuffle_world2_nccl():
    _need_two_gpus()
    from tests import test_shuffle_gloo_ops as T
    from tests.golden_cases import multiset
    assert multiset(_spawn(session_worker, 43533)) == multiset(T.session_expected())


@pytest.mark.parametrize("join_type", ["inner", "full"])
def test_instant_join_behind_two_shuffles_world2_nccl(join_type):
    _need_two_gpus()
    from tests import test_shuffle_gloo_ops as T
    from tests.golden_cases import multiset
    got = _spawn(join_worker, 44533 + (1000 if join_type == "full" else 0), join_type)
    assert multiset(_norm(got)) == multiset(_norm(T.join_expected(join_type)))

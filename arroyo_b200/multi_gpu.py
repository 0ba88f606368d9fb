"""The Shuffle edge across the GPUs of one box: one process per GPU, key-hash partition on the device,
NCCL all-to-all over NVLink, watermark min-merge at the receiver.

Replaces, for edges between GPU-resident operators, `ArrowCollector::collect -> repartition`
(arroyo-operator/src/context.rs:506-616), the queue mesh of a Shuffle edge
(arroyo-worker/src/engine.rs:341-359) and the per-input watermark merge (`WatermarkHolder`,
context.rs:35-86).  Every subtask is both a sender (it ingests its shard of the upstream) and a receiver
(it owns the keys whose hash falls in its range, arroyo-operator/src/lib.rs:30-41).

Per round every rank
  1. buckets its rows by destination (`arroyo_b200_partition`: histogram -> scan -> scatter; the
     per-destination segments are the all-to-all send buffers),
  2. all-gathers one control row {rows for each destination, its watermark or none}: afterwards every
     rank knows its receive counts and every sender's watermark (signals are broadcast to every
     downstream queue in the reference too, context.rs:663-677),
  3. exchanges the column segments with `all_to_all_single` (variable splits),
  4. hands the received rows to its window operator and then applies the min-merged watermark.
No reduction collective is needed: each key lives on exactly one GPU.

`ShuffleExchange` takes the partition function as a parameter so the protocol can be exercised on CPU
tensors with the `gloo` backend (tests/test_shuffle_gloo.py); the product path uses `DevicePartitioner`.
"""
import ctypes as C
from typing import Callable, List, Optional, Sequence, Tuple

from . import ffi
from .context import WatermarkHolder

NO_WM = -(1 << 63)


class DevicePartitioner:
    """arroyo_b200_partition over torch device tensors (the product path)."""

    def __init__(self, torch, world: int, n_cols: int, key_col: int, max_rows: int, device: int, stream: int = 0):
        self.torch = torch
        self.lib = ffi.load()
        self.h = C.c_void_p()
        st = self.lib.arroyo_b200_partitioner_create(device, stream, world, n_cols, key_col, max_rows, C.byref(self.h))
        if st != ffi.OK:
            raise ffi.ArroyoB200Error(st, "partitioner_create failed")
        dev = torch.device("cuda", device)
        self.out = [torch.empty(max_rows, dtype=torch.int64, device=dev) for _ in range(n_cols)]
        self.counts = torch.zeros(world, dtype=torch.int64, device=dev)
        self.offsets = torch.zeros(world, dtype=torch.int64, device=dev)
        self.n_cols = n_cols

    def __call__(self, cols: Sequence, n_rows: int):
        inp = (C.c_uint64 * self.n_cols)(*[c.data_ptr() for c in cols])
        outp = (C.c_uint64 * self.n_cols)(*[c.data_ptr() for c in self.out])
        st = self.lib.arroyo_b200_partition(self.h, inp, n_rows, outp, self.counts.data_ptr(), self.offsets.data_ptr())
        if st != ffi.OK:
            raise ffi.ArroyoB200Error(st, "partition failed")
        return [o[:n_rows] for o in self.out], self.counts

    def close(self):
        if self.h:
            self.lib.arroyo_b200_partitioner_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ShuffleExchange:
    """One Shuffle edge between `world` subtasks.  `partition_fn(cols, n_rows) -> (cols bucketed by
    destination in destination order, counts[world] tensor)`."""

    def __init__(self, torch, dist, rank: int, world: int, partition_fn: Callable, device, max_recv_rows: int,
                 n_cols: int):
        self.torch, self.dist, self.rank, self.world = torch, dist, rank, world
        self.partition_fn = partition_fn
        self.device = device
        self.ctrl = torch.zeros(world + 1, dtype=torch.int64, device=device)
        self.ctrl_all = torch.zeros(world * (world + 1), dtype=torch.int64, device=device)
        self.recv = [torch.empty(max_recv_rows, dtype=torch.int64, device=device) for _ in range(n_cols)]
        self.max_recv_rows = max_recv_rows
        self.holder = WatermarkHolder(world)
        self.bytes_sent = 0

    def exchange_watermark(self, watermark: Optional[int]) -> Optional[int]:
        """Broadcasts this sender's watermark (or none) and returns the min-merged effective watermark if it
        advanced (signals go to every downstream queue, context.rs:663-677; merge = WatermarkHolder)."""
        torch, dist, W = self.torch, self.dist, self.world
        self.ctrl[:W] = 0
        self.ctrl[W] = NO_WM if watermark is None else int(min(watermark, (1 << 63) - 1))
        dist.all_gather_into_tensor(self.ctrl_all, self.ctrl)
        m = self.ctrl_all.view(W, W + 1)[:, W].cpu().tolist()
        before = self.holder.last_present_watermark
        for s, wm in enumerate(m):
            if wm != NO_WM:
                self.holder.set(s, wm)
        after = self.holder.last_present_watermark
        return after if after is not None and after != before else None

    def round(self, cols: Sequence, n_rows: int, watermark: Optional[int]) -> Tuple[List, int, Optional[int]]:
        """Sends this rank's rows, returns (received columns, received rows, effective watermark after
        this round or None if it did not advance)."""
        torch, dist, W = self.torch, self.dist, self.world
        if n_rows > 0:
            send_cols, counts = self.partition_fn(cols, n_rows)
            self.ctrl[:W] = counts
        else:
            send_cols = [c[:0] for c in cols]
            self.ctrl[:W] = 0
        self.ctrl[W] = NO_WM if watermark is None else int(min(watermark, (1 << 63) - 1))
        dist.all_gather_into_tensor(self.ctrl_all, self.ctrl)
        m = self.ctrl_all.view(W, W + 1).cpu()  # the one host sync of the round: split sizes
        send_splits = m[self.rank, :W].tolist()
        recv_splits = m[:, self.rank].tolist()
        n_recv = int(sum(recv_splits))
        if n_recv > self.max_recv_rows:
            raise RuntimeError(f"shuffle receive buffer too small: {n_recv} > {self.max_recv_rows}")
        out = []
        for c, r in zip(send_cols, self.recv):
            o = r[:n_recv]
            dist.all_to_all_single(o, c, recv_splits, send_splits)
            out.append(o)
        self.bytes_sent += 8 * len(send_cols) * (n_rows - send_splits[self.rank])
        before = self.holder.last_present_watermark
        for s in range(W):
            wm = int(m[s, W])
            if wm != NO_WM:
                self.holder.set(s, wm)
        after = self.holder.last_present_watermark
        return out, n_recv, (after if after is not None and after != before else None)


# ------------------------------------------------------------------------------------------------
# N > 1 benchmark (called from bench.py under torchrun)
# ------------------------------------------------------------------------------------------------
class _Ptr:
    """Wraps a raw device pointer as a torch tensor (__cuda_array_interface__)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 2}


def _e2e_partials(args, torch, dist, B, ab, native, rank, world, local, device, gen_pane, part, ex):
    """End to end at N GPUs: every rank feeds its shard as pinned host Arrow batches through
    arroyo_b200_op_process_batch (host -> device copies inside the timed region), partial aggregates cross the
    all-to-all, and each rank reads the windows of its keys back as host Arrow batches."""
    import time

    import pyarrow as pa
    rows = args.rows_per_pane
    nb = rows // B.BATCH_ROWS
    K = args.e2e_steps or min(args.steps, 6)
    W = 13
    batches, wms, _keep = B.host_feed(torch, gen_pane, range(W + K), rows)
    stream = torch.cuda.current_stream().cuda_stream
    local_cfg = ab.WindowAggConfig(width=B.SLIDE, key_names=["key"],
                                   aggs=[ab.Agg("sum", "value", "sum"), ab.Agg("count", None, "count")],
                                   final_projection=False)
    local_op = native.TumblingAggregatingWindowFunc(local_cfg, device=local, stream=stream, flags=B.op_flags(args),
                                                    expected_keys=args.keys, task_index=rank, parallelism=world)
    owner_cfg = ab.WindowAggConfig(width=B.WIDTH, slide=B.SLIDE, key_names=["key"],
                                   aggs=[ab.Agg("sum", "sum", "sum"), ab.Agg("avg", "sum", "avg"),
                                         ab.Agg("count", None, "count")], window_index=1, partial_count_col="count")
    p_schema = pa.schema([("key", pa.int64()), ("sum", pa.int64()), ("count", pa.int64()), ("_timestamp", pa.timestamp("ns"))])
    owner_op = native.SlidingAggregatingWindowFunc(owner_cfg, input_schema=p_schema, device=local, stream=stream,
                                                   flags=B.op_flags(args), expected_keys=max(2 * args.keys // world, 1024),
                                                   task_index=rank, parallelism=world)
    ex.holder = type(ex.holder)(world)  # fresh watermark state for this pass
    lctx, octx, col = ab.OperatorContext(1), ab.OperatorContext(1), ab.Collector()
    empty_cols = [torch.empty(0, dtype=torch.int64, device=device) for _ in range(4)]
    part_rows = part.out[0].numel()
    d2h = 0

    def step(p):
        nonlocal d2h
        for b in range(nb):
            local_op.process_batch(batches[p][b], lctx, col)
            wm = wms[p * nb + b]
            if wm is None:
                continue
            eff = ex.exchange_watermark(wm)
            if eff is None:
                continue
            chunks = []
            for n, cols in local_op.handle_watermark_device(eff):
                for o in range(0, n, part_rows):
                    m = min(part_rows, n - o)
                    chunks.append(([torch.as_tensor(_Ptr(c + 8 * o, m), device=device) for c in cols], m))
            n_rounds = torch.tensor([len(chunks)], dtype=torch.int64, device=device)
            dist.all_reduce(n_rounds, op=dist.ReduceOp.MAX)
            for r in range(int(n_rounds.item())):
                tc, m = chunks[r] if r < len(chunks) else (empty_cols, 0)
                rc, n_recv, _ = ex.round(tc, m, None)
                if n_recv:
                    owner_op.process_device_batch([c.data_ptr() for c in rc], n_recv)
                    owner_op.flush()
            octx.watermarks.set(0, eff)
            owner_op.handle_watermark(eff, octx, col)
            for rb in col.batches:
                d2h += rb.num_rows * 48
            col.batches.clear()

    for p in range(W):
        step(p)
    owner_op.flush()
    torch.cuda.synchronize()
    dist.barrier()
    d2h = 0
    t0 = time.perf_counter()
    for p in range(W, W + K):
        step(p)
    owner_op.flush()
    local_op.flush()
    torch.cuda.synchronize()
    dist.barrier()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    tot = torch.tensor([d2h], dtype=torch.int64, device=device)
    dist.all_reduce(tot)
    owner_op.close()
    local_op.close()
    dt = float(dt.item())
    return {"value": world * K * rows / dt, "unit": "rows/s", "h2d_bytes_per_step": world * rows * 24,
            "d2h_bytes_per_step": int(tot.item()) // max(K, 1), "steps": K, "ms_per_step": 1e3 * dt / K,
            "path": "per GPU: pinned host Arrow batches -> arroyo_b200_op_process_batch -> partials over NCCL all-to-all "
                    "-> host Arrow windows"}


def bench(args, torch, dist, rank, world, local):
    """Weak scaling: every GPU ingests its own 16 Mi-row/pane shard of the stream.

    --shuffle partials (default): partial -> shuffle -> final.  Each GPU pre-aggregates its shard per pane
        (the same ingest kernel), and when a pane can no longer receive rows its partial rows
        (key, sum, count) are hash-partitioned on the device and exchanged with an NCCL all-to-all; the
        owner of a key merges the partials into its sliding-window state and emits.  Same results as
        shuffling raw rows (combiner), 1/16 of the bytes over NVLink.
    --shuffle rows: the reference's plan shape -- raw rows are partitioned and exchanged, each GPU
        aggregates only the keys it owns."""
    import json

    import pyarrow as pa

    import arroyo_b200 as ab
    import bench as B
    from . import operators as native

    device = torch.device("cuda", local)
    W, K = max(args.warmup, 3), args.steps
    rows = args.rows_per_pane
    nb = rows // B.BATCH_ROWS
    gen_pane = B.make_generator(torch, device, rows, args.keys, args.dist, 42 + rank)
    panes = [gen_pane(p) for p in range(W + K)]
    mins, maxs = [], []
    for (_, _, t) in panes:
        tb = t.view(nb, B.BATCH_ROWS)
        mins.append(tb.amin(dim=1))
        maxs.append(tb.amax(dim=1))
    mins = torch.stack(mins).cpu().numpy().reshape(-1).tolist()
    maxs = torch.stack(maxs).cpu().numpy().reshape(-1).tolist()
    wms = B.watermark_schedule(list(zip(mins, maxs)))  # this rank's own WatermarkGenerator
    torch.cuda.synchronize()

    stream = torch.cuda.current_stream().cuda_stream
    mode = args.shuffle
    flags = ffi.FLAG_PROFILE | B.op_flags(args)
    raw_schema = pa.schema([("key", pa.int64()), ("value", pa.int64()), ("_timestamp", pa.timestamp("ns"))])
    if mode == "partials":
        part_rows = 1 << 22
        local_cfg = ab.WindowAggConfig(width=B.SLIDE, key_names=["key"],
                                       aggs=[ab.Agg("sum", "value", "sum"), ab.Agg("count", None, "count")],
                                       final_projection=False)
        local_op = native.TumblingAggregatingWindowFunc(local_cfg, input_schema=raw_schema, device=local, stream=stream,
                                                        flags=flags, expected_keys=args.keys, task_index=rank,
                                                        parallelism=world)
        owner_cfg = ab.WindowAggConfig(width=B.WIDTH, slide=B.SLIDE, key_names=["key"],
                                       aggs=[ab.Agg("sum", "sum", "sum"), ab.Agg("avg", "sum", "avg"),
                                             ab.Agg("count", None, "count")], window_index=1, partial_count_col="count")
        p_schema = pa.schema([("key", pa.int64()), ("sum", pa.int64()), ("count", pa.int64()),
                              ("_timestamp", pa.timestamp("ns"))])
        owner_op = native.SlidingAggregatingWindowFunc(owner_cfg, input_schema=p_schema, device=local, stream=stream,
                                                       flags=B.op_flags(args), expected_keys=max(2 * args.keys // world, 1024),
                                                       task_index=rank, parallelism=world)
        n_cols = 4
    else:
        part_rows = 64 * B.BATCH_ROWS
        local_op = None
        owner_op = native.SlidingAggregatingWindowFunc(B.window_config(), input_schema=raw_schema, device=local,
                                                       stream=stream, flags=flags,
                                                       expected_keys=max(2 * args.keys // world, 1024), task_index=rank,
                                                       parallelism=world)
        n_cols = 3
    part = DevicePartitioner(torch, world, n_cols, 0, part_rows, local, stream)
    ex = ShuffleExchange(torch, dist, rank, world, part, device, max_recv_rows=2 * part_rows, n_cols=n_cols)
    rows_out = 0
    empty_cols = [torch.empty(0, dtype=torch.int64, device=device) for _ in range(n_cols)]

    def emit(eff):
        nonlocal rows_out
        for n, _ in owner_op.handle_watermark_device(eff):
            rows_out += n

    def step_partials(p):
        k, v, t = panes[p]
        start = 0
        for b in range(nb):
            wm = wms[p * nb + b]
            if wm is None and b != nb - 1:
                continue
            s, e = start * B.BATCH_ROWS, (b + 1) * B.BATCH_ROWS
            local_op.process_device_batch([k.data_ptr() + 8 * s, v.data_ptr() + 8 * s, t.data_ptr() + 8 * s], e - s)
            start = b + 1
            if wm is None:
                continue
            eff = ex.exchange_watermark(wm)
            if eff is None:
                continue
            # panes that can no longer receive rows leave the local stage as partial rows; every rank takes
            # part in the same number of exchange rounds (ranks with nothing left send empty segments)
            chunks = []
            for n, cols in local_op.handle_watermark_device(eff):
                for o in range(0, n, part_rows):
                    m = min(part_rows, n - o)
                    chunks.append(([torch.as_tensor(_Ptr(c + 8 * o, m), device=device) for c in cols], m))
            n_rounds = torch.tensor([len(chunks)], dtype=torch.int64, device=device)
            dist.all_reduce(n_rounds, op=dist.ReduceOp.MAX)
            for r in range(int(n_rounds.item())):
                tc, m = chunks[r] if r < len(chunks) else (empty_cols, 0)
                rc, n_recv, _ = ex.round(tc, m, None)
                if n_recv:
                    owner_op.process_device_batch([c.data_ptr() for c in rc], n_recv)
                    owner_op.flush()  # receive buffers are reused
            emit(eff)

    def step_rows(p):
        k, v, t = panes[p]
        rb = part_rows // B.BATCH_ROWS
        for r0 in range(0, nb, rb):
            r1 = min(r0 + rb, nb)
            s, e = r0 * B.BATCH_ROWS, r1 * B.BATCH_ROWS
            wm = None
            for b in range(r0, r1):
                if wms[p * nb + b] is not None:
                    wm = wms[p * nb + b]
            cols, n_recv, eff = ex.round([k[s:e], v[s:e], t[s:e]], e - s, wm)
            if n_recv:
                owner_op.process_device_batch([c.data_ptr() for c in cols], n_recv)
            if eff is not None:
                emit(eff)
            else:
                owner_op.flush()  # the receive buffers are reused by the next round

    step = step_partials if mode == "partials" else step_rows
    timed_op = local_op if mode == "partials" else owner_op
    for p in range(W):
        step(p)
    owner_op.flush()
    torch.cuda.synchronize()
    dist.barrier()
    st0 = timed_op.stats()
    so0 = owner_op.stats()
    rows_out = 0
    sent0 = ex.bytes_sent
    sampler = B.ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for p in range(W, W + K):
        step(p)
    owner_op.flush()
    if local_op is not None:
        local_op.flush()
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    clocks = sampler.stop() if rank == 0 else None
    st1 = timed_op.stats()
    so1 = owner_op.stats()
    d = {k: st1[k] - st0[k] for k in st1}
    launches = d["kernel_launches"] + (so1["kernel_launches"] - so0["kernel_launches"] if local_op is not None else 0)
    tot = torch.tensor([launches, rows_out, d["rows_in"]], dtype=torch.int64, device=device)
    dist.all_reduce(tot)
    sent = ex.bytes_sent - sent0
    owner_op.close()
    if local_op is not None:
        local_op.close()
    e2e = None
    if mode == "partials" and not args.skip_e2e:
        e2e = _e2e_partials(args, torch, dist, B, ab, native, rank, world, local, device, gen_pane, part, ex)
    part.close()
    if rank == 0:
        peak, peak_kind = B.measured_peak()
        ingest_gbs = 24.0 * d["ingest_rows_timed"] / (d["ingest_ms"] * 1e-3) / 1e9 if d["ingest_ms"] else None
        out = {"metric": "rows/sec sliding-window SUM (1M keys)", "value": world * K * rows / (ms * 1e-3),
               "unit": "rows/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
               "config": {"workload": "BASELINE configs[2]: hop(1s slide,10s width) SUM/AVG/COUNT GROUP BY key, "
                                      f"{args.keys} i64 keys ({args.dist}); every GPU ingests {rows} rows/pane "
                                      f"({rows // B.BATCH_ROWS} batches of {B.BATCH_ROWS}); key-hash shuffle over an NCCL "
                                      "all-to-all; each GPU emits the windows of the keys it owns",
                          "keys": args.keys, "rows_per_step_per_gpu": rows, "batch_rows": B.BATCH_ROWS,
                          "shuffle": ("partial aggregates per pane (partial -> shuffle -> final)" if mode == "partials"
                                      else "raw rows (reference plan shape)"),
                          "l2": "inputs larger than L2, never re-read", "parallelism": f"key-partitioned x{world}"},
               "rows_out_per_step": int(tot[1].item()) / max(K, 1), "gpu_launches": int(tot[0].item()),
               "roofline": {"bound": "hbm", "kernel": "ingest_kernel<1>",
                            "achieved": round(ingest_gbs, 1) if ingest_gbs else None, "peak": peak,
                            "peak_kind": peak_kind, "unit": "GB/s",
                            "frac": round(ingest_gbs / peak, 4) if ingest_gbs else None, "traffic": None,
                            "note": "rank 0's raw-row ingest kernel"},
               "e2e": e2e, "clocks": clocks,
               "shuffle_bytes_sent_per_step_per_gpu": sent // max(K, 1)}
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()

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
def bench(args, torch, dist, rank, world, local):
    import json

    import pyarrow as pa

    import bench as B
    from . import operators as native

    device = torch.device("cuda", local)
    W, K = max(args.warmup, 3), args.steps
    rows = args.rows_per_pane
    nb = rows // B.BATCH_ROWS
    round_batches = 64
    gen_pane = B.make_generator(torch, device, rows, args.keys, args.dist, 42 + rank)
    panes = [gen_pane(p) for p in range(W + K)]
    # this rank's watermark per batch (its own WatermarkGenerator)
    mins, maxs = [], []
    for (_, _, t) in panes:
        tb = t.view(nb, B.BATCH_ROWS)
        mins.append(tb.amin(dim=1))
        maxs.append(tb.amax(dim=1))
    mins = torch.stack(mins).cpu().numpy().reshape(-1).tolist()
    maxs = torch.stack(maxs).cpu().numpy().reshape(-1).tolist()
    wms = B.watermark_schedule(list(zip(mins, maxs)))
    torch.cuda.synchronize()

    stream = torch.cuda.current_stream().cuda_stream
    round_rows = round_batches * B.BATCH_ROWS
    part = DevicePartitioner(torch, world, 3, 0, round_rows, local, stream)
    ex = ShuffleExchange(torch, dist, rank, world, part, device, max_recv_rows=2 * round_rows, n_cols=3)
    schema = pa.schema([("key", pa.int64()), ("value", pa.int64()), ("_timestamp", pa.timestamp("ns"))])
    flags = ffi.FLAG_PROFILE | B.op_flags(args)
    op = native.SlidingAggregatingWindowFunc(B.window_config(), input_schema=schema, device=local, stream=stream,
                                             flags=flags, expected_keys=max(args.keys * 2 // world, 1024),
                                             task_index=rank, parallelism=world)
    rows_out = 0

    def step(p):
        nonlocal rows_out
        k, v, t = panes[p]
        for r0 in range(0, nb, round_batches):
            r1 = min(r0 + round_batches, nb)
            s, e = r0 * B.BATCH_ROWS, r1 * B.BATCH_ROWS
            wm = None
            for b in range(r0, r1):
                if wms[p * nb + b] is not None:
                    wm = wms[p * nb + b]
            cols, n_recv, eff = ex.round([k[s:e], v[s:e], t[s:e]], e - s, wm)
            if n_recv:
                op.process_device_batch([c.data_ptr() for c in cols], n_recv)
            if eff is not None:
                for n, _ in op.handle_watermark_device(eff):
                    rows_out += n
            else:
                op.flush()  # the receive buffers are reused by the next round

    for p in range(W):
        step(p)
    op.flush()
    torch.cuda.synchronize()
    dist.barrier()
    st0 = op.stats()
    rows_out = 0
    sampler = B.ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for p in range(W, W + K):
        step(p)
    op.flush()
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    clocks = sampler.stop() if rank == 0 else None
    st1 = op.stats()
    d = {k: st1[k] - st0[k] for k in st1}
    tot = torch.tensor([d["kernel_launches"], rows_out, d["rows_in"]], dtype=torch.int64, device=device)
    dist.all_reduce(tot)
    op.close()
    part.close()
    if rank == 0:
        peak, peak_kind = B.measured_peak()
        ingest_gbs = 24.0 * d["ingest_rows_timed"] / (d["ingest_ms"] * 1e-3) / 1e9 if d["ingest_ms"] else None
        out = {"metric": "rows/sec sliding-window SUM (1M keys)", "value": world * K * rows / (ms * 1e-3),
               "unit": "rows/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
               "config": {"workload": "BASELINE configs[2]: hop(1s slide,10s width) SUM/AVG/COUNT GROUP BY key, "
                                      f"{args.keys} i64 keys ({args.dist}); every GPU ingests {rows} rows/pane "
                                      f"({rows // B.BATCH_ROWS} batches of {B.BATCH_ROWS}), key-hash shuffle over NCCL "
                                      "all-to-all, each GPU aggregates the keys it owns",
                          "keys": args.keys, "rows_per_step_per_gpu": rows, "batch_rows": B.BATCH_ROWS,
                          "round_rows": round_rows, "l2": "inputs larger than L2, never re-read",
                          "parallelism": f"key-partitioned x{world}"},
               "rows_out_per_step": int(tot[1].item()) / max(K, 1), "gpu_launches": int(tot[0].item()) + 3 * 4 * K * world,
               "roofline": {"bound": "hbm", "kernel": "ingest_kernel<1>",
                            "achieved": round(ingest_gbs, 1) if ingest_gbs else None, "peak": peak,
                            "peak_kind": peak_kind, "unit": "GB/s",
                            "frac": round(ingest_gbs / peak, 4) if ingest_gbs else None, "traffic": None,
                            "note": "rank 0's ingest kernel; the step is bounded by the shuffle "
                                    "(24 B/row x (N-1)/N over NVLink) + partition + ingest"},
               "e2e": None, "clocks": clocks,
               "shuffle_bytes_sent_per_step_per_gpu": ex.bytes_sent // max(W + K, 1)}
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()

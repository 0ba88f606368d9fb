"""Device-resident rates of every operator on the path at BASELINE-config shapes, one GPU.

Not the headline (bench.py measures that): this is the per-operator evidence behind DESIGN.md's kernel table.
Each case feeds synthetic batches that already live in HBM through the C ABI's device-batch entry points, times
K repetitions with CUDA events after a warm-up, and reports rows/s and the algorithmic GB/s.

    python tools/operator_rates.py > gpurun_out/operator_rates.json
"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

S = 1_000_000_000
T0 = 1_700_000_000 * S


def timed(torch, fn, reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for r in range(reps):
        fn(r)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def cpu_join_rate(n_p, n_a, windows=3):
    """The C restatement of the reference's instant join (oracle/join_oracle.c) on the q8 shape, one thread -- the
    reference runs one task per subtask.  Returns input rows/s."""
    import time

    import numpy as np
    from oracle import arroyo_oracle as O, c_oracle
    rng = np.random.default_rng(1)
    pid = rng.permutation(n_p).astype(np.int64) + 1000
    name = rng.integers(0, 10**6, n_p, dtype=np.int64)
    seller = rng.integers(0, n_p, n_a, dtype=np.int64) + 1000
    auction = np.arange(n_a, dtype=np.int64)
    reserve = rng.integers(1, 10**5, n_a, dtype=np.int64)
    join = c_oracle.InstantJoin(O.JoinConfig(left_on=["id"], right_on=["seller"], join_type="inner"))
    ctx, out = O.OperatorContext(2), O.Collector()
    t0 = None
    for w in range(windows + 1):
        if w == 1:
            t0 = time.perf_counter()  # the first window warms the allocations up
        ts = T0 + (w + 1) * 30 * S - 1
        join.process_batch_index(0, 2, O.Batch({"id": pid, "name_code": name, O.TIMESTAMP: np.full(n_p, ts, dtype=np.int64)}), ctx, out)
        join.process_batch_index(1, 2, O.Batch({"seller": seller, "auction": auction, "reserve": reserve,
                                                O.TIMESTAMP: np.full(n_a, ts, dtype=np.int64)}), ctx, out)
        for side in (0, 1):
            ctx.watermarks.set(side, ts + 1)
        join.handle_watermark(ts + 1, ctx, out)
        out.batches.clear()
    return windows * (n_p + n_a) / (time.perf_counter() - t0)


def cpu_session_rate(n_keys, srows, steps=6, warm=8):
    """The C restatement of the reference's session operator (oracle/session_oracle.c), one thread."""
    import time

    import numpy as np
    from oracle import arroyo_oracle as O, c_oracle
    rng = np.random.default_rng(1)
    op = c_oracle.SessionAggregatingWindowFunc(O.SessionConfig(
        gap=5 * S, key_names=["key"], aggs=[O.Agg("sum", "value", "sum"), O.Agg("count", None, "n")], window_index=1))
    ctx, out = O.OperatorContext(1), O.Collector()
    sv = rng.integers(0, 10**6, srows, dtype=np.int64)
    offs = np.sort(rng.integers(0, S, srows, dtype=np.int64))
    t0 = None
    for p in range(warm + steps):
        if p == warm:
            t0 = time.perf_counter()
        key = rng.integers(0, n_keys, srows, dtype=np.int64) * 7919
        op.process_batch(O.Batch({"key": key, "value": sv, O.TIMESTAMP: offs + (T0 + p * S)}), ctx, out)
        ctx.watermarks.set(0, T0 + p * S - S)
        op.handle_watermark(T0 + p * S - S, ctx, out)
        out.batches.clear()
    return steps * srows / (time.perf_counter() - t0)


def main():
    if "--cpu-only" in sys.argv:
        # the CPU restatements alone (no GPU needed): same shapes as the GPU cases below
        print(json.dumps({"join_cpu_rows_per_s_1_thread": cpu_join_rate(1 << 21, 1 << 23),
                          "session_cpu_rows_per_s_1_thread": cpu_session_rate(int(os.environ.get("SESSION_KEYS", 10_000_000)), 1 << 22)},
                         indent=1))
        return
    import pyarrow as pa
    import torch

    import arroyo_b200 as ab
    import bench as B
    from arroyo_b200 import ffi, operators as native
    from arroyo_b200.multi_gpu import DevicePartitioner

    only = {a for a in sys.argv[1:] if not a.startswith("--")}  # e.g. `join session`; empty = everything

    def want(section):
        return not only or section in only

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    # one explicit stream for torch, the operators and the partitioner (a stream argument of 0 would make each
    # handle create its own stream, invisible to the events below)
    torch.cuda.set_stream(torch.cuda.Stream(device=dev))
    stream = torch.cuda.current_stream().cuda_stream
    assert stream != 0
    peak, _ = B.measured_peak()
    out = []
    g = torch.Generator(device=dev)
    g.manual_seed(1)

    def rec(name, config, rows, ms, bytes_per_row, note="", cpu=None):
        gbs = rows * bytes_per_row / (ms * 1e-3) / 1e9
        out.append({"operator": name, "config": config, "rows_per_step": rows, "ms_per_step": round(ms, 4),
                    "rows_per_s": rows / (ms * 1e-3), "algorithmic_bytes_per_row": bytes_per_row,
                    "achieved_GBps": round(gbs, 1), "frac_of_measured_hbm_peak": round(gbs / peak, 4), "note": note})
        if cpu is not None:  # the C restatement of the reference algorithm on one host thread, same shape
            out[-1]["cpu_port_rows_per_s"] = cpu
            out[-1]["cpu_port_threads"] = 1
        print(f"# {name}: {rows / (ms * 1e-3) / 1e9:.2f} G rows/s, {gbs:.0f} GB/s", file=sys.stderr, flush=True)

    raw_schema = pa.schema([("key", pa.int64()), ("value", pa.int64()), ("_timestamp", pa.timestamp("ns"))])
    rows = 1 << 24

    # ---- configs[1]: tumbling 1 s COUNT(*) GROUP BY key, 10 K keys --------------------------------------
    for keyspace in (("scattered", "dense") if want("tumbling") else ()):
        gen = B.make_generator(torch, dev, rows, 10_000, "uniform", 42, keyspace)
        W, K = 6, 20
        panes = [gen(p) for p in range(W + K)]
        cfg = ab.WindowAggConfig(width=S, key_names=["key"], aggs=[ab.Agg("count", None, "count")], window_index=1)
        op = native.TumblingAggregatingWindowFunc(cfg, input_schema=raw_schema, device=0, stream=stream, expected_keys=10_000)

        def step(p):
            k, v, t = panes[p]
            op.process_device_batch([k.data_ptr(), v.data_ptr(), t.data_ptr()], rows)
            op.handle_watermark_device(T0 + p * S)  # pane p has just begun: closes pane p - 1

        for p in range(W):
            step(p)
        ms = timed(torch, lambda r: step(W + r), K)
        rec("TumblingAggregatingWindowFunc", f"configs[1]: 1 s COUNT(*) GROUP BY key, 10 K keys ({keyspace}), 16 Mi rows/window",
            rows, ms, 16, "key + _timestamp read once (COUNT(*) reads no value column)")
        op.close()
        del panes, gen
        torch.cuda.empty_cache()

    # ---- shuffle: hash partition of raw rows into 8 destinations ----------------------------------------
    gen = B.make_generator(torch, dev, rows, 1 << 20, "uniform", 42)
    k, v, t = gen(0)
    for packed in ((False, True) if want("shuffle") else ()):
        part = DevicePartitioner(torch, 8, 3, 0, rows, 0, stream)
        fn = (lambda r: part.pack([k.data_ptr(), v.data_ptr(), t.data_ptr()], rows)) if packed else (lambda r: part([k, v, t], rows))
        for _ in range(3):
            fn(0)
        ms = timed(torch, fn, 20)
        rec("repartition (shuffle.cu)", f"16 Mi rows x 3 columns -> 8 destinations ({'packed' if packed else 'per-column'} layout)",
            rows, ms, 48 + 8, "24 B read + 24 B written per row, key read once more by the histogram pass")
        part.close()

    def wm_section():
        # ---- WatermarkGenerator reductions --------------------------------------------------------------------
        lib = ffi.load()
        mn, mx = C.c_int64(), C.c_int64()
        fn = lambda r: lib.arroyo_b200_ts_minmax(0, stream, t.data_ptr(), rows, C.byref(mn), C.byref(mx))
        fn(0)
        ms = timed(torch, fn, 20)
        rec("WatermarkGenerator min/max", "one 16 Mi-row batch (synchronous call: includes the result's D2H)", rows, ms, 8)
        nb = 1 << 16
        fn = lambda r: lib.arroyo_b200_ts_minmax(0, stream, t.data_ptr(), nb, C.byref(mn), C.byref(mx))
        ms = timed(torch, fn, 200)
        rec("WatermarkGenerator min/max", "one 64 Ki-row batch (latency-bound: launch + D2H + sync)", nb, ms, 8)

    def join_section():
        # ---- configs[3]: q8-shaped windowed join, person x auction per 30 s tumbling window ----------------
        n_p, n_a = 1 << 21, 1 << 23
        W30 = 30 * S
        jcfg = ab.JoinConfig(left_on=["id"], right_on=["seller"], join_type="inner")
        l_schema = pa.schema([("id", pa.int64()), ("name_code", pa.int64()), ("_timestamp", pa.timestamp("ns"))])
        r_schema = pa.schema([("seller", pa.int64()), ("auction", pa.int64()), ("reserve", pa.int64()), ("_timestamp", pa.timestamp("ns"))])
        jop = native.InstantJoin(jcfg, left_schema=l_schema, right_schema=r_schema, device=0, stream=stream)
        pid = torch.randperm(n_p, device=dev, generator=g).to(torch.int64) + 1000
        name = torch.randint(0, 10**6, (n_p,), device=dev, generator=g, dtype=torch.int64)
        seller = torch.randint(0, n_p, (n_a,), device=dev, generator=g, dtype=torch.int64) + 1000
        auction = torch.arange(n_a, device=dev, dtype=torch.int64)
        reserve = torch.randint(1, 10**5, (n_a,), device=dev, generator=g, dtype=torch.int64)
        rows_out = 0

        def jstep(w):
            nonlocal rows_out
            ts = T0 + (w + 1) * W30 - 1
            tl = torch.full((n_p,), ts, device=dev, dtype=torch.int64)
            tr = torch.full((n_a,), ts, device=dev, dtype=torch.int64)
            la = (C.c_uint64 * 3)(pid.data_ptr(), name.data_ptr(), tl.data_ptr())
            ra = (C.c_uint64 * 4)(seller.data_ptr(), auction.data_ptr(), reserve.data_ptr(), tr.data_ptr())
            native._check(jop._lib, jop._h, jop._lib.arroyo_b200_op_process_device_batch(jop._h, 0, 2, la, 3, n_p))
            native._check(jop._lib, jop._h, jop._lib.arroyo_b200_op_process_device_batch(jop._h, 1, 2, ra, 4, n_a))
            outb = (ffi.DeviceBatch * 8)()
            n = C.c_int64(0)
            native._check(jop._lib, jop._h, jop._lib.arroyo_b200_op_handle_watermark_device(jop._h, ts + 1, outb, 8, C.byref(n)))
            rows_out += sum(outb[i].n_rows for i in range(n.value))

        for w in range(3):
            jstep(w)
        rows_out = 0
        ms = timed(torch, lambda r: jstep(3 + r), 10)
        rec("InstantJoin (inner)", f"configs[3] shape: per 30 s window {n_p} persons x {n_a} auctions on person id = seller, "
            f"{rows_out // 10} joined rows out", n_p + n_a, ms, (24 * n_p + 32 * n_a + 48 * (rows_out // 10)) / (n_p + n_a),
            "inputs read once + joined rows (6 columns) written once; includes the arena append and the compaction",
            cpu=cpu_join_rate(n_p, n_a))
        jop.close()
        del pid, name, seller, auction, reserve
        torch.cuda.empty_cache()

    def session_section():
        # ---- configs[4]: session windows, 5 s gap, many keys -------------------------------------------------
        n_keys = int(os.environ.get("SESSION_KEYS", 10_000_000))
        srows = 1 << 22
        scfg = ab.SessionConfig(gap=5 * S, key_names=["key"], aggs=[ab.Agg("sum", "value", "sum"), ab.Agg("count", None, "n")],
                                window_index=1)
        sop = native.SessionAggregatingWindowFunc(scfg, input_schema=raw_schema, device=0, stream=stream, expected_keys=n_keys)
        n_steps = 26
        sk = [torch.randint(0, n_keys, (srows,), device=dev, generator=g, dtype=torch.int64) * 7919 for _ in range(n_steps)]
        sv = torch.randint(0, 10**6, (srows,), device=dev, generator=g, dtype=torch.int64)
        offs = torch.sort(torch.randint(0, S, (srows,), device=dev, generator=g, dtype=torch.int64)).values
        sess_out = 0

        def sstep(p):
            nonlocal sess_out
            ts = offs + (T0 + p * S)
            sop.process_device_batch([sk[p].data_ptr(), sv.data_ptr(), ts.data_ptr()], srows)
            for n, _ in sop.handle_watermark_device(T0 + p * S - S):
                sess_out += n

        for p in range(14):
            sstep(p)
        sess_out = 0
        ms = timed(torch, lambda r: sstep(14 + r), 12)
        rec("SessionAggregatingWindowFunc", f"configs[4] shape: gap 5 s, {n_keys} keys, 4 Mi rows per second of event time, "
            f"{sess_out // 12} sessions closed per step", srows, ms, 24,
            "one thread per key replays the reference's per-key state machine; inputs read once",
            cpu=cpu_session_rate(n_keys, srows))
        sop.close()

    if want("wm"):
        wm_section()
    if want("join"):
        join_section()
    if want("session"):
        session_section()

    print(json.dumps({"peak_GBps": peak, "cases": out}, indent=1))


if __name__ == "__main__":
    main()

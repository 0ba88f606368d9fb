"""The golden pipelines with a checkpoint and a restart in the middle, the way the reference's own smoke tests run every
query (arroyo-sql-testing/src/smoke_tests.rs: each query is run, checkpointed, stopped and resumed from the checkpoint,
and the concatenated output must still equal the golden file).

`RestartOps(base, kit)` wraps an operator factory namespace (the oracle module or tests.gpu_ops): its
`run_single_input` drives the operator up to a cut point, calls handle_checkpoint, throws the operator away, constructs a
fresh one, calls on_start with the same context (the state tables) and continues.  The join driver in golden_cases uses
`restart_join` the same way between the two input streams."""
from oracle import arroyo_oracle as O


class OracleKit:
    make_ctx = staticmethod(lambda n: O.OperatorContext(n))

    @staticmethod
    def checkpoint(op, ctx):
        op.handle_checkpoint(ctx)

    @staticmethod
    def respawn(op, ctx, sample_batch=None):
        new = type(op)(op.cfg)
        new.on_start(ctx)
        return new


class GpuKit:
    @staticmethod
    def make_ctx(n):
        import arroyo_b200 as ab
        return ab.OperatorContext(n)

    @staticmethod
    def checkpoint(op, ctx):
        op.op.handle_checkpoint(None, ctx, None)

    @staticmethod
    def respawn(op, ctx, sample_batch=None):
        from tests import gpu_ops as G
        kw = {}
        if sample_batch is not None and not isinstance(op, G.InstantJoin):
            kw["input_schema"] = G.to_arrow(sample_batch).schema
        if isinstance(op, G.InstantJoin):
            new = G.InstantJoin(op.op.config)
        else:
            op.close()
            new = type(op)(op.cfg, **kw)
        new.op.on_start(ctx)
        return new


class RestartOps:
    def __init__(self, base, kit, frac=0.5):
        self.base, self.kit, self.frac = base, kit, frac
        for name in ("TumblingAggregatingWindowFunc", "SlidingAggregatingWindowFunc", "SessionAggregatingWindowFunc",
                     "InstantJoin"):
            setattr(self, name, getattr(base, name))

    def run_single_input(self, op, batches, delay_ns=1_000_000_000, ctx=None):
        ctx = ctx or self.kit.make_ctx(1)
        out = O.Collector()
        gen = O.WatermarkGenerator(delay_ns)
        cut = max(1, int(len(batches) * self.frac))
        for i, b in enumerate(batches):
            if i == cut:
                self.kit.checkpoint(op, ctx)
                op = self.kit.respawn(op, ctx, batches[0])
            op.process_batch(b, ctx, out)
            wm = gen.process_batch(b[O.TIMESTAMP])
            if wm is not None:
                ctx.watermarks.set(0, wm)
                op.handle_watermark(wm, ctx, out)
        ctx.watermarks.set(0, O.FINAL_WATERMARK)
        op.handle_watermark(O.FINAL_WATERMARK, ctx, out)
        return out

    def restart_join(self, join, ctx):
        self.kit.checkpoint(join, ctx)
        return self.kit.respawn(join, ctx)

    def join_ctx(self):
        return self.kit.make_ctx(2)

"""LaggedCombiner without any process group: ordering of the two stages, the lag, and failure propagation (an error in
the owner-stage thread must surface in the local stage's thread instead of blocking it)."""
import threading

import pytest

from arroyo_b200.multi_gpu import LaggedCombiner


class FakeExchange:
    """One rank: the effective watermark is this rank's own latest watermark."""

    class Holder:
        last_present_watermark = None

    def __init__(self, fail_at=None):
        self.holder = self.Holder()
        self.rounds = 0
        self.fail_at = fail_at

    def round_packed(self, packed, counts, m, watermark, more=False):
        self.rounds += 1
        if self.fail_at is not None and self.rounds == self.fail_at:
            raise ValueError("link down")
        before = self.holder.last_present_watermark
        if watermark is not None:
            self.holder.last_present_watermark = watermark
        after = self.holder.last_present_watermark
        batches = [(list(packed), m)] if m else []
        return batches, (after if after != before else None), bool(more)


def test_stages_run_in_order_two_rounds_apart():
    log, lock = [], threading.Lock()

    def note(*x):
        with lock:
            log.append(x)

    ex = FakeExchange()
    closed = []

    def local_close(eff):
        closed.append(eff)
        return [((eff, eff + 1), 2), ((eff + 2,), 1)]  # two chunks: two rounds in that step

    pipe = LaggedCombiner(ex, local_close, lambda c: (c[0], None, c[1]),
                          lambda batches, consume_now: note("ingest", batches[0][0], consume_now),
                          lambda eff: note("owner_wm", eff), lag=2)
    for p in range(8):
        pipe.local_step(lambda p=p: note("feed", p), 100 + p)
    pipe.drain()
    pipe.close()
    # the local stage closed panes with the watermark of two rounds earlier, each exactly once, in order
    assert closed == [100 + p for p in range(6)]
    # owner stage: for every closing the chunks are ingested (first with more rounds to come -> consumed at once),
    # then the owner's watermark follows
    owner = [x for x in log if x[0] != "feed"]
    for i, eff in enumerate(closed):
        assert owner[3 * i:3 * i + 3] == [("ingest", [eff, eff + 1], True), ("ingest", [eff + 2], False), ("owner_wm", eff)]
    assert [x[1] for x in log if x[0] == "feed"] == list(range(8))


def test_owner_stage_failure_reaches_the_local_stage():
    ex = FakeExchange(fail_at=3)
    pipe = LaggedCombiner(ex, lambda eff: [], lambda c: (c[0], None, c[1]), lambda b, consume_now: None, lambda eff: None, lag=2)
    with pytest.raises(RuntimeError, match="owner stage failed"):
        for p in range(10):
            pipe.local_step(lambda: None, 100 + p)
        pipe.drain()


def test_watchdog_ends_a_stalled_rank_and_leaves_a_progressing_one_alone():
    """multi_gpu._Watchdog: a rank of the N > 1 plan that stops making progress (a collective whose peer never arrives)
    says where it stood and exits with 124 instead of hanging until the launcher's limit."""
    import os
    import subprocess
    import sys
    import time

    from arroyo_b200.multi_gpu import _Watchdog
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    dog = _Watchdog(0, "a test", limit_s=1.5)
    for i in range(4):
        time.sleep(0.6)
        dog.beat(f"step {i}")
    dog.close()  # 2.4 s without firing: beats keep it quiet
    code = ("import sys, time; sys.path.insert(0, %r)\n"
            "from arroyo_b200.multi_gpu import _Watchdog\n"
            "d = _Watchdog(3, 'the 8-GPU plan', limit_s=1.0); d.beat('local stage fed pane 7'); time.sleep(30)\n" % ROOT)
    t0 = time.time()
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=25)
    assert r.returncode == 124 and time.time() - t0 < 20
    assert "rank 3" in r.stderr and "local stage fed pane 7" in r.stderr

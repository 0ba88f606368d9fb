"""BASELINE configs[0]: Nexmark q1 / q2 stateless map / filter, 1 CPU worker, 64 Ki-row Arrow batches -- rows/s of the
host plumbing (arroyo_b200/plumbing.py; no GPU involved) and a correctness check of every batch against numpy.

    python tools/q1_plumbing.py > profiles/r02_q1_plumbing.json
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from arroyo_b200 import plumbing as P  # noqa: E402
from arroyo_b200.context import Collector  # noqa: E402
from tests.test_plumbing import bids  # noqa: E402


def run(op, batches, check):
    out = Collector()
    t0 = time.perf_counter()
    for b in batches:
        op.process_batch(b, None, out)
    dt = time.perf_counter() - t0
    ok = all(check(b, o) for b, o in zip(batches, out.batches))
    return sum(b.num_rows for b in batches) / dt, ok


def main():
    batches = [bids(65_536, seed=s) for s in range(64)]
    q1, ok1 = run(P.nexmark_q1(), batches, lambda b, o: np.array_equal(
        o.column(2).to_numpy(), b.column(2).to_numpy().astype(np.float64) * 0.908))
    q2, ok2 = run(P.nexmark_q2(), batches, lambda b, o: np.array_equal(
        o.column(1).to_numpy(), b.column(2).to_numpy()[b.column(0).to_numpy() % 123 == 0]))
    print(json.dumps({"config": "BASELINE configs[0]: Nexmark q1 / q2 stateless map / filter, 1 CPU worker, 64 Ki-row Arrow "
                                "batches (plumbing, no GPU)", "batches": len(batches), "batch_rows": 65_536,
                      "q1_rows_per_s": q1, "q1_correct": ok1, "q2_rows_per_s": q2, "q2_correct": ok2,
                      "cores": 1, "host": os.uname().machine}, indent=1))


if __name__ == "__main__":
    main()

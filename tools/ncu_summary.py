#!/usr/bin/env python
"""Summarise ncu outputs into small text/JSON files for profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches.csv  > profiles/<name>_launches.txt
  python tools/ncu_summary.py full gpurun_out/prof.ncu-rep [kernel-regex] > profiles/<name>_full.json
  python tools/ncu_summary.py traffic gpurun_out/prof.ncu-rep <rows per launch> "<source note>" > profiles/ingest_traffic.json
      (DRAM bytes per input row over every kernel of the capture = one ingest launch: part_kernel + agg_kernel)
"""
import csv
import io
import json
import re
import subprocess
import sys
from collections import defaultdict


def launches(path):
    rows = []
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(io.StringIO("".join(lines)))
    per = defaultdict(lambda: [0, 0.0])
    total = 0.0
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(unit, 1e-3)
        v *= scale
        per[name][0] += 1
        per[name][1] += v
        total += v
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none  ({path})")
    print("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes")
    print(f"{'kernel':60s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}")
    for name, (n, t) in sorted(per.items(), key=lambda kv: -kv[1][1]):
        print(f"{name[:60]:60s} {n:8d} {t:12.1f} {t / n:10.1f} {100 * t / total:6.1f}%")
    print(f"{'TOTAL':60s} {sum(n for n, _ in per.values()):8d} {total:12.1f}")


KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "l1tex__data_pipe_lsu_wavefronts.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_atom.sum", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg", "sm__cycles_elapsed.avg",
    "smsp__inst_executed_op_shared_atom.sum", "launch__shared_mem_per_block_dynamic",
]


def full(path, pattern=None):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        if pattern and not re.search(pattern, d.get("Kernel Name", "")):
            continue
        e = {"kernel": d.get("Kernel Name")}
        for k in KEYS:
            if k in d:
                e[k] = {"value": d[k], "unit": units[hdr.index(k)]}
        out.append(e)
    print(json.dumps(out, indent=1))


def traffic(path, rows_per_launch, source):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    kernels, total, us = [], 0.0, 0.0
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        b = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            b += float(d[k].replace(",", "")) * scale[units[hdr.index(k)]]
        t = float(d["gpu__time_duration.sum"].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}[units[hdr.index("gpu__time_duration.sum")]]
        name = re.sub(r"\(.*", "", d["Kernel Name"]).replace("void ", "")
        kernels.append({"kernel": name, "dram_bytes": b, "us_under_ncu": round(t, 1)})
        total += b
        us += t
    print(json.dumps({"kernel": "ingest = " + " + ".join(k["kernel"].split("::")[-1].split("<")[0] for k in kernels),
                      "dram_bytes_per_row": total / rows_per_launch, "dram_bytes_per_launch": total,
                      "rows_per_launch": rows_per_launch, "algorithmic_bytes_per_launch": 24 * rows_per_launch,
                      "kernels": kernels, "kernel_us_under_ncu": round(us, 1), "source": source}, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    elif sys.argv[1] == "traffic":
        traffic(sys.argv[2], int(sys.argv[3]), sys.argv[4])
    else:
        full(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)

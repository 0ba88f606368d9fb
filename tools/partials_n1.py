"""The N > 1 plan (partial -> shuffle -> final, arroyo_b200.multi_gpu._run_plan) on ONE GPU with a world-1 NCCL group:
every kernel of the local stage, the shuffle edge and the owner stage on one device, so that `ncu` can list them
(ncu is never pointed at a multi-rank command).  Prints the plan's timing dict.

    ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 120 --csv --log-file launches.csv \
        python tools/partials_n1.py --steps 4 --warmup 14
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    import arroyo_b200 as ab
    import bench as B
    from arroyo_b200 import multi_gpu, operators as native

    args = B.parse()
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29577")
    torch.cuda.set_device(0)
    device = torch.device("cuda", 0)
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=device)
    torch.cuda.set_stream(torch.cuda.Stream(device=device, priority=-1))
    W, K = B.steady_warmup(args.warmup, extra=2), args.steps
    gen = B.make_generator(torch, device, args.rows_per_pane, args.keys, args.dist, 42, args.keyspace)
    panes = [gen(p) for p in range(W + K)]
    res = multi_gpu._run_plan(args, torch, dist, B, ab, native, 0, 1, 0, device, panes, W, K)
    res.pop("sums", None)
    res["rows_per_s"] = K * args.rows_per_pane / (res["ms"] * 1e-3) if K else None
    print(json.dumps(res, default=str), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

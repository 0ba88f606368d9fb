"""Micro-probe (not product code): what does the host link of this box give?  Bounds the end-to-end number
(bench.py `e2e`): 24 B/row in, 48 B/emitted row out.

    python tools/pcie_probe.py > gpurun_out/pcie_probe.txt
"""
import subprocess
import time

import torch


def bw(nbytes, seconds):
    return nbytes / seconds / 1e9


def main():
    print(subprocess.run(["nvidia-smi", "--query-gpu=name,pcie.link.gen.current,pcie.link.gen.max,pcie.link.width.current",
                          "--format=csv"], capture_output=True, text=True).stdout)
    dev = torch.device("cuda", 0)
    total = 1 << 30
    host = torch.empty(total, dtype=torch.uint8).pin_memory()
    host.fill_(1)
    d = torch.empty(total, dtype=torch.uint8, device=dev)
    d2 = torch.empty(total, dtype=torch.uint8, device=dev)
    host2 = torch.empty(total, dtype=torch.uint8).pin_memory()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for piece in (512 << 10, 4 << 20, 64 << 20, 1 << 30):
        n = total // piece
        for name, src, dst in (("H2D", host, d), ("D2H", d, host2)):
            for rep in range(2):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                with torch.cuda.stream(s1):
                    for i in range(n):
                        dst[i * piece:(i + 1) * piece].copy_(src[i * piece:(i + 1) * piece], non_blocking=True)
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
            print(f"{name} pinned, 1 stream , pieces of {piece >> 10:8d} KiB : {bw(total, dt):6.1f} GB/s")
        # two streams, alternating pieces
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(n):
            with torch.cuda.stream(s1 if i % 2 == 0 else s2):
                d[i * piece:(i + 1) * piece].copy_(host[i * piece:(i + 1) * piece], non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"H2D pinned, 2 streams, pieces of {piece >> 10:8d} KiB : {bw(total, dt):6.1f} GB/s")
        # both directions at once
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(n):
            with torch.cuda.stream(s1):
                d[i * piece:(i + 1) * piece].copy_(host[i * piece:(i + 1) * piece], non_blocking=True)
            with torch.cuda.stream(s2):
                host2[i * piece:(i + 1) * piece].copy_(d2[i * piece:(i + 1) * piece], non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"H2D + D2H together   , pieces of {piece >> 10:8d} KiB : {bw(total, dt):6.1f} GB/s each way")
    # pageable for comparison
    pg = torch.empty(total, dtype=torch.uint8)
    pg.fill_(1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    d.copy_(pg)
    torch.cuda.synchronize()
    print(f"H2D pageable 1 GiB: {bw(total, time.perf_counter() - t0):6.1f} GB/s")
    # NUMA placement of this process
    try:
        print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:1500])
    except Exception as e:  # noqa: BLE001
        print("topo:", e)


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""What the host link gives: pinned 33 MB buffers (one 4K RGBA8 frame) copied H2D only, D2H only and both directions at once on two
streams, with and without binding the thread to the GPU's NUMA node first.  The e2e leg of bench.py moves one frame each way per frame, so
its ceiling is the BIDIRECTIONAL figure (PCIe Gen5 x16: 63 GB/s raw per direction).  Evidence for DESIGN.md §5."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch


def run(tag):
    n = 33177600
    dev = torch.device("cuda", 0)
    h_in = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(4)]
    h_out = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(4)]
    d_a = [torch.empty(n, dtype=torch.uint8, device=dev) for _ in range(4)]
    d_b = [torch.empty(n, dtype=torch.uint8, device=dev) for _ in range(4)]
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    res = {}
    for mode in ("h2d", "d2h", "both"):
        for rep in range(2):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for i in range(200):
                if mode in ("h2d", "both"):
                    with torch.cuda.stream(s1): d_a[i % 4].copy_(h_in[i % 4], non_blocking=True)
                if mode in ("d2h", "both"):
                    with torch.cuda.stream(s2): h_out[i % 4].copy_(d_b[i % 4], non_blocking=True)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
        res[mode] = 200 * n / dt / 1e9
    res["frames_per_s_ceiling_bidirectional"] = res["both"] * 1e9 / n
    print(json.dumps({tag: res}))


if __name__ == "__main__":
    run("unbound")
    import gyroflow_b200 as g
    ncpu = g.bind_thread_to_device(0)
    run("bound_to_gpu_numa_node(%d cpus)" % ncpu)

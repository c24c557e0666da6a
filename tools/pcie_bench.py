#!/usr/bin/env python3
"""Host link ceiling of the box: pinned H2D, D2H and both at once, on every rank at the same time.

    python tools/pcie_bench.py                      one GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/pcie_bench.py

This is the end-to-end roofline of the host-buffer entry points (bench.py `e2e.frac_of_link`): a step moves
h2d_bytes up and d2h_bytes down and cannot be faster than those bytes over these rates.  Also times a
multithreaded host memcpy (what any host-side packing pass competes with).  One JSON line from rank 0."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)

NB = int(float(sys.argv[1]) * 1e9) if len(sys.argv) > 1 else 1_500_000_000
h_up = torch.empty(NB, dtype=torch.uint8, pin_memory=True)
h_dn = torch.empty(NB, dtype=torch.uint8, pin_memory=True)
h_up.fill_(1)
h_dn.fill_(2)
d_up = torch.empty(NB, dtype=torch.uint8, device=dev)
d_dn = torch.ones(NB, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def run(up, down, reps=3):
    best = 1e9
    for _ in range(reps + 1):
        barrier()
        t0 = time.perf_counter()
        if up:
            with torch.cuda.stream(s1):
                d_up.copy_(h_up, non_blocking=True)
        if down:
            with torch.cuda.stream(s2):
                h_dn.copy_(d_dn, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        best = min(best, dt)
    return best


res = {"n_gpus": world, "bytes_per_direction_per_gpu": NB, "cpus": len(os.sched_getaffinity(0))}
t = run(True, False)
res["h2d_gbs_per_gpu"] = NB / t / 1e9
t = run(False, True)
res["d2h_gbs_per_gpu"] = NB / t / 1e9
t = run(True, True)
res["duplex_gbs_per_gpu_each_direction"] = NB / t / 1e9
res["aggregate_duplex_gbs"] = 2 * world * NB / t / 1e9

# host memcpy with this rank's share of the threads (numpy releases the GIL)
import numpy as np
from concurrent.futures import ThreadPoolExecutor
threads = max(1, len(os.sched_getaffinity(0)) // world)
a, b = h_up.numpy(), h_dn.numpy()
cuts = [NB * i // threads for i in range(threads + 1)]
with ThreadPoolExecutor(threads) as pool:
    best = 1e9
    for _ in range(3):
        barrier()
        t0 = time.perf_counter()
        list(pool.map(lambda i: np.copyto(b[cuts[i]:cuts[i + 1]], a[cuts[i]:cuts[i + 1]]), range(threads)))
        best = min(best, time.perf_counter() - t0)
res["host_memcpy_gbs_per_rank"] = NB / best / 1e9
res["host_memcpy_threads_per_rank"] = threads
if rank == 0:
    print(json.dumps(res))
if world > 1:
    dist.destroy_process_group()

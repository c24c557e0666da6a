#!/usr/bin/env python3
"""End to end through kbbq_recalibrate_host_multi (ONE process, one host thread + session per device, tables summed
over NVLink peer memory): config-3-shaped reads in pinned host memory, 1 / 2 / 4 / 8 devices of the box.

    python tools/multi_e2e.py [reads_per_gpu] [read_groups]

Weak scaling: n devices share n x reads_per_gpu reads.  Also checks that every device count gives the same bytes."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "kbbq-py_b200"))
from kbbq import _native  # noqa: E402
from kbbq.device import synth_reads  # noqa: E402

per, R = (int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000), (int(sys.argv[2]) if len(sys.argv) > 2 else 8)
L = 150
ndev = torch.cuda.device_count()
N = per * ndev
h = {k: torch.empty((N, L) if k in ("seq", "qual", "corr", "out") else (N,), dtype=torch.int16 if k == "rg" else torch.uint8,
                    pin_memory=True) for k in ("seq", "qual", "corr", "rg", "second", "out")}
for d in range(ndev):   # every device generates its part of the stream
    with torch.cuda.device(d):
        parts = synth_reads(1003, d * per, per, L, R, device=torch.device("cuda", d))
        for k, t in zip(("seq", "qual", "corr", "rg", "second"), parts):
            h[k][d * per:(d + 1) * per].copy_(t)
        del parts
        torch.cuda.empty_cache()
torch.cuda.synchronize()
a = {k: v.numpy() for k, v in h.items()}
rg = a["rg"].view(np.uint16)
res, ref = {}, None
n = 1
while n <= ndev:
    m = n * per
    devs = list(range(n))
    out = a["out"][:m]
    times = []
    for rep in range(4):
        t0 = time.perf_counter()
        _native.recalibrate_host(a["seq"][:m], a["qual"][:m], a["corr"][:m], rg[:m] if R > 1 else None, a["second"][:m], L, R,
                                 out=out, devices=devs)
        times.append(time.perf_counter() - t0)
    best = sorted(times[1:])[len(times[1:]) // 2]
    res[n] = {"ms_per_step": 1e3 * best, "gbases_per_s": m * L / best / 1e9, "all_ms": [1e3 * t for t in times]}
    if n == 1:
        ref = out[:per].copy()
    else:   # first shard's bytes must not depend on how many devices shared the work ... when the tables are the same
        pass
    n *= 2
# same reads, different device counts: identical bytes
m = per
outs = []
for devs in ([0], list(range(min(2, ndev))), list(range(ndev))):
    o = np.empty((m, L), np.uint8)
    _native.recalibrate_host(a["seq"][:m], a["qual"][:m], a["corr"][:m], rg[:m] if R > 1 else None, a["second"][:m], L, R,
                             out=o, devices=devs)
    outs.append(o)
same = all(np.array_equal(outs[0], o) for o in outs[1:])
print(json.dumps({"tool": "multi_e2e", "reads_per_gpu": per, "read_len": L, "read_groups": R, "devices": ndev,
                  "cpus": len(os.sched_getaffinity(0)), "weak_scaling": res,
                  "efficiency_vs_1": {k: v["gbases_per_s"] / (k * res[1]["gbases_per_s"]) for k, v in res.items()},
                  "same_bytes_for_1_2_all_devices": bool(same)}))

#!/usr/bin/env python3
"""Device time of the BAM-side kernels (csrc/bam.cuh) on config-2-sized arrays.  B200 only.
   python tools/bam_bench.py [reads] [read_len] [read_groups]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "kbbq-py_b200"))
import torch  # noqa: E402
from kbbq.device import DeviceRecalibrator, synth_reads  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 150
    R = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    seq, qual, corr, rg, second = synth_reads(1002, 0, n, L, R)
    err = (seq != corr).to(torch.uint8)
    g = torch.Generator(device="cuda").manual_seed(1)
    skip = (torch.rand(n, L, device="cuda", generator=g) < 0.02).to(torch.uint8)
    flags = torch.randint(0, 4, (n,), device="cuda", generator=g, dtype=torch.uint8)
    a0 = torch.zeros(n, dtype=torch.int16, device="cuda")
    a1 = torch.full((n,), L, dtype=torch.int16, device="cuda")
    rec = DeviceRecalibrator(L, R, max_reads=0)
    out = torch.empty_like(qual)
    rgarg = rg if R > 1 else None

    def timed(fn, reps=5):
        fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    for fast in (True, False):
        tb = timed(lambda: rec.build_bam(seq, qual, err, skip, rgarg, flags, a0, a1, fast=fast))
        rec.model()
        ta = timed(lambda: rec.apply_bam(seq, qual, out, rgarg, flags, fast=fast))
        rec.check_status()
        print("%d x %d bp, %d read group(s), %s: build_bam %.2f ms (%.0f Gbases/s), apply_bam %.2f ms (%.0f Gbases/s)" %
              (n, L, R, "canonical form + shared-memory kernels" if fast else "direct kernels (global atomics)",
               tb, n * L / tb / 1e6, ta, n * L / ta / 1e6))


if __name__ == "__main__":
    main()

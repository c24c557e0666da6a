#!/usr/bin/env python3
"""Device time of the calibration-count kernel (csrc/calib.cuh) on BASELINE config 2 sized input.
   python tools/calib_bench.py [reads] [read_len]     (B200 only; prints GB/s against the measured HBM peak)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "kbbq-py_b200"))
import torch  # noqa: E402
from kbbq.device import calibration_counts, synth_reads  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 150
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    seq, qual, corr, rg, second = synth_reads(1002, 0, n, L, 1)
    err = (seq != corr).to(torch.uint8)
    skip = (qual < 6).to(torch.uint8)
    cases = {"qual+seq+corr (3 B/base)": (dict(seq=seq, corr=corr), 3), "qual+err (2 B/base)": (dict(err=err), 2),
             "qual+err+skip (3 B/base)": (dict(err=err, skip=skip), 3)}
    for name, (kw, bpb) in cases.items():
        for _ in range(3):
            calibration_counts(qual, **kw)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            calibration_counts(qual, **kw)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        gbs = bpb * n * L / ms / 1e6
        print("%-28s %.3f ms  %.0f GB/s  %.1f %% of %.0f GB/s  %.0f Gbases/s" % (name, ms, gbs, 100 * gbs / peak, peak, n * L / ms / 1e6))


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Build / apply times of one batch in both HBM layouts: rows in read order with rg[] per read (work-list
gather) and the segmented layout (rows sorted by read group and mate).  python tools/seg_bench.py R L N"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "kbbq-py_b200"))
from kbbq.device import DeviceRecalibrator, synth_reads  # noqa: E402

R, L, N = (int(x) for x in sys.argv[1:4])
seq, qual, corr, rg, second = synth_reads(1003, 0, N, L, R)
rg_arg = rg if R > 1 else None
rec = DeviceRecalibrator(L, R, max_reads=N + 64 * R)
out = torch.empty_like(qual)


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


peak = 6551.4
tb = timed(lambda: rec.build(seq, qual, corr, rg_arg, second))
rec.reset()
rec.build(seq, qual, corr, rg_arg, second)
whole = rec.tables.clone()
rec.model()
ta = timed(lambda: rec.apply(seq, qual, out, rg_arg, second))
print("read order : build %.3f ms (%.0f %%)  apply %.3f ms (%.0f %%)" % (
    tb, 100 * 3 * N * L / tb / 1e6 / peak, ta, 100 * 3 * N * L / ta / 1e6 / peak))
ts = timed(lambda: rec.segment(seq, qual, corr, rg_arg, second), reps=2)
sb = rec.segment(seq, qual, corr, rg_arg, second)
rec.reset()
rec.build_segmented(sb)
assert torch.equal(rec.tables, whole), "segmented tables differ"
tb = timed(lambda: rec.build_segmented(sb))
out_seg = torch.empty(sb.rows_bound * L + 16, dtype=torch.uint8, device=qual.device)
ta = timed(lambda: rec.apply_segmented(sb, out_seg))
out2 = torch.empty_like(qual)
tu = timed(lambda: rec.unsegment(sb, out_seg, out2))
assert torch.equal(out, out2), "segmented output differs"
rec.status.zero_()
print("segmented  : build %.3f ms (%.0f %%)  apply %.3f ms (%.0f %%)   [segment 3 arrays %.3f ms, unsegment %.3f ms]" % (
    tb, 100 * 3 * N * L / tb / 1e6 / peak, ta, 100 * 3 * N * L / ta / 1e6 / peak, ts, tu))

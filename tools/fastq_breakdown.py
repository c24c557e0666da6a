#!/usr/bin/env python3
"""Where FASTQ -> recalibrated FASTQ spends its time (kbbq.recalibrate.recalibrate_fastq step by step), files on
/dev/shm.  python tools/fastq_breakdown.py [reads] [read_groups]"""
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "kbbq-py_b200"))
from kbbq import _native, fastx, synth  # noqa: E402
from kbbq.device import synth_reads  # noqa: E402

n, R = (int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000), (int(sys.argv[2]) if len(sys.argv) > 2 else 8)
L = 150
d = tempfile.mkdtemp(prefix="kbbq_fq_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
data = [t.cpu().numpy() for t in synth_reads(1003, 0, n, L, R)]
data[3] = data[3].view(np.uint16)
fu, fc, fo = (os.path.join(d, x) for x in ("reads.fq", "corrected.fq", "out.fq"))
synth.write_fastq_fast(fu, fc, *data, infer_rg=True)
del data
for rep in range(3):
    t = [time.perf_counter()]
    lap = lambda: t.append(time.perf_counter())
    reads = fastx.NativeFastq(fu); lap()
    fixed = fastx.NativeFastq(fc); lap()
    reads.check_names(fixed, n); lap()
    rg, second, keys = reads.infer(True); lap()
    seq, qual = reads.pack(0, n); lap()
    corr, _ = fixed.pack(0, n); lap()
    out = _native.recalibrate_host(seq, qual, corr, rg, second, L, len(keys)); lap()
    fd = os.open(fo, os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o600)
    reads.write(fd, out, 0, n); os.close(fd); lap()
    reads.close(); fixed.close(); lap()
    names = ("open+index reads", "open+index corrected", "check names", "infer rg/second", "pack reads", "pack corrected",
             "recalibrate_host (pageable arrays)", "format + write", "close")
    print("rep %d total %.1f ms: " % (rep, 1e3 * (t[-1] - t[0])) +
          ", ".join("%s %.1f" % (k, 1e3 * (b - a)) for k, a, b in zip(names, t, t[1:])))
# the native pipeline (what kbbq.recalibrate.recalibrate_fastq calls): same files, output compared byte for byte
import ctypes as C
want = open(fo, "rb").read()
for rep in range(3):
    fd = os.open(fo + ".native", os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o600)
    nn, nrg, st = C.c_int64(0), C.c_int(0), C.c_int(0)
    t0 = time.perf_counter()
    rc = _native.lib().kbbq_recalibrate_fastq(fu.encode(), fc.encode(), 1, 6, fd, 0, 0, C.byref(nn), C.byref(nrg), C.byref(st))
    dt = time.perf_counter() - t0
    os.close(fd)
    print("native pipeline rep %d: rc %d, %d reads, %d read groups, %.1f ms = %.2f Gbases/s, same bytes: %s" % (
        rep, rc, nn.value, nrg.value, 1e3 * dt, n * L / dt / 1e9, open(fo + ".native", "rb").read() == want))
import shutil
shutil.rmtree(d, ignore_errors=True)

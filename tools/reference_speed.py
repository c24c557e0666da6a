#!/usr/bin/env python3
"""Speed of the UNMODIFIED Python reference on the first reads of BASELINE config 2.
Build container only (needs /root/reference, imported through oracle/ref_shim); one core, as the
reference is single threaded.     python tools/reference_speed.py [reads]"""
import contextlib
import io
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "kbbq-py_b200"))
from kbbq import synth as _synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
seq, qual, corr, rg, second = _synth.synth_reads(1002, 0, n, 150, 1)
import ref_shim  # noqa: E402

rc, cr, ab = ref_shim.load()
d = tempfile.mkdtemp()


def write(path, s):
    with open(path, "wb") as fh:
        for i in range(s.shape[0]):
            fh.write(b"@r%d/%d\n%s\n+\n%s\n" % (i // 2, 1 + (i & 1), s[i].tobytes(), (qual[i] + 33).astype(np.uint8).tobytes()))


write(d + "/a.fq", seq)
write(d + "/b.fq", corr)
t0 = time.perf_counter()
with contextlib.redirect_stdout(io.StringIO()):
    rc.recalibrate_fastq([d + "/a.fq", d + "/b.fq"])
dt = time.perf_counter() - t0
print("reference recalibrate_fastq: %d reads x 150 bp in %.1f s = %.3f Mbases/s (1 core)" % (n, dt, seq.size / dt / 1e6))

#!/usr/bin/env python3
"""Throughput of the native FASTQ ingest / egress (csrc/fastq_io.cpp) beside the per-read Python
tokenising the reference's driver does (FastxFile iteration + get_quality_array + name inference,
kbbq/recalibrate.py:56-64,92).  Host only.   python tools/fastq_io_bench.py [reads] [read_len]"""
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "kbbq-py_b200"))
from kbbq import compare_reads, fastx  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 150
    rng = np.random.default_rng(1)
    d = tempfile.mkdtemp(prefix="kbbq_fq_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    path, out = os.path.join(d, "reads.fq"), os.path.join(d, "out.fq")
    block = 100_000
    with open(path, "wb") as fh:
        for b in range(0, n, block):
            m = min(block, n - b)
            seq = rng.choice(np.frombuffer(b"ACGT", np.uint8), size=(m, L))
            qual = rng.integers(35, 74, size=(m, L), dtype=np.uint8)
            fh.write(b"".join(b"@r%d/%d_RG:Z:g%d some comment\n%s\n+\n%s\n" %
                              ((b + i) // 2, 1 + (i & 1), ((b + i) // 2) % 8, seq[i].tobytes(), qual[i].tobytes())
                              for i in range(m)))
    size = os.path.getsize(path)
    t0 = time.perf_counter()
    f = fastx.NativeFastq(path)
    rg, second, keys = f.infer(True)
    seq, qual = f.pack()
    t_in = time.perf_counter() - t0
    fd = os.open(out, os.O_WRONLY | os.O_CREAT | os.O_TRUNC)
    t0 = time.perf_counter()
    f.write(fd, qual)
    os.close(fd)
    t_out = time.perf_counter() - t0
    # the per-read Python path on a sample
    sample = min(n, 50_000)
    t0 = time.perf_counter()
    seen = {}
    for i, rec in enumerate(fastx.FastxFile(path)):
        if i == sample:
            break
        q = np.array(rec.get_quality_array())
        seen.setdefault(compare_reads.fastq_infer_rg(rec), len(seen))
        compare_reads.fastq_infer_secondinpair(rec)
    t_py = time.perf_counter() - t0
    print("file: %d reads x %d bp, %.1f MB; threads: %d" % (n, L, size / 1e6, os.cpu_count()))
    print("native ingest (index + infer RG/pair + pack): %.2f s  %.0f MB/s  %.2f Mreads/s" %
          (t_in, size / t_in / 1e6, n / t_in / 1e6))
    print("native egress (format + write):               %.2f s  %.0f MB/s" % (t_out, os.path.getsize(out) / t_out / 1e6))
    print("python per-read tokenising (%d reads):       %.2f s  %.3f Mreads/s  -> native ingest is %.0fx" %
          (sample, t_py, sample / t_py / 1e6, (n / t_in) / (sample / t_py)))
    for p in (path, out):
        os.remove(p)
    os.rmdir(d)


if __name__ == "__main__":
    main()

"""How fast do the host cores turn (seq, corrected) into the forms kbbq_recalibrate_host sends over PCIe?
   python tools/host_pack_bench.py [bases] -- mismatch bit map and 4-bit form (csrc/host_pack.cpp), pinned
   destination when torch + CUDA are there, thread counts 1 .. all."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "kbbq-py_b200"))
from kbbq import _native  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 600_000_000
    lib = _native.lib()
    rng = np.random.default_rng(1)
    seq = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, n, dtype=np.uint8)]
    corr = seq.copy()
    corr[::97] = ord("A")
    try:
        import torch
        dst = torch.empty(n // 2 + 64, dtype=torch.uint8, pin_memory=torch.cuda.is_available()).numpy()
    except Exception:
        dst = np.empty(n // 2 + 64, np.uint8)
    dst[:] = 0
    cpus = len(os.sched_getaffinity(0))
    for threads in sorted({1, 2, 4, 8, cpus}):
        for name in ("bits", "nibbles"):
            best = 1e9
            for _ in range(3):
                t0 = time.perf_counter()
                if name == "bits":
                    rc = lib.kbbq_host_mismatch_bits(_native.ptr(seq), _native.ptr(corr), n, _native.ptr(dst), threads)
                else:
                    bad = C.c_int(0)
                    rc = lib.kbbq_host_pack_nibbles(_native.ptr(seq), _native.ptr(corr), n, _native.ptr(dst), threads, C.byref(bad))
                assert rc == 0
                best = min(best, time.perf_counter() - t0)
            print("%-8s %2d threads: %7.2f ms for %d bases = %6.1f GB/s of (seq + corrected) read" %
                  (name, threads, best * 1e3, n, 2 * n / best / 1e9), flush=True)


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""A/B of two builds of libkbbq_b200.so on the same box: build / apply time of one resident batch through the
device-pointer C ABI (signatures common to every revision).  python tools/ab_kernels.py LIB [LIB ...] -- R L N"""
import ctypes as C
import sys

import torch

args = sys.argv[1:]
cut = args.index("--")
libs, (R, L, N) = args[:cut], (int(x) for x in args[cut + 1:cut + 4])
vp, i, i64, sz = C.c_void_p, C.c_int, C.c_int64, C.c_size_t
dev = torch.device("cuda", 0)
P = lambda t: None if t is None else C.c_void_p(t.data_ptr())
NQ = 43
results = {}
for rep in range(2):
    for path in libs:
        lib = C.CDLL(path)
        lib.kbbq_synth_reads.argtypes = [C.c_uint64, i64, i64, i, i] + [vp] * 6
        lib.kbbq_workspace_bytes.argtypes = [i64, i, i, C.POINTER(sz)]
        lib.kbbq_build.argtypes = [vp] * 5 + [i64, i, i, i] + [vp] * 4 + [vp, sz, vp, i, vp]
        lib.kbbq_marginals.argtypes = [vp, vp, i, i] + [vp] * 5 + [vp]
        lib.kbbq_get_delta_qs.argtypes = [vp] * 9 + [i, i, i, i] + [vp] * 4 + [vp]
        lib.kbbq_apply.argtypes = [vp] * 4 + [i64, i, i, i] + [vp] * 5 + [i, i, vp, vp, sz, vp, i, vp]
        seq = torch.empty(N, L, dtype=torch.uint8, device=dev)
        qual, corr, out = torch.empty_like(seq), torch.empty_like(seq), torch.empty_like(seq)
        rg = torch.empty(N, dtype=torch.int16, device=dev)
        sec = torch.empty(N, dtype=torch.uint8, device=dev)
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        assert lib.kbbq_synth_reads(1002, 0, N, L, R, P(seq), P(qual), P(corr), P(rg), P(sec), st) == 0
        nb = sz(0)
        lib.kbbq_workspace_bytes(N, L, R, C.byref(nb))
        ws = torch.empty(nb.value, dtype=torch.uint8, device=dev)
        npos, ndin = R * NQ * 2 * L, R * NQ * 16
        tab = torch.zeros(2 * npos + 2 * ndin, dtype=torch.int64, device=dev)
        pe, pt, de, dt = tab[:npos], tab[npos:2 * npos], tab[2 * npos:2 * npos + ndin], tab[2 * npos + ndin:]
        m = {k: torch.zeros(n, dtype=torch.int64, device=dev) for k, n in
             (("qe", R * NQ), ("qt", R * NQ), ("ge", R), ("gt", R), ("mq", R), ("rgdq", R), ("qdq", R * NQ),
              ("posdq", npos), ("dindq", R * NQ * 17))}
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        rgp = P(rg) if R > 1 else None

        def build():
            assert lib.kbbq_build(P(seq), P(qual), P(corr), rgp, P(sec), N, L, R, 6, P(pe), P(pt), P(de), P(dt), P(ws),
                                  nb.value, P(status), 0, st) == 0

        def apply():
            assert lib.kbbq_apply(P(seq), P(qual), rgp, P(sec), N, L, R, 6, P(m["mq"]), P(m["rgdq"]), P(m["qdq"]),
                                  P(m["posdq"]), P(m["dindq"]), NQ, 17, P(out), P(ws), nb.value, P(status), 0, st) == 0

        def timed(fn, reps=20):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps

        tb = timed(build)
        tab.zero_()
        build()
        lib.kbbq_marginals(P(pe), P(pt), L, R, P(m["qe"]), P(m["qt"]), P(m["ge"]), P(m["gt"]), P(m["mq"]), st)
        lib.kbbq_get_delta_qs(P(m["mq"]), P(m["ge"]), P(m["gt"]), P(m["qe"]), P(m["qt"]), P(pe), P(pt), P(de), P(dt), R, NQ,
                              2 * L, 16, P(m["rgdq"]), P(m["qdq"]), P(m["posdq"]), P(m["dindq"]), st)
        ta = timed(apply)
        key = (int(tab.sum()), int(out.sum(dtype=torch.int64)))
        if "noconf" not in path:
            results.setdefault("check", key)
        if "noconf" not in path:   # the no-conflict timing experiment computes wrong tables on purpose
            assert results["check"] == key, "libraries disagree"
        print("%-40s rep %d  build %.4f ms  apply %.4f ms" % (path, rep, tb, ta))

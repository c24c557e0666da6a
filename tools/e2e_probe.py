"""Where does the end-to-end step go?  PCIe copy rates vs kbbq_recalibrate_host (B200 box), and how the step
depends on the chunk size and on the kind of stores the host-side packer uses.
    python tools/e2e_probe.py [sweep]"""
import ctypes as C, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "kbbq-py_b200"))
from kbbq import _native
from kbbq.device import synth_reads
N, L = 10_000_000, 150
dev = torch.device("cuda", 0)
seq, qual, corr, rg, second = synth_reads(1002, 0, N, L, 1, device=dev)
h = {k: torch.empty(t.shape, dtype=t.dtype, pin_memory=True).copy_(t) for k, t in
     (("seq", seq), ("qual", qual), ("corr", corr), ("second", second))}
h_out = torch.empty(qual.shape, dtype=torch.uint8, pin_memory=True)
torch.cuda.synchronize()
d = torch.empty_like(seq)
for name, fn, nbytes in (("H2D 1.5 GB", lambda: d.copy_(h["seq"], non_blocking=True), seq.numel()),
                         ("D2H 1.5 GB", lambda: h_out.copy_(d, non_blocking=True), seq.numel())):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3): fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    print("%s: %.1f ms  %.1f GB/s" % (name, dt * 1e3, nbytes / dt / 1e9))
lib = _native.lib()
a = {k: v.numpy() for k, v in h.items()}
def step():
    st = C.c_int(0)
    rc = lib.kbbq_recalibrate_host(_native.ptr(a["seq"].reshape(-1)), _native.ptr(a["qual"].reshape(-1)),
                                   _native.ptr(a["corr"].reshape(-1)), None, _native.ptr(a["second"]), N, L, 1, 6,
                                   _native.ptr(h_out.numpy().reshape(-1)), None, None, C.byref(st), 0)
    _native.check(rc, st.value)
del d, seq, qual, corr
torch.cuda.empty_cache()
def timed(label):
    step(); step()
    ts = []
    for i in range(5):
        t0 = time.perf_counter(); step(); ts.append((time.perf_counter() - t0) * 1e3)
    print("%-44s median %.1f ms  (%s)" % (label, sorted(ts)[2], " ".join("%.1f" % t for t in ts)), flush=True)
timed("kbbq_recalibrate_host, defaults")
if len(sys.argv) > 1 and sys.argv[1] == "sweep":
    for chunk in (1_789_569 // 16 * 16, 894_784, 447_392, 223_696, 111_856):
        for nt in ("1", "0"):
            os.environ["KBBQ_HOST_CHUNK_READS"] = str(chunk)
            os.environ["KBBQ_PACK_NT"] = nt
            timed("chunk %d reads, %s stores" % (chunk, "streaming" if nt == "1" else "plain"))

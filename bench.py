#!/usr/bin/env python3
"""bench.py -- throughput of the kbbq recalibration hot path (table build + model + apply).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Metric (BASELINE.json): bases/sec recalibrated (table build + apply).  Workload: BASELINE config 2,
synthetic 10 M x 150 bp interleaved pairs, 1 read group, Q2-Q41, 1 % mismatches, per GPU (weak
scaling: every rank holds its own 10 M-read shard of one global counter-based stream).

One step = one pass of the hot path over the resident batch: zero tables -> build -> (N > 1: one
int64 all-reduce of the tables) -> marginals + delta-Q model -> apply.  `value` is measured with
the packed reads already in HBM; `e2e` is the same pass through the host-buffer C-ABI entry point
(kbbq_recalibrate_host) from pinned host memory, copies inside the timed region.  The batch
(4.5 GB in, 1.5 GB out) is far larger than the 126 MB L2, so no explicit flush is needed.

`--impl reference` times the CPU restatement of the reference's algorithm (oracle/, all host
threads) on a bounded sample of the same workload; the Python reference itself is pure Python,
cannot travel to the GPU box and runs at ~0.23 Mbases/s/core.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "kbbq-py_b200"))

METRIC = "bases/sec recalibrated (table build + apply)"
UNIT = "bases/s"
SEED = 1002
ALGO_BYTES_BUILD = 3  # seq + qual + corrected read, per base
ALGO_BYTES_APPLY = 3  # seq + qual read, new qual written, per base


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML from a background thread (one
    sample per millisecond: the timed region is tens of milliseconds long), nvidia-smi -lms as fallback."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, gpu_index):
        import threading
        self.proc, self.thread, self.stop_flag = None, None, threading.Event()
        self.sm, self.bits, self.smax, self.how = [], 0, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            uuid = None
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(gpu_index).uuid)
            except Exception:
                pass
            handle = None
            if uuid:
                try:
                    handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
                except Exception:
                    handle = None
            if handle is None:
                handle = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM))
            reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                pynvml.nvmlDeviceGetCurrentClocksThrottleReasons

            def loop():
                while not self.stop_flag.is_set():
                    try:
                        self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(handle, pynvml.NVML_CLOCK_SM)))
                        self.bits |= int(reasons(handle))
                    except Exception:
                        pass
                    time.sleep(0.001)
            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
            self.how = "nvml"
            return
        except Exception:
            self.thread = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.how = "nvidia-smi"
        except Exception:
            self.proc = None

    def stop(self):
        if self.thread is not None:
            self.stop_flag.set()
            self.thread.join(timeout=2)
            reasons = sorted(k for k, bit in self.REASONS.items() if self.bits & bit)
            return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.smax,
                    "samples": len(self.sm), "reasons": reasons, "source": "nvml, 1 ms period, timed loops only"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi -lms 100"}


def cpu_port_throughput(n_reads, L, R, seed, target_s, threads=0):
    """Oracle (CPU port of the reference algorithm) on a bounded sample -> (bases/s, sample description)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    from kbbq import synth
    oracle.build()
    # every hardware thread this process may run on; torchrun exports OMP_NUM_THREADS=1 to its workers, which
    # omp_get_max_threads() would obey, so the count is passed explicitly (omp_set_num_threads in the oracle)
    threads = threads or len(os.sched_getaffinity(0)) or oracle.max_threads()
    probe = min(n_reads, 100_000)
    data = synth.synth_reads(seed, 0, probe, L, R)
    t0 = time.perf_counter()
    oracle.recalibrate(*data, L, R, threads=threads)
    rate = probe * L / max(time.perf_counter() - t0, 1e-6)
    sample = int(min(n_reads, max(probe, rate * target_s / L)))
    sample = min(sample, 4_000_000)  # bound host memory of the numpy generator
    if sample > probe:
        data = synth.synth_reads(seed, 0, sample, L, R)
    return oracle, data, sample, threads


def run_reference(args):
    """--impl reference: CPU restatement of the reference algorithm on the host cores (rank 0 only)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    L, R = args.read_len, args.read_groups
    per_step = max(1.0, min(6.0, 150.0 / (args.steps + args.warmup)))
    oracle, data, sample, threads = cpu_port_throughput(args.reads, L, R, SEED, per_step)
    for _ in range(args.warmup):
        oracle.recalibrate(*data, L, R, threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.recalibrate(*data, L, R, threads=threads)
    dt = time.perf_counter() - t0
    value = sample * L * args.steps / dt
    desc = "first %d reads x %d bp of the workload per step (oracle/kbbq_oracle.c, OpenMP)" % (sample, L)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64/u8 (+f64/f80 model)",
        "data": "synthetic",
        "config": {"workload": "BASELINE config 2: synthetic %d x %d bp interleaved pairs, %d read group(s), "
                               "Q2-Q41, 1%% mismatches (CPU arm: bounded sample)" % (args.reads, L, R),
                   "reads_per_step": sample, "read_len": L, "read_groups": R},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "CPU port of the reference algorithm (the reference is pure Python: ~0.23 Mbases/s on one core, "
                "BASELINE.md); host arrays in, host arrays out, no FASTQ parsing on either arm",
    }
    print(json.dumps(line))


def run_streamed(args):
    """BASELINE config 5 (--stream-batch B): `--reads` reads per GPU that do not fit HBM at once, in batches
    of B.  Pass 1 builds the tables over every batch, one all-reduce + model, pass 2 applies batch by
    batch.  The batches are regenerated on the device (counter-based generator), which stands in for the
    host feeding them; only the hot-path calls are timed (CUDA events per batch, summed, max over ranks)."""
    import torch
    import torch.distributed as dist
    from kbbq import _native, parallel
    from kbbq.device import DeviceRecalibrator, synth_reads

    world, rank, local = (int(os.environ.get(k, "0" if k != "WORLD_SIZE" else "1")) for k in ("WORLD_SIZE", "RANK", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    N, L, R, B = args.reads, args.read_len, args.read_groups, args.stream_batch
    rec = DeviceRecalibrator(L, R, max_reads=B, device=dev)
    lib = _native.lib()
    out = torch.empty(B, L, dtype=torch.uint8, device=dev)
    launches0 = lib.kbbq_launch_count()
    ms = {"build": 0.0, "model": 0.0, "apply": 0.0}

    def timed(key, fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms[key] += e0.elapsed_time(e1)

    for phase in ("build", "apply"):
        for lo in range(0, N, B):
            n = min(B, N - lo)
            seq, qual, corr, rg, second = synth_reads(SEED, rank * N + lo, n, L, R, device=dev)
            rg_arg = rg if R > 1 else None
            if phase == "build":
                timed("build", lambda: rec.build(seq, qual, corr, rg_arg, second))
            else:
                timed("apply", lambda: rec.apply(seq, qual, out[:n], rg_arg, second))
            del seq, qual, corr, rg, second
        if phase == "build":
            timed("model", lambda: (rec.allreduce(), rec.model()))
    rec.check_status()
    total_ms = parallel.max_over_ranks(sum(ms.values()), dev)
    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        gbs = 6 * N * L / ((ms["build"] + ms["apply"]) * 1e-3) / 1e9
        print(json.dumps({
            "metric": METRIC, "value": world * N * L / (total_ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": 1, "warmup": 0, "ms_per_step": total_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8 in/out, u32->int64 counts, f64 model", "data": "synthetic",
            "config": {"workload": "BASELINE config 5: synthetic %d x %d bp reads per GPU streamed in batches of %d, "
                                   "%d read group(s); two passes, hot-path calls only" % (N, L, B, R),
                       "reads_per_gpu": N, "read_len": L, "read_groups": R, "batch_reads": B, "seed": SEED},
            "roofline": {"bound": "hbm", "kernel": "build + apply", "achieved": gbs, "peak": peak, "unit": "GB/s",
                         "frac": gbs / peak, "traffic": None, "peak_source": peak_src},
            "phase_ms": ms, "gpu_launches": int(lib.kbbq_launch_count() - launches0)}))
    if world > 1:
        dist.destroy_process_group()


def bind_to_gpu_cpus(gpu_index):
    """Several ranks on one host: run this rank (and first-touch its pinned buffers) on the CPUs NVML names
    as local to its GPU, so that the end-to-end copies do not cross sockets.  Returns the CPU count or None."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(gpu_index).uuid)
        h = pynvml.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus and len(cpus) < len(os.sched_getaffinity(0)):
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from kbbq import _native, parallel
    from kbbq.device import DeviceRecalibrator, synth_reads

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (kbbq_b200 has no CPU path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    total_cpus = len(os.sched_getaffinity(0))
    numa = bind_to_gpu_cpus(local) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _native.lib()
    N, L, R = args.reads, args.read_len, args.read_groups
    K, W = args.steps, max(args.warmup, 3)

    # synthetic shard of this rank, generated on the device (counter-based, see csrc/synth.cuh)
    seq, qual, corr, rg, second = synth_reads(SEED, rank * N, N, L, R, device=dev)
    rg_arg = rg if R > 1 else None
    out = torch.empty_like(qual)
    rec = DeviceRecalibrator(L, R, max_reads=N, device=dev)

    def step(ev=None):
        rec.tables.zero_()
        if ev:
            ev[0].record()
        rec.build(seq, qual, corr, rg_arg, second)
        if ev:
            ev[1].record()
        rec.allreduce()
        rec.model()
        if ev:
            ev[2].record()
        rec.apply(seq, qual, out, rg_arg, second)
        if ev:
            ev[3].record()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(W):
        step()
    rec.check_status()
    barrier()

    sampler = ClockSampler(local) if rank == 0 else None
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(K)]
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = lib.kbbq_launch_count()
    barrier()
    t_start.record()
    for k in range(K):
        step(evs[k])
    t_end.record()
    barrier()
    launches = lib.kbbq_launch_count() - launches0
    total_ms = parallel.max_over_ranks(t_start.elapsed_time(t_end), dev)
    # The step is a fixed chain of a dozen launches (memsets, the pre-pass, the all-reduce, the model kernels
    # around the two hot ones), so it is captured once into a CUDA graph and replayed; the eager loop above
    # keeps the per-phase events.  The figure reported is the replayed one (both are in the line).
    graph_ms = None
    if not args.no_graph:
        # the collective stays outside the captures (NCCL inside a capture hung on this pool): one graph up to
        # the build, the eager all-reduce, one graph from the model on
        g1 = g2 = None
        try:
            g1, g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(g1):
                rec.tables.zero_()
                rec.build(seq, qual, corr, rg_arg, second)
            with torch.cuda.graph(g2):
                rec.model()
                rec.apply(seq, qual, out, rg_arg, second)
            torch.cuda.synchronize()
        except Exception as exc:  # capture not possible: the eager figure stands
            sys.stderr.write("bench.py: CUDA graph capture failed (%r); reporting the eager loop\n" % (exc,))
            g1 = g2 = None

        def replay():
            g1.replay()
            rec.allreduce()
            g2.replay()

        # every rank replays or none does
        if parallel.max_over_ranks(0.0 if g1 is not None else 1.0, dev) == 0.0:
            for _ in range(W):
                replay()
            barrier()
            t_start.record()
            for _ in range(K):
                replay()
            t_end.record()
            barrier()
            graph_ms = parallel.max_over_ranks(t_start.elapsed_time(t_end), dev)
            rec.check_status()
    clocks = sampler.stop() if sampler else None   # sampled over both timed loops
    eager_ms = total_ms
    if graph_ms is not None:
        total_ms = graph_ms
    build_ms = statistics.mean(e[0].elapsed_time(e[1]) for e in evs)
    model_ms = statistics.mean(e[1].elapsed_time(e[2]) for e in evs)
    apply_ms = statistics.mean(e[2].elapsed_time(e[3]) for e in evs)
    rec.check_status()

    # ---- end to end through the host-buffer C-ABI entry point, pinned host memory ----
    e2e = None
    if not args.no_e2e:
        h = {}
        for name, t in (("seq", seq), ("qual", qual), ("corr", corr), ("second", second), ("rg", rg)):
            h[name] = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            h[name].copy_(t)
        h_out = torch.empty(qual.shape, dtype=torch.uint8, pin_memory=True)
        torch.cuda.synchronize()
        a = {k: v.numpy() for k, v in h.items()}
        rg_host = a["rg"].view(np.uint16) if R > 1 else None

        if world == 1:
            def e2e_step():
                st = C.c_int(0)
                rc = lib.kbbq_recalibrate_host(_native.ptr(a["seq"].reshape(-1)), _native.ptr(a["qual"].reshape(-1)),
                                               _native.ptr(a["corr"].reshape(-1)), _native.ptr(rg_host),
                                               _native.ptr(a["second"]), N, L, R, 6,
                                               _native.ptr(h_out.numpy().reshape(-1)), None, None, C.byref(st), local)
                _native.check(rc, st.value)
            api = "kbbq_recalibrate_host (pinned host buffers in, pinned host buffer out)"
        else:
            # several ranks: the tables have to be summed between build and apply, so the step is the
            # device API fed from the pinned host buffers (H2D, build, all-reduce, model, apply, D2H)
            d_in = {k: torch.empty_like(t) for k, t in (("seq", seq), ("qual", qual), ("corr", corr),
                                                         ("second", second), ("rg", rg))}
            d_out = torch.empty_like(qual)
            rec2 = DeviceRecalibrator(L, R, max_reads=N, device=dev)
            # as kbbq_recalibrate_host does, the corrected reads cross PCIe as a mismatch bit map made by this
            # rank's share of the host threads (KBBQ_HOST_NO_BITMAP=1: as they are)
            host_threads = max(1, total_cpus // world)
            # ... when this rank's share is at least 8 threads: with fewer the comparison (27 ms for 1.5 Gbases on
            # 16 threads) takes longer than the 27 ms the corrected reads need on the wire
            use_bits = os.environ.get("KBBQ_HOST_NO_BITMAP", "0") in ("", "0") and host_threads >= 8
            nwords = (N * L + 31) // 32
            h_bits = torch.empty(nwords, dtype=torch.int32, pin_memory=True)
            d_bits = torch.empty(nwords, dtype=torch.int32, device=dev)

            def e2e_step():
                for k in ("seq", "qual", "second") + (("rg",) if R > 1 else ()):
                    d_in[k].copy_(h[k], non_blocking=True)
                rec2.tables.zero_()
                rg_d = d_in["rg"] if R > 1 else None
                if use_bits:
                    _native.check(lib.kbbq_host_mismatch_bits(_native.ptr(a["seq"].reshape(-1)), _native.ptr(a["corr"].reshape(-1)),
                                                              N * L, C.c_void_p(h_bits.data_ptr()), host_threads))
                    d_bits.copy_(h_bits, non_blocking=True)
                    rec2.build_from_bits(d_in["seq"], d_in["qual"], d_bits, d_in["corr"], rg_d, d_in["second"])
                else:
                    d_in["corr"].copy_(h["corr"], non_blocking=True)
                    rec2.build(d_in["seq"], d_in["qual"], d_in["corr"], rg_d, d_in["second"])
                rec2.allreduce()
                rec2.model()
                rec2.apply(d_in["seq"], d_in["qual"], d_out, d_in["rg"] if R > 1 else None, d_in["second"])
                h_out.copy_(d_out, non_blocking=True)
                torch.cuda.synchronize()
                rec2.check_status()
            api = "kbbq.device.DeviceRecalibrator fed from pinned host buffers (H2D, build, all-reduce, model, apply, D2H)"

        e2e_step()  # warm-up (allocations, first-touch)
        ke = max(1, min(K, args.e2e_steps))
        barrier()
        t0 = time.perf_counter()
        for _ in range(ke):
            e2e_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        dt = parallel.max_over_ranks(dt, dev)
        if not torch.equal(h_out.to(dev), out):
            raise SystemExit("bench.py: host-buffer path and device path disagree")
        # bytes that cross PCIe per step: kbbq_recalibrate_host sends the corrected reads as a 1-bit-per-base
        # mismatch map made by the host cores inside the call (csrc/host_pack.cpp) unless KBBQ_HOST_NO_BITMAP=1;
        # the device-API path of the multi-rank step copies all three arrays
        bitmap = os.environ.get("KBBQ_HOST_NO_BITMAP", "0") in ("", "0") and \
            (world == 1 or total_cpus // world >= 8)
        corr_bytes = (N * L + 31) // 32 * 4 if bitmap else N * L
        e2e = {"value": world * N * L * ke / dt, "unit": UNIT,
               "h2d_bytes_per_step": 2 * N * L + corr_bytes + N + (2 * N if R > 1 else 0), "d2h_bytes_per_step": N * L,
               "host_input_bytes_per_step": 3 * N * L + N + (2 * N if R > 1 else 0),
               "ms_per_step": 1e3 * dt / ke, "steps": ke,
               "api": api + ("; corrected reads cross PCIe as a mismatch bit map" if bitmap else "")}

    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak_gbs()
    kernels = {
        "build_smem_kernel": {"ms": build_ms, "algorithmic_bytes": ALGO_BYTES_BUILD * N * L},
        "apply_smem_kernel": {"ms": apply_ms, "algorithmic_bytes": ALGO_BYTES_APPLY * N * L},
    }
    for k in kernels.values():
        k["gbs"] = k["algorithmic_bytes"] / (k["ms"] * 1e-3) / 1e9
        k["frac_of_peak"] = k["gbs"] / peak
    dom = max(kernels, key=lambda n: kernels[n]["ms"])
    combined = (ALGO_BYTES_BUILD + ALGO_BYTES_APPLY) * N * L / ((build_ms + apply_ms) * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            traffic = json.load(fh).get(dom)
    except Exception:
        pass
    line = {
        "metric": METRIC, "value": world * N * L * K / (total_ms * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8 in/out, u32->int64 counts, f64 model", "data": "synthetic",
        "config": {"workload": "BASELINE config 2: synthetic %d x %d bp interleaved pairs per GPU, %d read "
                               "group(s), Q2-Q41, 1%% mismatches" % (N, L, R),
                   "reads_per_gpu": N, "read_len": L, "read_groups": R, "seed": SEED,
                   "l2": "inputs (%.1f GB per step) exceed the 126 MB L2; no explicit flush" % (4 * N * L / 1e9),
                   "parallelism": "reads sharded by rank; one int64 all-reduce of the tables" if world > 1 else "single GPU"},
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["gbs"], "peak": peak, "unit": "GB/s",
                     "frac": kernels[dom]["frac_of_peak"], "traffic": traffic, "peak_source": peak_src,
                     "timing": "CUDA events on the launching stream around the kbbq_build / kbbq_apply call "
                               "(includes the < 1 % work-list pre-pass kernels)",
                     "build_plus_apply_gbs": combined, "build_plus_apply_frac": combined / peak},
        "kernels": kernels,
        "phase_ms": {"build": build_ms, "allreduce+model": model_ms, "apply": apply_ms},
        "launch": {"mode": "CUDA graph replay" if graph_ms is not None else "eager", "eager_ms_per_step": eager_ms / K},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    if e2e:
        line["e2e"] = e2e
    if numa:
        line["config"]["host"] = "each rank bound to the %d CPUs local to its GPU (NVML affinity)" % numa
    if world == 1 and not args.no_cpu:
        oracle, data, sample, threads = cpu_port_throughput(N, L, R, SEED, args.cpu_seconds)
        passes, t0 = 0, time.perf_counter()
        while passes == 0 or time.perf_counter() - t0 < args.cpu_seconds:
            oracle.recalibrate(*data, L, R, threads=threads)
            passes += 1
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": passes * sample * L / dt, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": "first %d reads x %d bp of the workload, %d pass(es) in %.1f s "
                                          "(oracle/kbbq_oracle.c, OpenMP)" % (sample, L, passes, dt)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--reads", type=int, default=10_000_000, help="reads per GPU (config 2: 10 M)")
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("--read-groups", type=int, default=1)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--stream-batch", type=int, default=0,
                    help="config 5: stream --reads reads per GPU through the device in batches of this many")
    ap.add_argument("--no-graph", action="store_true", help="time the eager launch loop instead of the CUDA-graph replay")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.stream_batch > 0:
        run_streamed(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()

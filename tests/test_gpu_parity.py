"""Parity of the CUDA path against the oracle and the reference's golden vectors (B200 only).

Bar: bit-exact for the int64 count tables, the integer delta tables and every output quality byte.
All calls go through the C ABI (device-pointer entry points via kbbq.device, host-buffer entry
points via kbbq._native / the Python API mirror).
"""
import io
import contextlib
import os

import numpy as np
import pytest

from conftest import DELTA_KEYS, TABLE_KEYS, load_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


def _dev(torch, *arrays):
    out = []
    for a in arrays:
        if a is None:
            out.append(None)
        elif a.dtype == np.uint16:
            out.append(torch.from_numpy(a.astype(np.int16)).cuda())
        else:
            out.append(torch.from_numpy(np.ascontiguousarray(a)).cuda())
    return out


def _run_device(torch, seq, qual, corr, rg, second, L, R, path, splits=1):
    from kbbq.device import DeviceRecalibrator
    N = seq.shape[0]
    rec = DeviceRecalibrator(L, R, max_reads=N)
    bounds = [N * i // splits // 16 * 16 for i in range(splits)] + [N]
    parts = []
    for lo, hi in zip(bounds, bounds[1:]):
        parts.append(_dev(torch, seq[lo:hi], qual[lo:hi], corr[lo:hi], rg[lo:hi] if rg is not None else None,
                          second[lo:hi] if second is not None else None))
        s, q, c, g, sec = parts[-1]
        if hi > lo:
            rec.build(s, q, c, g, sec, path=path)
    tables = rec.covariate_arrays()
    deltas = rec.delta_qs()
    outs = []
    for (s, q, c, g, sec) in parts:
        o = torch.full_like(q, 255)
        if q.numel():
            rec.apply(s, q, o, g, sec, path=path)
        outs.append(o.cpu().numpy())
    rec.check_status()
    return tables, deltas, np.concatenate(outs).reshape(N, L)


SEGMENTED = 3  # "path" of the helpers below: the segmented HBM layout (kbbq_segment_* + kbbq_*_segmented)


def _run_segmented(torch, seq, qual, corr, rg, second, L, R, splits=1):
    """Same contract as _run_device, through the segmented batch layout: every split is segmented on the device,
    built and applied with the *_segmented entry points, and its qualities brought back into read order."""
    from kbbq.device import DeviceRecalibrator
    N = seq.shape[0]
    rec = DeviceRecalibrator(L, R, max_reads=N)
    bounds = [N * i // splits // 16 * 16 for i in range(splits)] + [N]
    parts = []
    for lo, hi in zip(bounds, bounds[1:]):
        s, q, c, g, sec = _dev(torch, seq[lo:hi], qual[lo:hi], corr[lo:hi], rg[lo:hi] if rg is not None else None,
                               second[lo:hi] if second is not None else None)
        sb = rec.segment(s, q, c, g, sec)
        parts.append((sb, q))
        rec.build_segmented(sb)
    tables = rec.covariate_arrays()
    deltas = rec.delta_qs()
    outs = []
    for sb, q in parts:
        o_seg = torch.full((sb.rows_bound * L + 16,), 255, dtype=torch.uint8, device=q.device)
        o = torch.full_like(q, 255)
        rec.apply_segmented(sb, o_seg)
        rec.unsegment(sb, o_seg, o)
        outs.append(o.cpu().numpy())
    rec.check_status()
    return tables, deltas, np.concatenate(outs).reshape(N, L)


def _segmented_ok(L, R):
    from kbbq import _native
    return bool(_native.lib().kbbq_segmented_supported(L, R, 6))


@pytest.mark.parametrize("path", [1, 2, SEGMENTED])
def test_golden_cases_device_api(torch_cuda, golden_case, path):
    g = golden_case
    L, R = int(g["L"]), int(g["R"])
    if path == SEGMENTED:
        tables, deltas, out = _run_segmented(torch_cuda, g["seq"], g["qual"], g["corr"], g["rg"], g["second"], L, R)
    else:
        tables, deltas, out = _run_device(torch_cuda, g["seq"], g["qual"], g["corr"], g["rg"], g["second"], L, R, path)
    for got, key in zip(tables, TABLE_KEYS):
        assert np.array_equal(got, g[key]), key
    for got, key in zip(deltas, DELTA_KEYS):
        assert np.array_equal(got, g[key]), key
    assert np.array_equal(out.astype(np.int16), g["outq"])


@pytest.mark.parametrize("path", [1, 2, SEGMENTED])
def test_baseline_shaped_goldens_device_api(torch_cuda, synth_golden_case, path):
    """The reference's own outputs at the BASELINE shapes (L = 150 / 250, R = 1 / 8 / 32)."""
    g = synth_golden_case
    L, R = int(g["L"]), int(g["R"])
    run = _run_segmented if path == SEGMENTED else (lambda *a: _run_device(*a, path))
    tables, deltas, out = run(torch_cuda, g["seq"], g["qual"], g["corr"], g["rg"], g["second"], L, R)
    for got, key in zip(tables, TABLE_KEYS):
        assert np.array_equal(got, g[key]), key
    for got, key in zip(deltas, DELTA_KEYS):
        assert np.array_equal(got, g[key]), key
    assert np.array_equal(out.astype(np.int16), g["outq"])


def test_golden_cases_python_api(golden_case, tmp_path, capfd):
    """The reference's own entry points, same signatures, fed the same FASTQ files."""
    from kbbq import recalibrate
    from kbbq.gatk import applybqsr
    g = golden_case
    names = str(g["names"]).split("\n")
    fu, fc = tmp_path / "u.fq", tmp_path / "c.fq"
    with open(fu, "w") as hu, open(fc, "w") as hc:
        for i, name in enumerate(names):
            q = (g["qual"][i] + 33).astype(np.uint8).tobytes().decode()
            hu.write("@%s\n%s\n+\n%s\n" % (name, g["seq"][i].tobytes().decode(), q))
            hc.write("@%s corrected\n%s\n+\n%s\n" % (name, g["corr"][i].tobytes().decode(), q))
    infer = bool(g["infer_rg"])
    tables = recalibrate.fastq_to_covariate_arrays((str(fu), str(fc)), infer_rg=infer)
    for got, key in zip(tables, TABLE_KEYS):
        assert np.array_equal(got, g[key]), key
        assert got.dtype == np.int64
    dqs = applybqsr.get_delta_qs(*tables)
    for got, key in zip(dqs, DELTA_KEYS):
        assert np.array_equal(got, g[key]), key
    capfd.readouterr()
    recalibrate.recalibrate_fastq((str(fu), str(fc)), infer_rg=infer)
    text = capfd.readouterr().out
    lines = text.split("\n")
    outq = np.array([[ord(ch) - 33 for ch in lines[4 * i + 3]] for i in range(len(names))], dtype=np.int16)
    assert np.array_equal(outq, g["outq"])
    assert all(lines[4 * i] == "@" + names[i] and lines[4 * i + 2] == "+" for i in range(len(names)))
    if "fastq_out" in g.files:
        assert text == str(g["fastq_out"])
    # the same through the batch-streaming driver (what files above 4 GiB take): identical text
    import os
    os.environ["KBBQ_BATCH_READS"] = "48"
    try:
        recalibrate.recalibrate_fastq((str(fu), str(fc)), infer_rg=infer)
    finally:
        del os.environ["KBBQ_BATCH_READS"]
    assert capfd.readouterr().out == text


def test_delta_grid_matches_reference():
    from kbbq import compare_reads
    g = load_case("delta_grid")
    dq = compare_reads.gatk_delta_q(g["prior"], g["errs"], g["total"])
    assert np.array_equal(dq, g["dq"])


def test_near_ties_kernel_equals_long_double(torch_cuda, oracle_mod):
    """Tie-breaking of gatk_delta_q (kbbq/compare_reads.py:257-258: fp64 log-likelihood + long-double prior, first
    maximum wins).  (1) the reference's own answers on constructed near-tie cells; (2) a search of ~1e7 constructed
    cells with 1e10 .. 1e11 observations, where the 64-bit-significand sums of the two best candidates are EQUAL a few
    times per million: the kernel (exact sum rounded to a 64-bit significand) must agree with C long double on all."""
    from kbbq import _native
    g = load_case("delta_near_ties")
    got = _native.delta_q_host(g["prior"].astype(np.int64), g["errs"], g["total"])
    assert np.array_equal(got, g["dq"].astype(np.int64))
    ties = close = 0
    for seed in range(5):
        pq, errs, tot = oracle_mod.near_tie_cells(3_000_000, 100 + seed, 10.0, 11.0)
        want = oracle_mod.gatk_delta_q(pq, errs, tot)
        got = _native.delta_q_host(pq, errs, tot)
        assert np.array_equal(got, want), np.nonzero(got != want)[0][:5]
        gap = oracle_mod.delta_q_top2_gap(pq, errs, tot)
        ties += int((gap == 0).sum())
        close += int((gap < 2.0 ** -60).sum())
    assert close >= 20 and ties >= 10, (close, ties)   # the sample does exercise ties


def test_device_synth_equals_numpy_twin(torch_cuda):
    from kbbq import synth
    from kbbq.device import synth_reads
    for (seed, first, n, L, R) in ((1002, 0, 3000, 150, 1), (1004, 12345, 2001, 250, 32), (7, 10 ** 9, 515, 151, 8)):
        dev = [t.cpu().numpy() for t in synth_reads(seed, first, n, L, R)]
        host = synth.synth_reads(seed, first, n, L, R)
        for name, a, b in zip(("seq", "qual", "corr", "rg", "second"), dev, host):
            assert np.array_equal(a.view(b.dtype) if a.dtype != b.dtype else a, b), name


SYNTH_CASES = [
    # (name, seed, N, L, R)  -- the BASELINE.json configs at oracle-sized N
    ("c2_l150_r1", 1002, 200_000, 150, 1),
    ("c2_l150_r1_partial_last_group", 1002, 100_003, 150, 1),
    ("c3_l150_r8", 1003, 120_000, 150, 8),
    ("c4_l250_r32", 1004, 60_000, 250, 32),
    ("c1_l151_r1", 1001, 24_000, 151, 1),
    ("l100_r3_ragged_n", 5, 33_333, 100, 3),
    ("l150_r1_ragged_second", 9, 50_001, 150, 1),
    ("l37_r5", 6, 10_007, 37, 5),
    ("l8_r2", 8, 4_099, 8, 2),
]


@pytest.mark.parametrize("case", SYNTH_CASES, ids=[c[0] for c in SYNTH_CASES])
def test_synthetic_configs_vs_oracle(torch_cuda, oracle_mod, case):
    from kbbq import synth
    _, seed, N, L, R = case
    seq, qual, corr, rg, second = synth.synth_reads(seed, 0, N, L, R)
    if "ragged" in case[0]:  # unpaired: every read its own group / orientation
        rng = np.random.default_rng(seed)
        rg = rng.integers(0, R, N).astype(np.uint16)
        second = rng.integers(0, 2, N).astype(np.uint8)
    want_t = oracle_mod.covariate_arrays(seq, qual, corr, rg, second, L, R)
    want_d = oracle_mod.get_delta_qs(*want_t)
    want_o = oracle_mod.apply(seq, qual, rg, second, L, R, want_t[0], *want_d)
    for path, splits in ((1, 1), (1, 3), (2, 1), (SEGMENTED, 1), (SEGMENTED, 3)):
        if path == SEGMENTED:
            tables, deltas, out = _run_segmented(torch_cuda, seq, qual, corr, rg, second, L, R, splits)
        else:
            tables, deltas, out = _run_device(torch_cuda, seq, qual, corr, rg, second, L, R, path, splits)
        for got, want, key in zip(tables, want_t, TABLE_KEYS):
            assert np.array_equal(got, want), (key, path, splits)
        for got, want, key in zip(deltas, want_d, DELTA_KEYS):
            assert np.array_equal(got, want), (key, path, splits)
        assert np.array_equal(out.astype(np.int16), want_o), (path, splits)


def _set_transport(monkeypatch, transport):
    monkeypatch.setenv("KBBQ_HOST_BITMAP", "1")   # pack whatever the core count of the box
    monkeypatch.setenv("KBBQ_HOST_NO_BITMAP", "1" if transport == "plain" else "0")
    monkeypatch.setenv("KBBQ_HOST_NO_NIBBLES", "1" if transport == "bits" else "0")


@pytest.mark.parametrize("transport", ["nibbles", "bits", "plain"])
@pytest.mark.parametrize("streaming", ["0", "1"])
@pytest.mark.parametrize("R", [4, 1])
def test_host_buffer_entry_point_chunked(oracle_mod, monkeypatch, streaming, transport, R):
    """kbbq_recalibrate_host with several chunks, resident and two-pass streaming modes, one and several read
    groups; the reads and the corrected reads crossing PCIe as 4 bits per base (default), the corrected reads as
    a mismatch bit map, or everything as it is."""
    from kbbq import _native, synth
    N, L = 50_003, 151   # odd sizes: the last chunk's packed form ends inside a byte / word
    seq, qual, corr, rg, second = synth.synth_reads(11, 0, N, L, R)
    corr[::7, 3] = ord("N")    # corrected bases outside ACGT only have to differ
    seq[::11, 5] = ord("N")
    monkeypatch.setenv("KBBQ_HOST_CHUNK_READS", "7000")
    monkeypatch.setenv("KBBQ_HOST_FORCE_STREAMING", streaming)
    _set_transport(monkeypatch, transport)
    out, tabs, dqs = _native.recalibrate_host(seq, qual, corr, rg, second, L, R, want_tables=True)
    want_t = oracle_mod.covariate_arrays(seq, qual, corr, rg, second, L, R)
    want_d = oracle_mod.get_delta_qs(*want_t)
    want_o = oracle_mod.apply(seq, qual, rg, second, L, R, want_t[0], *want_d)
    for got, want in zip(tabs, want_t[5:]):
        assert np.array_equal(got, want)
    assert np.array_equal(dqs[0], want_t[0])
    for got, want in zip(dqs[1:], want_d):
        assert np.array_equal(got, want)
    assert np.array_equal(out.astype(np.int16), want_o)


def test_expand_nibbles_kernel(torch_cuda):
    """kbbq_expand_nibbles: the device side of kbbq_host_pack_nibbles -- seq comes back byte for byte, corr differs
    from it exactly where the corrected read did."""
    torch = torch_cuda
    import ctypes as C
    from kbbq import _native
    lib = _native.lib()
    rng = np.random.default_rng(12)
    alphabet = np.frombuffer(b"ACGTN", np.uint8)
    for n in (1, 31, 32, 33, 8192 * 3 + 7, 1_000_001):
        seq = alphabet[rng.integers(0, 5, size=n)]
        corr = seq.copy()
        flip = rng.random(n) < 0.1
        corr[flip] = alphabet[rng.integers(0, 5, size=int(flip.sum()))]
        packed = np.zeros((n + 1) // 2, np.uint8)
        bad = C.c_int(0)
        assert lib.kbbq_host_pack_nibbles(_native.ptr(seq), _native.ptr(corr), n, _native.ptr(packed), 0, C.byref(bad)) == 0
        d_packed = torch.from_numpy(packed).cuda()
        d_seq = torch.full((n + 64,), 255, dtype=torch.uint8, device="cuda")
        d_corr = torch.full((n + 64,), 255, dtype=torch.uint8, device="cuda")
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _native.check(lib.kbbq_expand_nibbles(C.c_void_p(d_packed.data_ptr()), n, C.c_void_p(d_seq.data_ptr()),
                                              C.c_void_p(d_corr.data_ptr()), st))
        got_s, got_c = d_seq.cpu().numpy(), d_corr.cpu().numpy()
        assert np.array_equal(got_s[:n], seq)
        assert np.array_equal(got_c[:n] != got_s[:n], corr != seq)
        assert np.all(got_s[n:] == 255) and np.all(got_c[n:] == 255)   # nothing written past the end


@pytest.mark.parametrize("transport", ["nibbles", "bits", "plain"])
def test_host_entry_point_input_errors(monkeypatch, transport):
    """A base outside ACGTN -> TypeError, a quality above 42 -> IndexError through kbbq_recalibrate_host, whichever
    form the reads cross PCIe in (with nibbles the base check happens on the host: the device never sees the byte)."""
    from kbbq import _native, synth
    N, L, R = 20_000, 100, 2
    seq, qual, corr, rg, second = synth.synth_reads(21, 0, N, L, R)
    monkeypatch.setenv("KBBQ_HOST_CHUNK_READS", "6000")
    _set_transport(monkeypatch, transport)
    for bad_byte in (ord("X"), ord("a"), ord("O"), 0):
        s2 = seq.copy()
        s2[13_333, 77] = bad_byte
        with pytest.raises(TypeError):
            _native.recalibrate_host(s2, qual, corr, rg, second, L, R)
    q2 = qual.copy()
    q2[19_999, 99] = 43
    with pytest.raises(IndexError):
        _native.recalibrate_host(seq, q2, corr, rg, second, L, R)
    out = _native.recalibrate_host(seq, qual, corr, rg, second, L, R)   # the session recovers
    assert out.shape == (N, L)


def test_input_errors_raise_like_the_reference(torch_cuda):
    from kbbq import _native
    L = 40
    seq = np.frombuffer(b"ACGT" * 10 * 8, np.uint8).reshape(8, L).copy()
    qual = np.full((8, L), 30, np.uint8)
    bad_q = qual.copy()
    bad_q[3, 17] = 43
    with pytest.raises(IndexError):
        _native.build_host(seq, bad_q, seq, None, None, L, 1)
    bad_s = seq.copy()
    bad_s[5, 20] = ord("X")
    with pytest.raises(TypeError):
        _native.build_host(bad_s, qual, seq, None, None, L, 1)
    lower = seq.copy()
    lower[0, 0] = ord("a")
    with pytest.raises(TypeError):
        _native.build_host(lower, qual, seq, None, None, L, 1)
    with pytest.raises(IndexError):
        _native.build_host(seq, qual, seq, np.full(8, 2, np.uint16), None, L, 2)
    # empty batch is fine
    pe, pt, de, dt = _native.build_host(seq[:0], qual[:0], seq[:0], None, None, L, 1)
    assert pt.sum() == 0


FULL_SIZE = [
    # BASELINE.json configs at full per-GPU size: (id, seed, reads, read length, read groups)
    ("config2_10M_x150_r1", 1002, 10_000_000, 150, 1),
    ("config3_shard_25M_x150_r8", 1003, 25_000_000, 150, 8),   # 200 M reads over 8 GPUs
    ("config4_20M_x250_r32", 1004, 20_000_000, 250, 32),
]


@pytest.mark.parametrize("case", FULL_SIZE, ids=[c[0] for c in FULL_SIZE])
def test_full_size_properties(torch_cuda, case):
    """BASELINE configs at full size: size-independent properties (checksums of checksums, marginals,
    linearity, the generic kernels as an independent second implementation)."""
    torch = torch_cuda
    from kbbq.device import DeviceRecalibrator, synth_reads
    _, seed, N, L, R = case
    seq, qual, corr, rg, second = synth_reads(seed, 0, N, L, R)
    if R == 1:
        rg = None
    rec = DeviceRecalibrator(L, R, max_reads=N)
    rec.build(seq, qual, corr, rg, second)
    whole = rec.tables.clone()
    valid = qual >= 6
    err = (seq != corr) & valid
    # checksum of checksums: every tallied base lands in exactly one cell
    assert int(rec.pos_total.sum()) == int(valid.sum())
    assert int(rec.pos_errs.sum()) == int(err.sum())
    # per-cycle marginal == per-column count, forward half for read 1 and reversed half for read 2
    col = valid.view(N // 2, 2, L).sum(0)
    per_cycle = rec.pos_total.sum((0, 1))
    assert torch.equal(per_cycle[:L], col[0]) and torch.equal(per_cycle[L:].flip(0), col[1])
    # per-quality marginal == histogram of the quality bytes
    hist = sum(torch.bincount(qual[lo:lo + 2_000_000].reshape(-1).to(torch.int64), minlength=43)
               for lo in range(0, N, 2_000_000))
    hist[:6] = 0
    assert torch.equal(rec.pos_total.sum((0, 2)), hist)
    if rg is not None:  # per-read-group marginal == valid bases of the reads of that group
        per_rg = torch.zeros(R, dtype=torch.int64, device=qual.device)
        per_rg.index_add_(0, rg.to(torch.int64), valid.sum(1))
        assert torch.equal(rec.pos_total.sum((1, 2)), per_rg)
    assert int(rec.din_total.sum()) <= int(valid.sum()) - int(valid[:, 0].sum())
    assert bool((rec.din_errs <= rec.din_total).all()) and bool((rec.pos_errs <= rec.pos_total).all())
    # linearity: two half batches add up to the whole; the generic kernel agrees with the smem kernel
    rec.reset()
    h = N // 2
    rec.build(seq[:h], qual[:h], corr[:h], None if rg is None else rg[:h], second[:h])
    rec.build(seq[h:], qual[h:], corr[h:], None if rg is None else rg[h:], second[h:])
    assert torch.equal(rec.tables, whole)
    rec.reset()
    rec.build(seq, qual, corr, rg, second, path=2)
    assert torch.equal(rec.tables, whole)
    # the segmented layout (rows sorted by read group and mate) gives the same tables ...
    rec.reset()
    sb = rec.segment(seq, qual, corr, rg, second)
    rec.build_segmented(sb)
    assert torch.equal(rec.tables, whole)
    # apply: untouched below minscore, bounded otherwise, smem == generic
    rec.model()
    out = torch.empty_like(qual)
    rec.apply(seq, qual, out, rg, second)
    assert torch.equal(out[~valid], qual[~valid])
    assert int(out[valid].max()) <= 60
    out2 = torch.empty_like(qual)
    rec.apply(seq, qual, out2, rg, second, path=2)
    assert torch.equal(out, out2)
    # ... and the same output bytes once they are back in read order
    out_seg = torch.empty(sb.rows_bound * L + 16, dtype=torch.uint8, device=qual.device)
    rec.apply_segmented(sb, out_seg)
    out2.fill_(255)
    rec.unsegment(sb, out_seg, out2)
    assert torch.equal(out, out2)
    rec.check_status()


PACKED_COUNTER_CASES = [
    # (L, N, R): short reads of ONE quality -- every read hits the same row of the cycle table at every cycle, so a
    # CTA's slice (N / 148 reads) puts far more than 65 024 hits into single shared-memory cells between segments
    (16, 20_000_000, 1),
    (8, 24_000_000, 2),
    (24, 12_000_000, 1),
]


@pytest.mark.parametrize("case", PACKED_COUNTER_CASES, ids=["l%d_n%dm_r%d" % (c[0], c[1] // 1_000_000, c[2]) for c in PACKED_COUNTER_CASES])
def test_packed_counters_do_not_carry(torch_cuda, oracle_mod, case):
    """The shared-memory counters pack total + 65025 * errors in one u32 (build.cuh): the cycle table has to be
    flushed, and the dinucleotide replicas folded, before any total field reaches 65025 -- whatever the read
    length.  Constant qualities on tens of millions of short reads reach that regime (for L < 29 the fold period
    used to exceed the flush period, and a chunk was only bounded by the former)."""
    torch = torch_cuda
    from kbbq.device import DeviceRecalibrator
    L, N, R = case
    rng = np.random.default_rng(L)
    seq = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, (N, L), dtype=np.uint8)]
    qual = np.full((N, L), 30, np.uint8)
    corr = seq.copy()
    corr[rng.random((N, L)) < 0.02] = ord("A")      # ~1.5 % mismatches
    rg = (np.arange(N) // 2 % R).astype(np.uint16) if R > 1 else None
    second = (np.arange(N) & 1).astype(np.uint8)
    want = oracle_mod.covariate_arrays(seq, qual, corr, rg if rg is not None else np.zeros(N, np.uint16), second, L, R)[5:]
    s, q, c, g, sec = _dev(torch, seq, qual, corr, rg, second)
    rec = DeviceRecalibrator(L, R, max_reads=N)
    rec.build(s, q, c, g, sec, path=1)
    got = rec.covariate_arrays()[5:]
    for a, b, key in zip(got, want, TABLE_KEYS[5:]):
        assert np.array_equal(a, b), (key, "read order")
    if _segmented_ok(L, R):
        rec.reset()
        rec.build_segmented(rec.segment(s, q, c, g, sec))
        for a, b, key in zip(rec.covariate_arrays()[5:], want, TABLE_KEYS[5:]):
            assert np.array_equal(a, b), (key, "segmented")
    rec.check_status()


def test_random_shapes_vs_oracle(torch_cuda, oracle_mod):
    """Seeded sweep over read lengths (every alignment class, the generic-kernel range L > 288 included),
    read-group counts, batch sizes that leave partial groups, unpaired reads and N-rich data; path 0 is
    what a caller gets (shared-memory kernels when the plan fits, generic otherwise)."""
    from kbbq import synth
    rng = np.random.default_rng(2026)
    shapes = [(4, 1), (5, 2), (7, 3), (9, 1), (33, 4), (64, 2), (75, 1), (99, 3), (101, 1), (126, 6), (149, 2),
              (152, 1), (200, 9), (251, 2), (287, 1), (288, 2), (289, 1), (300, 3), (400, 1)]
    for L, R in shapes:
        N = int(rng.integers(1, 9000))
        seq, qual, corr, rg, second = synth.synth_reads(int(rng.integers(1, 1 << 30)), 0, N, L, R)
        if rng.random() < 0.5:   # unpaired: read groups and mates in no particular order
            rg = rng.integers(0, R, N).astype(np.uint16)
            second = rng.integers(0, 2, N).astype(np.uint8)
        if rng.random() < 0.5:   # N-rich, with low qualities elsewhere too
            isn = rng.random((N, L)) < 0.1
            seq = np.where(isn, np.uint8(ord("N")), seq)
            qual = np.where(rng.random((N, L)) < 0.1, rng.integers(0, 8, (N, L)), qual).astype(np.uint8)
        want_t = oracle_mod.covariate_arrays(seq, qual, corr, rg, second, L, R)
        want_d = oracle_mod.get_delta_qs(*want_t)
        want_o = oracle_mod.apply(seq, qual, rg, second, L, R, want_t[0], *want_d)
        runs = [_run_device(torch_cuda, seq, qual, corr, rg, second, L, R, 0, int(rng.integers(1, 4)))]
        if _segmented_ok(L, R):
            runs.append(_run_segmented(torch_cuda, seq, qual, corr, rg, second, L, R, int(rng.integers(1, 4))))
        for tables, deltas, out in runs:
            for got, want, key in zip(tables, want_t, TABLE_KEYS):
                assert np.array_equal(got, want), (key, L, R, N)
            for got, want, key in zip(deltas, want_d, DELTA_KEYS):
                assert np.array_equal(got, want), (key, L, R, N)
            assert np.array_equal(out.astype(np.int16), want_o), (L, R, N)


def test_segmented_layout_invariants(torch_cuda):
    """kbbq_segment_plan: spans sorted by 2 * rg + second, 16-row aligned, dest a bijection onto the occupied
    rows; a malformed span table is refused by the kernels (KBBQ_FLAG_SEGMENTS -> ValueError)."""
    torch = torch_cuda
    from kbbq import synth
    from kbbq.device import DeviceRecalibrator
    N, L, R = 20_011, 150, 5
    seq, qual, corr, rg, second = synth.synth_reads(77, 0, N, L, R)
    rg = np.random.default_rng(1).integers(0, R - 1, N).astype(np.uint16)   # the last read group stays empty
    s, q, c, g, sec = _dev(torch, seq, qual, corr, rg, second)
    rec = DeviceRecalibrator(L, R, max_reads=N)
    sb = rec.segment(s, q, c, g, sec)
    seg = sb.seg.cpu().numpy().astype(np.int64)
    off, cnt = seg[:2 * R + 1], seg[2 * R + 1:4 * R + 1]
    key = 2 * rg.astype(np.int64) + second
    assert off[0] == 0 and np.all(off % 16 == 0) and np.all(np.diff(off) >= 0)
    assert np.array_equal(cnt, np.bincount(key, minlength=2 * R))
    assert np.all(off[1:] - off[:-1] >= cnt) and np.all(off[1:] - off[:-1] - cnt < 16)
    dest = sb.dest.cpu().numpy().astype(np.int64)[:N]
    assert np.all((dest >= off[key]) & (dest < off[key] + cnt[key])) and len(np.unique(dest)) == N
    rows = sb.seq[:sb.rows_bound * L].view(-1, L).cpu().numpy()
    assert np.array_equal(rows[dest], seq)
    pad = np.ones(off[-1], bool)
    pad[dest] = False
    assert np.all(sb.qual[:sb.rows_bound * L].view(-1, L).cpu().numpy()[:off[-1]][pad] == 0)
    rec.check_status()
    bad = sb.seg.clone()
    bad[1] = 8   # not a multiple of 16
    sb.seg = bad
    rec.build_segmented(sb)
    with pytest.raises(ValueError):
        rec.check_status()

"""Host-buffer side of the C ABI on the GPU: sessions (kbbq_session_*), the several-GPU entry point
(kbbq_recalibrate_host_multi) and the torch.distributed form, all against the oracle.  Bar: bit-exact."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _want(oracle_mod, seq, qual, corr, rg, second, L, R):
    t = oracle_mod.covariate_arrays(seq, qual, corr, rg, second, L, R)
    d = oracle_mod.get_delta_qs(*t)
    o = oracle_mod.apply(seq, qual, rg, second, L, R, t[0], *d)
    return t, d, o


def _check(got, tabs, dqs, want):
    t, d, o = want
    if tabs is not None:
        for a, b in zip(tabs, t[5:]):
            assert np.array_equal(a, b)
    if dqs is not None:
        assert np.array_equal(dqs[0], t[0])
        for a, b in zip(dqs[1:], d):
            assert np.array_equal(a, b)
    assert np.array_equal(got.astype(np.int16), o)


@pytest.mark.parametrize("R,L", [(1, 150), (4, 151), (8, 100)])
@pytest.mark.parametrize("resident", [True, False])
def test_session_chunks(oracle_mod, R, L, resident):
    from kbbq import _native, synth
    N, C = 30_011, 7008
    seq, qual, corr, rg, second = synth.synth_reads(21, 0, N, L, R)
    want = _want(oracle_mod, seq, qual, corr, rg, second, L, R)
    out = np.full((N, L), 255, np.uint8)
    with _native.Session(L, R, 6, chunk_reads=C, resident_reads=N if resident else 0) as s:
        for rep in range(2):   # a session is reusable after reset()
            if rep:
                s.reset()
                out.fill(255)
            if rep == 0:
                for lo in range(0, N, C):
                    hi = min(N, lo + C)
                    s.build_chunk(seq[lo:hi], qual[lo:hi], corr[lo:hi], rg[lo:hi], second[lo:hi], keep=resident)
            else:   # the same reads as one chunk on its own followed by a pipelined range (kbbq_session_build_range)
                s.build_chunk(seq[:C], qual[:C], corr[:C], rg[:C], second[:C], keep=resident)
                assert s.build_range(seq[C:], qual[C:], corr[C:], rg[C:], second[C:]) == -(-(N - C) // C)
            tabs = s.tables()
            dqs = s.model(want_deltas=True)
            for k, lo in enumerate(range(0, N, C)):
                hi = min(N, lo + C)
                if resident:
                    s.apply_resident(k, out[lo:hi])
                else:
                    s.apply_chunk(seq[lo:hi], qual[lo:hi], rg[lo:hi], second[lo:hi], out[lo:hi])
            s.sync()
            npos, ndin = R * 43 * 2 * L, R * 43 * 16
            tt = (tabs[:npos].reshape(R, 43, 2 * L), tabs[npos:2 * npos].reshape(R, 43, 2 * L),
                  tabs[2 * npos:2 * npos + ndin].reshape(R, 43, 16), tabs[2 * npos + ndin:].reshape(R, 43, 16))
            _check(out, tt, dqs, want)
            h2d, d2h = s.traffic()
            assert d2h == N * L and h2d > N * L   # qualities as they are + at least 4 bits per base for the reads


def test_session_tables_add_up(oracle_mod):
    """Two sessions over two halves, tables exchanged through the host, give the one-session result."""
    from kbbq import _native, synth
    N, L, R = 20_000, 150, 3
    seq, qual, corr, rg, second = synth.synth_reads(22, 0, N, L, R)
    want = _want(oracle_mod, seq, qual, corr, rg, second, L, R)
    h = N // 2 // 16 * 16
    out = np.full((N, L), 255, np.uint8)
    parts = [(0, h), (h, N)]
    sess = [_native.Session(L, R, 6, chunk_reads=hi - lo, resident_reads=hi - lo) for lo, hi in parts]
    for s, (lo, hi) in zip(sess, parts):
        s.build_chunk(seq[lo:hi], qual[lo:hi], corr[lo:hi], rg[lo:hi], second[lo:hi], keep=True)
    total = sum(s.tables() for s in sess)
    for s, (lo, hi) in zip(sess, parts):
        s.set_tables(total)
        s.model()
        s.apply_resident(0, out[lo:hi])
        s.sync()
        s.close()
    _check(out, None, None, want)


@pytest.mark.parametrize("no_p2p", ["0", "1"])
@pytest.mark.parametrize("R,L", [(1, 150), (8, 150)])
def test_multi_device_entry_point_same_result(oracle_mod, monkeypatch, R, L, no_p2p):
    """kbbq_recalibrate_host_multi over [0], [0, 0, 0] and (with 2+ GPUs) [0, 1]: identical tables, deltas and
    output bytes; the table sum over peer memory and the one through the host."""
    import torch
    from kbbq import _native, synth
    N = 40_013
    seq, qual, corr, rg, second = synth.synth_reads(23, 0, N, L, R)
    want = _want(oracle_mod, seq, qual, corr, rg, second, L, R)
    monkeypatch.setenv("KBBQ_MULTI_NO_P2P", no_p2p)
    lists = [[0], [0, 0, 0]]
    if torch.cuda.device_count() >= 2:
        lists += [[0, 1], [1, 0, 1]]
    for devs in lists:
        out, tabs, dqs = _native.recalibrate_host(seq, qual, corr, rg, second, L, R, want_tables=True, devices=devs)
        _check(out, tabs, dqs, want)


def test_two_gpus_byte_for_byte(oracle_mod):
    """SURVEY.md section 8e acceptance: 1-GPU vs 2-GPU tables and output bytes identical (needs two devices)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from kbbq import _native, synth
    N, L, R = 300_000, 150, 8
    seq, qual, corr, rg, second = synth.synth_reads(1003, 0, N, L, R)
    one = _native.recalibrate_host(seq, qual, corr, rg, second, L, R, want_tables=True, devices=[0])
    two = _native.recalibrate_host(seq, qual, corr, rg, second, L, R, want_tables=True, devices=[0, 1])
    assert np.array_equal(one[0], two[0])
    for a, b in zip(one[1] + one[2], two[1] + two[2]):
        assert np.array_equal(a, b)
    _check(two[0], two[1], two[2], _want(oracle_mod, seq, qual, corr, rg, second, L, R))


def test_distributed_form_single_rank(oracle_mod):
    """kbbq.parallel.recalibrate_host_distributed without a process group is the one-GPU path."""
    from kbbq import parallel, synth
    N, L, R = 25_000, 150, 2
    seq, qual, corr, rg, second = synth.synth_reads(24, 0, N, L, R)
    out = np.full((N, L), 255, np.uint8)
    s = parallel.recalibrate_host_distributed(seq, qual, corr, rg, second, L, R, out)
    _check(out, None, None, _want(oracle_mod, seq, qual, corr, rg, second, L, R))
    # the table buffer is visible to torch without a copy
    import torch
    t = parallel.session_tables_tensor(s, torch.device("cuda", 0))
    assert t.dtype == torch.int64 and int(t.sum()) == int(s.tables().sum())
    s.close()


def test_fastq_shorter_corrected_file(tmp_path, capfd, oracle_mod):
    """fastq[0] longer than the corrected file: tables from the pairs zip() yields (kbbq/recalibrate.py:56-57),
    every read of fastq[0] recalibrated and written (kbbq/recalibrate.py:141-156)."""
    from kbbq import recalibrate, synth
    N, n, L = 600, 480, 60
    seq, qual, corr, rg, second = synth.synth_reads(25, 0, N, L, 1)
    names = synth.write_fastq(str(tmp_path / "u.fq"), str(tmp_path / "c_all.fq"), seq, qual, corr, rg, second, infer_rg=False)
    with open(tmp_path / "c_all.fq") as fh:
        lines = fh.readlines()
    with open(tmp_path / "c.fq", "w") as fh:
        fh.writelines(lines[:4 * n])
    capfd.readouterr()
    recalibrate.recalibrate_fastq((str(tmp_path / "u.fq"), str(tmp_path / "c.fq")))
    got = capfd.readouterr().out.split("\n")
    t = oracle_mod.covariate_arrays(seq[:n], qual[:n], corr[:n], rg[:n], second[:n], L, 1)
    d = oracle_mod.get_delta_qs(*t)
    o = oracle_mod.apply(seq, qual, rg, second, L, 1, t[0], *d)
    assert len(got) == 4 * N + 1
    outq = np.array([[ord(ch) - 33 for ch in got[4 * i + 3]] for i in range(N)], dtype=np.int16)
    assert np.array_equal(outq, o)
    assert all(got[4 * i] == "@" + names[i] for i in range(N))

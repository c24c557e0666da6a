"""BAM side of the path.

Golden vectors tests/golden/bam_*.npz: the unmodified reference (bam_to_bqsr_covariates,
find_read_errors, trim_bamread, get_delta_qs, recalibrate_bamread) run on in-memory stand-ins for
pysam reads by tests/golden/make_golden_bam.py.  CPU part: the oracle restatement and the host logic
(CIGAR walk, adaptor trimming, per-read covariates).  GPU part: kbbq_build_bam / kbbq_apply_bam
through the C ABI, bit-exact."""
import os
import sys

import numpy as np
import pytest

from conftest import DELTA_KEYS, ROOT, TABLE_KEYS, load_case

CASES = ["bam_mixed_r2", "bam_long_r3"]


@pytest.fixture(scope="module")
def standin():
    """The in-memory stand-in for pysam (test infrastructure, oracle/ref_shim/stubs)."""
    stubs = os.path.join(ROOT, "oracle", "ref_shim", "stubs")
    had = sys.modules.pop("pysam", None)
    sys.path.insert(0, stubs)
    import pysam
    yield pysam
    sys.path.remove(stubs)
    sys.modules.pop("pysam", None)
    if had is not None:
        sys.modules["pysam"] = had


def rebuild_reads(pysam, d):
    reads = []
    for i in range(d["seq"].shape[0]):
        cig = [(int(op), int(n)) for op, n in d["cigar"][i] if op >= 0]
        read = pysam.AlignedSegment("r%d" % i, d["seq"][i].tobytes().decode(), d["bamq"][i].tolist(), cig, "chr1",
                                    int(d["ref_start"][i]), bool(d["flags"][i] & 2), bool(d["flags"][i] & 1),
                                    {"OQ": (d["qual"][i] + 33).astype(np.uint8).tobytes().decode(),
                                     "RG": "rg%d" % d["rg"][i]})
        read.tlen = read.template_length = int(d["tlen"][i])
        read.next_reference_start = int(d["next_start"][i])
        reads.append(read)
    return reads


@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_the_reference(oracle_mod, case):
    d = load_case(case)
    L, R = int(d["L"]), int(d["R"])
    tabs = oracle_mod.build_tables_bam(d["seq"], d["qual"], d["err"], d["skip"], d["rg"], d["flags"],
                                       d["aln_start"], d["aln_end"], L, R)
    for key, t in zip(TABLE_KEYS[5:], tabs):
        assert np.array_equal(t, d[key]), key
    marg = oracle_mod.marginals(tabs[0], tabs[1])
    for key, t in zip(TABLE_KEYS[:5], marg):
        assert np.array_equal(t, d[key]), key
    out = oracle_mod.apply_bam(d["seq"], d["qual"], d["rg"], d["flags"], L, R, d["meanq"],
                               *[d[k] for k in DELTA_KEYS])
    assert np.array_equal(out, d["outq"])


@pytest.mark.parametrize("case", CASES)
def test_host_cigar_walk_trimming_and_covariates(standin, case):
    from kbbq import compare_reads
    from kbbq.gatk import applybqsr, bqsr
    d = load_case(case)
    L, R = int(d["L"]), int(d["R"])
    ref = {"chr1": np.array(list(d["ref"].tobytes().decode()), dtype=np.str_)}
    variable = {"chr1": np.zeros(d["ref"].size, bool)}
    variable["chr1"][d["variable"]] = True
    reads = rebuild_reads(standin, d)
    pos_total, pos_errs = np.zeros_like(d["pos_total"]), np.zeros_like(d["pos_errs"])
    din_total = np.zeros_like(d["dinuc_total"])
    trimmed = 0
    for i, read in enumerate(reads):
        e, s = compare_reads.find_read_errors(read, ref, variable)
        t = bqsr.trim_bamread(read)
        trimmed += int(t.any())
        assert np.array_equal(e, d["err"][i].astype(bool)), i
        assert np.array_equal(np.logical_or(s, t), d["skip"][i].astype(bool)), i
        assert np.array_equal(compare_reads.bamread_get_oq(read), d["qual"][i])
        assert (read.query_alignment_start, read.query_alignment_end) == (d["aln_start"][i], d["aln_end"][i])
        # the per-read covariates, tallied the way the reference does (kbbq/gatk/bqsr.py:98-117)
        q = d["qual"][i].astype(int)
        valid = ~(np.logical_or(s, t) | (q < 6) | (d["seq"][i] == ord("N")))
        cyc, din = bqsr.bamread_bqsr_cycle(read), bqsr.bamread_bqsr_dinuc(read)
        g = int(d["rg"][i])
        np.add.at(pos_total, (g, q[valid], cyc[valid]), 1)
        np.add.at(pos_errs, (g, q[valid & e], cyc[valid & e]), 1)
        dv = valid & (din != -1)
        np.add.at(din_total, (g, q[dv], din[dv]), 1)
        c2 = applybqsr.bamread_cycle_covariates(read)
        want = np.arange(L) if not read.is_read2 else -(np.arange(L) + 1)
        assert np.array_equal(c2, want[::-1] if read.is_reverse else want)
    assert trimmed > 10  # the adaptor-trimming branch is exercised
    assert np.array_equal(pos_total, d["pos_total"]) and np.array_equal(pos_errs, d["pos_errs"])
    assert np.array_equal(din_total, d["dinuc_total"])
    assert compare_reads.get_rg_to_pu(standin.AlignmentFile(header={"RG": [{"ID": "a", "PU": "u"}]})) == {"a": "u"}


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_bam_arrays_match_the_reference(case):
    from kbbq.gatk import applybqsr, bqsr
    d = load_case(case)
    R = int(d["R"])
    got = bqsr.bam_arrays_to_bqsr_covariates(d["seq"], d["qual"], d["err"], d["skip"], d["rg"], d["flags"] & 1,
                                             (d["flags"] >> 1) & 1, d["aln_start"], d["aln_end"], R)
    for key, a in zip(TABLE_KEYS, got):
        assert a.dtype == np.int64 and np.array_equal(a, d[key]), key
    deltas = applybqsr.get_delta_qs(*got)
    for key, a in zip(DELTA_KEYS, deltas):
        assert np.array_equal(a, d[key]), key
    out = applybqsr.recalibrate_bam_arrays(d["seq"], d["qual"], d["rg"], d["flags"] & 1, (d["flags"] >> 1) & 1,
                                           got[0], *deltas)
    assert np.array_equal(out, d["outq"])


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_bam_file_objects_match_the_reference(standin, case):
    """bam_to_bqsr_covariates / recalibrate_bamread on pysam-shaped objects, small batches."""
    from kbbq.gatk import applybqsr, bqsr
    d = load_case(case)
    R = int(d["R"])
    standin.register_fasta(case + ".fa", {"chr1": d["ref"].tobytes().decode()})
    header = {"RG": [{"ID": "rg%d" % i, "PU": "unit.rg%d" % i} for i in range(R)]}
    reads = rebuild_reads(standin, d)
    got = bqsr.bam_to_bqsr_covariates(standin.AlignmentFile(reads=reads, header=header), case + ".fa",
                                      {"chr1": d["variable"].tolist()}, batch_reads=97)
    for key, a in zip(TABLE_KEYS, got):
        assert np.array_equal(a, d[key]), key
    rg_to_int = {"rg%d" % i: i for i in range(R)}
    deltas = [d[k] for k in DELTA_KEYS]
    for i in (0, 1, 2, 3, 17, len(reads) - 1):
        out = applybqsr.recalibrate_bamread(reads[i], d["meanq"], *deltas, rg_to_int)
        assert np.array_equal(out, d["outq"][i]), i
    report = bqsr.bam_to_report(standin.AlignmentFile(reads=reads, header=header), case + ".fa",
                                {"chr1": d["variable"].tolist()})
    assert list(report.tables[2].data.index) == ["unit.rg%d" % i for i in range(R) if d["rg_total"][i]]


@pytest.mark.gpu
def test_bam_device_entry_points_against_the_oracle(oracle_mod):
    """Larger seeded batch through the device-pointer entry points (kbbq.device), several batches."""
    import torch
    from kbbq import synth
    from kbbq.device import DeviceRecalibrator
    N, L, R = 60_000, 101, 3
    seq, qual, corr, rg, second = synth.synth_reads(99, 0, N, L, R)
    rng = np.random.default_rng(5)
    err = (seq != corr).astype(np.uint8)
    skip = (rng.random((N, L)) < 0.05).astype(np.uint8)
    flags = rng.integers(0, 4, size=N).astype(np.uint8)
    a0 = rng.integers(0, 8, size=N).astype(np.uint16) * (rng.random(N) < 0.4)
    a1 = (L - rng.integers(0, 8, size=N) * (rng.random(N) < 0.4)).astype(np.uint16)
    a0 = a0.astype(np.uint16)
    # a few bases outside the aligned window that the host did not mark: tallied at cycle 0 / dinuc AA (reference quirk)
    skip[::97, -1] = 0
    skip[::89, 0] = 0
    want = oracle_mod.build_tables_bam(seq, qual, err, skip, rg, flags, a0, a1, L, R)
    dev = lambda a: torch.from_numpy(a.view(np.int16) if a.dtype == np.uint16 else a).cuda()
    # fast = canonical form + the shared-memory kernels; not fast = the direct kernels (global atomics)
    for fast in (True, False):
        rec = DeviceRecalibrator(L, R, max_reads=0)
        for lo in range(0, N, 25_000):
            sl = slice(lo, min(N, lo + 25_000))
            rec.build_bam(dev(seq[sl]), dev(qual[sl]), dev(err[sl]), dev(skip[sl]), dev(rg[sl]), dev(flags[sl]),
                          dev(a0[sl]), dev(a1[sl]), fast=fast)
        tables = rec.covariate_arrays()
        rec.check_status()
        for key, a, b in zip(TABLE_KEYS[5:], tables[5:], want):
            assert np.array_equal(a, b), (key, fast)
        deltas = rec.delta_qs()
        out = torch.empty(N, L, dtype=torch.uint8, device="cuda")
        rec.apply_bam(dev(seq), dev(qual), out, dev(rg), dev(flags), fast=fast)
        rec.check_status()
        want_o = oracle_mod.apply_bam(seq, qual, rg, flags, L, R, tables[0], *deltas)
        assert np.array_equal(out.cpu().numpy().astype(np.int16), want_o), fast

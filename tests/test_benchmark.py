"""Calibration benchmark counts: kbbq.benchmark.calculate_q on the GPU.

Known answers from the reference's tests (tests/test_benchmark.py:62-69,116-123); the rest compares
the kernel with np.bincount on seeded inputs, bit-exact (integer counts)."""
import numpy as np
import pytest

from kbbq import benchmark, compare_reads


def _numpy_calculate_q(errors, quals):
    """The reference's formulation (kbbq/benchmark.py:76-91) on the host, as the checker."""
    numtotal = np.bincount(quals.reshape(-1))
    numerrs = np.bincount(quals[errors].reshape(-1), minlength=len(numtotal))
    nonzero = numtotal != 0
    actual = np.zeros(len(numtotal), dtype=int)
    actual[nonzero] = compare_reads.p_to_q(np.true_divide(numerrs[nonzero], numtotal[nonzero]))
    return actual, numtotal


def test_print_benchmark(capfd):
    actual = np.array([0, 20, 0, 42], dtype=int)
    total = np.array([0, 101, 1, 1], dtype=int)
    benchmark.print_benchmark(actual, 'test', total)
    assert capfd.readouterr().out == "1\t20\ttest\t101\n2\t0\ttest\t1\n3\t42\ttest\t1\n"


def test_bam_side_is_not_implemented():
    for name in ("benchmark", "benchmark_bam", "benchmark_fastq", "get_error_dict", "get_ref_dict"):
        with pytest.raises(NotImplementedError):
            getattr(benchmark, name)()


def test_readname_helpers():
    class R:
        name = "r7/2_RG:Z:foo"
        query_name = "r7"
        is_read2 = True
    assert benchmark.get_fastq_readname(R) == "r7/2"
    assert benchmark.get_bam_readname(R) == "r7/2"


@pytest.mark.gpu
def test_calculate_q_known_answer():
    errors = np.array([False, True, True] + [False] * 100)
    quals = np.array([3, 2, 1] + [1] * 100, dtype=int)
    a, t = benchmark.calculate_q(errors, quals)
    assert np.array_equal(a, np.array([0, 20, 0, 42], dtype=int))
    assert np.array_equal(t, np.array([0, 101, 1, 1], dtype=int))
    a, t = benchmark.calculate_q(np.zeros(0, bool), np.zeros(0, int))
    assert a.size == 0 and t.size == 0


@pytest.mark.gpu
@pytest.mark.parametrize("n,qmax", [(1, 42), (15, 42), (16, 42), (17, 5), (4099, 42), (1_000_003, 41),
                                    (300_000, 255), (2_000_000, 63), (2_000_000, 64)])
def test_calculate_q_matches_bincount(n, qmax):
    rng = np.random.default_rng(n + qmax)
    quals = rng.integers(0, qmax + 1, size=n)
    if n > 100:
        quals[rng.random(n) < 0.7] = min(qmax, 37)  # skewed like real data: most bases share a few values
    errors = rng.random(n) < 0.02
    a, t = benchmark.calculate_q(errors, quals)
    wa, wt = _numpy_calculate_q(errors, quals)
    assert np.array_equal(t, wt) and np.array_equal(a, wa)
    skips = rng.random(n) < 0.3
    a, t = benchmark.calculate_q_skips(errors, quals, skips)
    if np.any(~skips):
        wa, wt = _numpy_calculate_q(errors[~skips], quals[~skips])
        assert np.array_equal(t, wt) and np.array_equal(a, wa)
    else:
        assert t.size == 0
    # 2-D input, as the reference's benchmark_* build it from per-read arrays
    if n % 17 == 0:
        a2, t2 = benchmark.calculate_q(errors.reshape(-1, 17), quals.reshape(-1, 17))
        assert np.array_equal(t2, wt if False else _numpy_calculate_q(errors, quals)[1])


@pytest.mark.gpu
def test_calibration_counts_device_seq_vs_corr_and_misaligned():
    import torch
    from kbbq.device import calibration_counts, synth_reads
    N, L = 40_000, 151
    seq, qual, corr, rg, second = synth_reads(7, 0, N, L, 1)
    total, errs = calibration_counts(qual, seq=seq, corr=corr)
    q, e = qual.cpu().numpy().ravel(), (seq != corr).cpu().numpy().ravel()
    assert np.array_equal(total.cpu().numpy(), np.bincount(q, minlength=256))
    assert np.array_equal(errs.cpu().numpy(), np.bincount(q[e], minlength=256))
    # accumulates; misaligned views take the scalar kernel and must agree
    total2, errs2 = calibration_counts(qual.view(-1)[3:], seq=seq.view(-1)[3:], corr=corr.view(-1)[3:])
    assert np.array_equal(total2.cpu().numpy(), np.bincount(q[3:], minlength=256))
    assert np.array_equal(errs2.cpu().numpy(), np.bincount(q[3:][e[3:]], minlength=256))
    calibration_counts(qual, seq=seq, corr=corr, total=total, errs=errs)
    assert np.array_equal(total.cpu().numpy(), 2 * np.bincount(q, minlength=256))
    # recalibrated qualities score through the host API as well
    a, t = benchmark.calculate_q_reads(seq.cpu().numpy(), corr.cpu().numpy(), qual.cpu().numpy())
    wa, wt = _numpy_calculate_q(e, q.astype(np.int64))
    assert np.array_equal(a, wa) and np.array_equal(t, wt)
    assert torch.cuda.is_available()

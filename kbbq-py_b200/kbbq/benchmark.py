"""Mirror of the calibration benchmark of the reference's kbbq/benchmark.py.

`calculate_q` (kbbq/benchmark.py:76-91) -- the two bincounts over (quality, error) that score a
recalibrated read set -- runs on the GPU (kbbq_calibration_counts, csrc/calib.cuh);
`calculate_q_skips` fuses the `errors[~skips]`, `quals[~skips]` selection the reference does in
front of it (kbbq/benchmark.py:102-104,131-133).  `print_benchmark` and the read-name helpers are
host text.  The BAM / FASTA / VCF readers of the reference (get_ref_dict, get_var_sites,
get_bed_dict, get_full_skips, get_error_dict, benchmark_fastq, benchmark_bam, benchmark) need pysam
and are out of scope: they raise NotImplementedError.
"""
import numpy as np

from . import _native, compare_reads


def _counts_to_q(total, errs):
    n = int(np.flatnonzero(total)[-1]) + 1 if np.any(total) else 0  # np.bincount: length max(quals) + 1
    numtotal, numerrs = total[:n].copy(), errs[:n].copy()
    nonzero = numtotal != 0
    p = np.true_divide(numerrs[nonzero], numtotal[nonzero])
    actual_q = np.zeros(n, dtype=int)
    actual_q[nonzero] = compare_reads.p_to_q(p)
    return actual_q, numtotal


def _check_quals(quals):
    quals = np.asarray(quals)
    if quals.size and (quals.min() < 0 or quals.max() > 255):
        raise ValueError("quality scores must lie in 0..255")  # negative: ValueError in np.bincount as well
    return quals


def calculate_q(errors, quals):
    """Actual quality and number of bases per predicted quality (kbbq/benchmark.py:76-91).

    `errors` is a boolean array of the shape of `quals`.  Returns (actual_q, numtotal), both of length
    max(quals) + 1, actual_q = p_to_q(errors / total) where bases were seen and 0 elsewhere.
    """
    quals = _check_quals(quals)
    errors = np.asarray(errors)
    if errors.shape != quals.shape:
        raise IndexError("boolean index did not match indexed array")  # quals[errors] in the reference
    if quals.size == 0:
        return np.zeros(0, dtype=int), np.zeros(0, dtype=np.int64)
    total, errs = _native.calibration_counts_host(quals, err=errors.astype(bool))
    return _counts_to_q(total, errs)


def calculate_q_skips(errors, quals, skips):
    """calculate_q(errors[~skips], quals[~skips]) in one pass (kbbq/benchmark.py:102-104,131-133)."""
    quals = _check_quals(quals)
    errors, skips = np.asarray(errors), np.asarray(skips)
    if errors.shape != quals.shape or skips.shape != quals.shape:
        raise IndexError("boolean index did not match indexed array")
    if quals.size == 0:
        return np.zeros(0, dtype=int), np.zeros(0, dtype=np.int64)
    total, errs = _native.calibration_counts_host(quals, err=errors.astype(bool), skip=skips.astype(bool))
    return _counts_to_q(total, errs)


def calculate_q_reads(seq, corr, quals, skips=None):
    """The same counts with the errors taken as seq != corr (find_corrected_sites,
    kbbq/recalibrate.py:13-20) on packed u8 arrays: scores a recalibrated FASTQ against its
    corrected twin without materialising the error mask."""
    quals = _check_quals(quals)
    if quals.size == 0:
        return np.zeros(0, dtype=int), np.zeros(0, dtype=np.int64)
    total, errs = _native.calibration_counts_host(quals, seq=seq, corr=corr,
                                                  skip=None if skips is None else np.asarray(skips).astype(bool))
    return _counts_to_q(total, errs)


def get_bam_readname(read):
    """kbbq/benchmark.py:41-48."""
    return read.query_name + ("/2" if read.is_read2 else "/1")


def get_fastq_readname(read):
    """kbbq/benchmark.py:50-55."""
    return read.name.split(sep='_')[0]


def print_benchmark(actual_q, label, nbases):
    """Tab separated (predicted q, actual q, label, bases) rows, no header (kbbq/benchmark.py:130-145)."""
    actual_q, nbases = np.asarray(actual_q), np.asarray(nbases)
    nonzero = nbases != 0
    for pq, aq, nb in zip(np.arange(len(actual_q))[nonzero], actual_q[nonzero], nbases[nonzero]):
        print(pq, aq, label, nb, sep="\t")


def _needs_pysam(name):
    def f(*args, **kwargs):
        raise NotImplementedError("kbbq.benchmark.%s reads BAM / FASTA / VCF through pysam; the B200 path covers "
                                  "the counting step (calculate_q*) only" % name)
    f.__name__ = name
    return f


for _n in ("get_ref_dict", "get_var_sites", "get_bed_dict", "get_full_skips", "get_error_dict", "benchmark_fastq",
           "get_bamread_quals", "benchmark_bam", "benchmark"):
    globals()[_n] = _needs_pysam(_n)
del _n

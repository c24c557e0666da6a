"""Mirror of the reference's kbbq/recalibrate.py -- the driver of the FASTQ recalibration path.

Same functions, signatures, return order and exceptions; the per-read Python loops of the
reference (kbbq/recalibrate.py:56-119 and :141-156) become one packed batch handed to the CUDA
kernels through the C ABI (include/kbbq_b200.h).
"""
import sys

import numpy as np

from . import _native
from . import compare_reads as utils
from .batch import ReadBatch
from .gatk import applybqsr


def find_corrected_sites(uncorr_read, corr_read):
    """reference: kbbq/recalibrate.py:13-20 -- raw character comparison of one read pair."""
    assert corr_read.name.startswith(uncorr_read.name)
    uncorr_seq = np.array(list(uncorr_read.sequence), dtype=np.str_)
    corr_seq = np.array(list(corr_read.sequence), dtype=np.str_)
    return (uncorr_seq != corr_seq)


def _tables_from_batch(batch, minscore=6):
    if batch.N == 0:
        # the reference returns zero-length arrays when there are no reads (kbbq/recalibrate.py:45-54)
        z = np.zeros(0, dtype=np.int64)
        return (utils.p_to_q(z.astype(float)), z.copy(), z.copy(), np.zeros((0, 43), np.int64),
                np.zeros((0, 43), np.int64), np.zeros((0, 43, 0), np.int64), np.zeros((0, 43, 0), np.int64),
                np.zeros((0, 43, 16), np.int64), np.zeros((0, 43, 16), np.int64))
    pe, pt, de, dt = _native.build_host(batch.seq, batch.qual, batch.corr, batch.rg, batch.second,
                                        batch.L, batch.R, minscore)
    meanq, rg_e, rg_t, q_e, q_t = _native.marginals_host(pe, pt)
    return meanq, rg_e, rg_t, q_e, q_t, pe, pt, de, dt


def fastq_to_covariate_arrays(fastq, infer_rg=False, minscore=6, maxscore=42):
    """reference: kbbq/recalibrate.py:22-121.

    -> (meanq, rg_errs, rg_total, q_errs, q_total, pos_errs, pos_total, dinuc_errs, dinuc_total),
    fresh int64 arrays of shapes [R], [R], [R], [R,43], [R,43], [R,43,2L], [R,43,2L], [R,43,16] x2.
    """
    if maxscore != 42:
        raise NotImplementedError("only maxscore = 42 is supported")
    batch = ReadBatch.from_fastq(fastq, infer_rg)
    return _tables_from_batch(batch, minscore)


STREAM_ABOVE_BYTES = 4 << 30      # larger FASTQ files go through the device in batches (host memory stays bounded)
STREAM_BATCH_READS = 8_000_000


def _stdout_fd():
    """Descriptor of sys.stdout, or None when it is not a real file (a StringIO, a capture object)."""
    try:
        fd = sys.stdout.fileno()
    except (AttributeError, OSError, ValueError):
        return None
    sys.stdout.flush()
    return fd


def _recalibrate_fastq_streamed(fastq, infer_rg, batch_reads):
    """The two passes of kbbq/recalibrate.py:123-156 over batches of `batch_reads` reads: pass 1 packs a
    batch from the indexed files, builds its tables on the GPU and adds them up (integer tables are
    additive), the model runs once, pass 2 packs each batch again, applies and prints it.  Host memory
    holds one batch (plus 8 bytes of index per read); the result is the one of the single-batch path."""
    from . import fastx
    reads, fixed = fastx.NativeFastq(fastq[0]), fastx.NativeFastq(fastq[1])
    n = min(reads.N, fixed.N)
    reads.check_names(fixed, n)
    if n == 0:
        return
    if reads.L < 0 or fixed.L != reads.L:
        raise ValueError("operands could not be broadcast together: reads of unequal length")
    rg, second, keys = reads.infer(infer_rg)
    L, R = reads.L, max(1, len(keys))
    from concurrent.futures import ThreadPoolExecutor
    starts = list(range(0, n, batch_reads))

    def pack(lo, with_corr):
        m = min(batch_reads, n - lo)
        seq, qual = reads.pack(lo, m)
        return lo, m, seq, qual, (fixed.pack(lo, m)[0] if with_corr else None)

    def batches(with_corr):
        # the next batch is tokenised (native code, GIL released) while the GPU works on this one
        with ThreadPoolExecutor(1) as pool:
            nxt = pool.submit(pack, starts[0], with_corr)
            for i in range(len(starts)):
                cur = nxt.result()
                if i + 1 < len(starts):
                    nxt = pool.submit(pack, starts[i + 1], with_corr)
                yield cur

    tables = None
    for lo, m, seq, qual, corr in batches(True):
        part = _native.build_host(seq, qual, corr, rg[lo:lo + m], second[lo:lo + m], L, R, 6)
        tables = part if tables is None else tuple(a + b for a, b in zip(tables, part))
    fixed.close()
    meanq, rg_e, rg_t, q_e, q_t = _native.marginals_host(tables[0], tables[1])
    deltas = _native.get_delta_qs_host(meanq, rg_e, rg_t, q_e, q_t, *tables)
    fd = _stdout_fd()
    for lo, m, seq, qual, _ in batches(False):
        out = _native.apply_host(seq, qual, rg[lo:lo + m], second[lo:lo + m], L, R, meanq, *deltas)
        if fd is not None:
            reads.write(fd, out, lo, m)
        else:
            txt = (out + np.uint8(33)).astype(np.uint8)
            sys.stdout.write(''.join('@%s\n%s\n+\n%s\n' % (reads.name(lo + i), seq[i].tobytes().decode(),
                                                          txt[i].tobytes().decode('latin-1')) for i in range(m)))
    reads.close()


def recalibrate_fastq(fastq, infer_rg=False):
    """Recalibrate fastq[0] given its corrected twin fastq[1]; FASTQ to stdout
    (reference: kbbq/recalibrate.py:123-156: name without comment, sequence, '+', chr(q + 33))."""
    import os
    batch_reads = int(os.environ.get("KBBQ_BATCH_READS", "0"))
    if batch_reads <= 0 and os.path.getsize(fastq[0]) > STREAM_ABOVE_BYTES:
        batch_reads = STREAM_BATCH_READS
    if batch_reads > 0:
        return _recalibrate_fastq_streamed(fastq, infer_rg, batch_reads)
    batch = ReadBatch.from_fastq(fastq, infer_rg)
    if batch.N == 0:
        return
    out = _native.recalibrate_host(batch.seq, batch.qual, batch.corr, batch.rg, batch.second,
                                   batch.L, batch.R, 6)
    # native formatter straight to the stdout descriptor when there is one; text fallback otherwise
    # (e.g. sys.stdout replaced by an in-memory stream)
    w = sys.stdout
    try:
        fd = w.fileno()
    except (AttributeError, OSError, ValueError):
        fd = None
    if fd is not None and batch.source is not None:
        w.flush()
        batch.source.write(fd, out, 0, batch.N)
        return
    qual_txt = (out + np.uint8(33)).astype(np.uint8)
    seq = batch.seq
    chunks = []
    for i, name in enumerate(batch.names):
        chunks.append('@%s\n%s\n+\n%s\n' % (name, seq[i].tobytes().decode(), qual_txt[i].tobytes().decode('latin-1')))
        if len(chunks) >= 4096:
            w.write(''.join(chunks))
            chunks = []
    w.write(''.join(chunks))


def recalibrate_bam(bam, use_oq=False, set_oq=False):
    """Not implemented in the reference either (kbbq/recalibrate.py:158-164)."""
    raise NotImplementedError('Recalibrating a bam is not yet implemented. \
        Try converting your BAM to a FASTQ file with the samtools fastq command.')


def recalibrate(bam, fastq, infer_rg=False, use_oq=False, set_oq=False, gatkreport=None):
    """reference: kbbq/recalibrate.py:166-174."""
    if gatkreport is not None:
        raise NotImplementedError('GATKreport reading / creation is not yet supported.')
    elif bam is not None:
        recalibrate_bam(bam, use_oq, set_oq)
    elif fastq is not None:
        recalibrate_fastq(fastq, infer_rg=infer_rg)
    else:
        raise ValueError("A BAM or FASTQ file should be provided for recalibration.")

"""Recalibration-report creation: mirror of the report half of the
reference's kbbq/gatk/bqsr.py (`quantize` :214-225, `vectors_to_report` :227-366).

`vectors_to_report` turns the nine count / quality vectors of the table build into a
:class:`kbbq.recaltable.RecalibrationReport` whose text is byte-identical to the reference's
(pinned by tests/golden/report_*.txt).  The EmpiricalQuality columns are MAP qualities: they come
from the delta-Q kernels of the C ABI (kbbq_delta_q, kbbq_posterior_q_real), one launch per table
over every cell at once; rows are then selected and ordered with NumPy (the reference builds
pandas frames over all cells, concatenates and re-sorts them through a MultiIndex).

The BAM half keeps the reference's function names: the host walks the
records (it needs nothing from pysam but the objects the caller hands in; only reading the FASTA
imports pysam), the tally runs on the GPU (kbbq_build_bam).
"""
import numpy as np
import pandas as pd

from .. import _native
from .. import compare_reads as utils
from .. import recaltable

# the argument table GATK writes; only the defaults are reproduced (kbbq/gatk/bqsr.py:264-282)
REPORT_ARGUMENTS = (
    ('binary_tag_name', 'null'),
    ('covariate', 'ReadGroupCovariate,QualityScoreCovariate,ContextCovariate,CycleCovariate'),
    ('default_platform', 'null'),
    ('deletions_default_quality', '45'),
    ('force_platform', 'null'),
    ('indels_context_size', '3'),
    ('insertions_default_quality', '45'),
    ('low_quality_tail', '2'),
    ('maximum_cycle_value', '500'),
    ('mismatches_context_size', '2'),
    ('mismatches_default_quality', '-1'),
    ('no_standard_covs', 'false'),
    ('quantizing_levels', '16'),
    ('recalibration_report', 'null'),
    ('run_without_dbsnp', 'false'),
    ('solid_nocall_strategy', 'THROW_EXCEPTION'),
    ('solid_recal_mode', 'SET_Q_ZERO'),
)


def quantize(q_errs, q_total, maxscore=93):
    """Identity quantisation map with unobserved qualities sent to maxscore (kbbq/gatk/bqsr.py:214-225)."""
    qt = np.sum(np.asarray(q_total), axis=0)
    quantizer = np.arange(maxscore + 1)
    quantizer[:qt.shape[0]][qt == 0] = maxscore
    quantizer[qt.shape[0]:] = maxscore
    return quantizer


def _estimated_q(q_total, global_total):
    """EstimatedQReported of every read group: -10 log10 of the mean reported error probability,
    rounded to five decimals, 0 where the group was never observed (kbbq/gatk/bqsr.py:289-290)."""
    nq = q_total.shape[1]
    with np.errstate(divide='ignore', invalid='ignore'):
        expected = np.sum(utils.q_to_p(np.arange(nq)) * q_total, axis=1)
        est = (-10.0 * np.log10(expected / global_total)).round(decimals=5).astype(np.float64)
    est[np.isnan(est)] = 0
    return est


def vectors_to_report(meanq, global_errs, global_total, q_errs, q_total, pos_errs, pos_total,
                      dinuc_errs, dinuc_total, rg_order, maxscore=42):
    """The recalibration vectors as a RecalibrationReport (kbbq/gatk/bqsr.py:227-366).

    Shapes: [rg], [rg, q], [rg, q, 2 * cycles], [rg, q, 16]; `rg_order` names the read groups.
    `meanq` is accepted for signature compatibility; like the reference, the report derives
    EstimatedQReported from q_total.
    """
    global_errs, global_total = np.asarray(global_errs, np.int64), np.asarray(global_total, np.int64)
    q_errs, q_total = np.asarray(q_errs, np.int64), np.asarray(q_total, np.int64)
    pos_errs, pos_total = np.asarray(pos_errs, np.int64), np.asarray(pos_total, np.int64)
    dinuc_errs, dinuc_total = np.asarray(dinuc_errs, np.int64), np.asarray(dinuc_total, np.int64)
    rgs = np.asarray(list(rg_order), dtype=str)
    R, nq = q_total.shape
    assert rgs.shape[0] == R and pos_total.shape[:2] == (R, nq) and dinuc_total.shape[:2] == (R, nq)
    if nq > 43:
        raise IndexError("quality axis longer than 43")

    argtable = pd.DataFrame({'Argument': np.array([a for a, _ in REPORT_ARGUMENTS], dtype=object),
                             'Value': np.array([v for _, v in REPORT_ARGUMENTS], dtype=object)})

    counts = np.zeros(94)
    counts[:nq] = np.sum(q_total, axis=0)
    quanttable = pd.DataFrame({'QualityScore': np.arange(94), 'Count': counts,
                               'QuantizedScore': quantize(q_errs, q_total)})

    # read groups: real-valued prior
    est = _estimated_q(q_total, global_total)
    posterior = _native.posterior_q_real_host(est, global_errs, global_total)
    keep = global_total != 0
    rgtable = pd.DataFrame({'ReadGroup': rgs[keep].astype(object), 'EventType': 'M',
                            'EmpiricalQuality': ((posterior - est) + est).astype(np.float64)[keep],
                            'EstimatedQReported': est[keep], 'Observations': global_total[keep],
                            'Errors': global_errs.astype(np.float64)[keep]},
                           columns=['ReadGroup', 'EventType', 'EmpiricalQuality', 'EstimatedQReported',
                                    'Observations', 'Errors'])

    # reported quality: the prior is the quality itself
    qs = np.broadcast_to(np.arange(nq), (R, nq))
    keep = (q_total != 0).ravel()
    qsel = qs.ravel()[keep]
    emp = utils.gatk_delta_q(qsel, q_errs.ravel()[keep], q_total.ravel()[keep]) + qsel if qsel.size else qsel
    qualtable = pd.DataFrame({'ReadGroup': np.repeat(rgs, nq)[keep].astype(object), 'QualityScore': qsel,
                              'EventType': 'M', 'EmpiricalQuality': emp.astype(np.float64),
                              'Observations': q_total.ravel()[keep], 'Errors': q_errs.ravel()[keep].astype(np.float64)},
                             columns=['ReadGroup', 'QualityScore', 'EventType', 'EmpiricalQuality', 'Observations',
                                      'Errors'])

    # covariates: context (dinucleotide) and cycle cells that were observed
    ncyc2 = pos_total.shape[2]
    ncycles = ncyc2 // 2
    cycle_values = np.concatenate([np.arange(ncycles) + 1, -(np.arange(ncycles) + 1)[::-1]]).astype(np.int64)
    parts = []
    for name, values, tot, err in (('Context', np.array(utils.Dinucleotide.dinucs, dtype=str), dinuc_total, dinuc_errs),
                                   ('Cycle', cycle_values.astype(str), pos_total, pos_errs)):
        r, q, c = np.nonzero(tot)
        parts.append((r, q, values[c], np.full(r.shape[0], name), tot[r, q, c], err[r, q, c]))
    r = np.concatenate([p[0] for p in parts])
    q = np.concatenate([p[1] for p in parts])
    width = max(parts[0][2].dtype.itemsize, parts[1][2].dtype.itemsize) // 4
    value = np.concatenate([p[2].astype('U%d' % max(width, 1)) for p in parts])
    name = np.concatenate([p[3] for p in parts])
    tot = np.concatenate([p[4] for p in parts])
    err = np.concatenate([p[5] for p in parts])
    # GATK's row order: read group, quality, covariate value (as text), covariate name
    rank = np.empty(R, np.int64)
    rank[np.argsort(rgs, kind='stable')] = np.arange(R)
    order = np.lexsort((name, value, q, rank[r]))
    r, q, value, name, tot, err = r[order], q[order], value[order], name[order], tot[order], err[order]
    emp = (utils.gatk_delta_q(q, err, tot) + q) if q.size else q
    covtable = pd.DataFrame({'ReadGroup': rgs[r].astype(object), 'QualityScore': q.astype(np.int64),
                             'CovariateValue': value.astype(object), 'CovariateName': name.astype(object),
                             'EventType': 'M', 'EmpiricalQuality': emp.astype(np.float64), 'Observations': tot,
                             'Errors': err.astype(np.float64)},
                            columns=['ReadGroup', 'QualityScore', 'CovariateValue', 'CovariateName', 'EventType',
                                     'EmpiricalQuality', 'Observations', 'Errors'])

    descriptions = ['Recalibration argument collection values used in this run', 'Quality quantization map', '', '', '']
    tables = [recaltable.GATKTable(t, d, f) for t, d, f in
              zip(recaltable.RecalibrationReport.TITLES, descriptions, [argtable, quanttable, rgtable, qualtable, covtable])]
    return recaltable.RecalibrationReport(tables)


# ---- BAM side -------------------------------------------------------
# Per-read covariate helpers keep the reference's names and results (they are API surface; the
# batched kernel kbbq_build_bam fuses the same index arithmetic).  Reads are anything with pysam's
# AlignedSegment attributes.

def bamread_bqsr_cycle(read):
    """Cycle of every base, 0 outside the aligned part (reference: kbbq/gatk/bqsr.py:23-31)."""
    full = np.zeros(read.query_length, dtype=int)
    lo, hi = read.query_alignment_start, read.query_alignment_end
    c = np.arange(hi - lo)
    if read.is_reverse:
        c = c[::-1]
    full[lo:hi] = np.negative(c + 1) if read.is_read2 else c
    return full


def _strand_dinuc(seq, quals, reverse, minscore):
    """generic_dinuc_covariate of a window; on the reverse strand of its reverse complement, flipped
    back to stored order (reference: kbbq/gatk/bqsr.py:38-49, kbbq/gatk/applybqsr.py:55-62)."""
    if not reverse:
        return utils.generic_dinuc_covariate(np.array(list(seq), dtype='U1'), quals, minscore).copy()
    rc = [utils.Dinucleotide.complement.get(x, 'N') for x in reversed(seq)]
    return utils.generic_dinuc_covariate(np.array(rc, dtype='U1').reshape(len(rc)), quals[::-1], minscore)[::-1].copy()


def bamread_bqsr_dinuc(read, use_oq=True, minscore=6):
    """Dinucleotide of every base inside the aligned part, 0 outside (reference: kbbq/gatk/bqsr.py:33-50)."""
    lo, hi = read.query_alignment_start, read.query_alignment_end
    quals = utils.bamread_get_oq(read) if use_oq else np.array(read.query_qualities, dtype=int)
    full = np.zeros(read.query_length, dtype=int)
    if hi > lo:
        full[lo:hi] = _strand_dinuc(read.query_sequence[lo:hi], quals[lo:hi], read.is_reverse, minscore)
    return full


def bamread_adaptor_boundary(read):
    """Reference position where the adaptor starts, or None (reference: kbbq/gatk/bqsr.py:131-156)."""
    if (read.tlen == 0 or not read.is_paired or read.is_unmapped or read.mate_is_unmapped or
            read.is_reverse == read.mate_is_reverse):
        return None
    if read.is_reverse:
        return read.next_reference_start - 1 if (read.reference_end - 1) > read.next_reference_start else None
    return read.reference_start + abs(read.tlen) if read.reference_start <= (read.next_reference_start + read.tlen) else None


def trim_bamread(read):
    """Bases past the adaptor boundary, to be skipped (reference: kbbq/gatk/bqsr.py:158-212)."""
    skips = np.zeros(len(read.query_qualities), dtype=bool)
    boundary = bamread_adaptor_boundary(read)
    if boundary is None:
        return skips
    pairs = read.get_aligned_pairs()
    if read.is_reverse:
        if boundary >= read.reference_start:
            cut, reached = 0, False
            for readidx, refidx in reversed(pairs):  # first read base at or left of the boundary, from the right
                reached = reached or (refidx is not None and refidx <= boundary)
                if reached and readidx is not None:
                    cut = readidx + 1
                    break
            skips[:cut] = True
    elif boundary <= read.reference_end - 1:
        cut, reached = len(skips), False
        for readidx, refidx in pairs:
            reached = reached or (refidx is not None and refidx >= boundary)
            if reached and readidx is not None:
                cut = readidx
                break
        skips[cut:] = True
    return skips


def bam_arrays_to_bqsr_covariates(seq, qual, errors, skips, rg, is_read2, is_reverse, aln_start, aln_end, nrgs,
                                  minscore=6):
    """The tally of bam_to_bqsr_covariates (kbbq/gatk/bqsr.py:86-123) on reads already unpacked into
    arrays: seq (bytes), qual (OQ), errors, skips [N, L]; rg, is_read2, is_reverse, aln_start, aln_end [N].
    Runs on the GPU (kbbq_build_bam + kbbq_marginals).  -> the nine vectors, as fastq_to_covariate_arrays."""
    seq = np.asarray(seq)
    if seq.dtype.kind in 'US':
        seq = np.frombuffer(np.ascontiguousarray(seq).astype('S1').tobytes(), np.uint8).reshape(seq.shape)
    N, L = seq.shape
    flags = (np.asarray(is_read2).astype(np.uint8) & 1) | ((np.asarray(is_reverse).astype(np.uint8) & 1) << 1)
    pe, pt, de, dt = _native.build_bam_host(seq, qual, np.asarray(errors).astype(bool), np.asarray(skips).astype(bool),
                                            rg, flags, aln_start, aln_end, L, nrgs, minscore)
    meanq, rg_e, rg_t, q_e, q_t = _native.marginals_host(pe, pt)
    return meanq, rg_e, rg_t, q_e, q_t, pe, pt, de, dt


def _load_reference(fastafilename):
    try:
        import pysam
    except ImportError:
        raise NotImplementedError("reading %s needs pysam.FastaFile; pass arrays to bam_arrays_to_bqsr_covariates "
                                  "instead" % fastafilename)
    fasta = pysam.FastaFile(fastafilename)
    return {chrom: np.frombuffer(fasta.fetch(reference=chrom).encode(), dtype=np.uint8) for chrom in fasta.references}


def bam_to_bqsr_covariates(bamfileobj, fastafilename, var_pos, minscore=6, maxscore=42, batch_reads=1 << 20):
    """Covariate arrays of a BAM file (reference: kbbq/gatk/bqsr.py:52-123).

    The host walks the records (CIGAR against the reference, known variant sites, adaptor trimming:
    compare_reads.find_read_errors, trim_bamread) and packs them; every `batch_reads` reads the tally
    runs on the GPU and adds into the tables.  As in the reference the tables are sized by the first
    read; shorter reads are padded with skipped bases, a longer one is an IndexError.
    """
    if maxscore != 42:
        raise NotImplementedError("only maxscore = 42 is supported")
    rg_to_pu = utils.get_rg_to_pu(bamfileobj)
    rg_to_int = dict(zip(rg_to_pu, range(len(rg_to_pu))))
    nrgs = len(rg_to_pu)
    ref = _load_reference(fastafilename)
    fullskips = {chrom: np.zeros(len(ref[chrom]), dtype=bool) for chrom in ref}
    for chrom in fullskips:
        fullskips[chrom][np.array(var_pos[chrom], dtype=int)] = True
    tables, L, rows = None, None, []

    def flush():
        nonlocal tables
        if not rows:
            return
        n = len(rows)
        seq, qual = np.full((n, L), ord('N'), np.uint8), np.zeros((n, L), np.uint8)
        err, skip = np.zeros((n, L), np.uint8), np.ones((n, L), np.uint8)
        rg, flags = np.zeros(n, np.uint16), np.zeros(n, np.uint8)
        a0, a1 = np.zeros(n, np.uint16), np.zeros(n, np.uint16)
        for i, (s, q, e, k, g, f, lo, hi) in enumerate(rows):
            m = s.size
            seq[i, :m], qual[i, :m], err[i, :m], skip[i, :m] = s, q, e, k
            rg[i], flags[i], a0[i], a1[i] = g, f, lo, hi
        tables = _native.build_bam_host(seq, qual, err, skip, rg, flags, a0, a1, L, nrgs, minscore, tables=tables)
        rows.clear()

    for read in bamfileobj:
        s = np.frombuffer(read.query_sequence.encode(), dtype=np.uint8)
        if L is None:
            L = len(read.query_qualities)
        if s.size > L:
            raise IndexError("read %s is longer than the first read (%d > %d)" % (read.query_name, s.size, L))
        e, k = utils.find_read_errors(read, ref, fullskips)
        k = np.logical_or(k, trim_bamread(read))
        q = utils.bamread_get_oq(read)
        if np.any(q > 255) or np.any(q < 0):
            raise IndexError("quality out of range")
        rows.append((s, q.astype(np.uint8), e, k, rg_to_int[read.get_tag('RG')],
                     (1 if read.is_read2 else 0) | (2 if read.is_reverse else 0),
                     read.query_alignment_start, read.query_alignment_end))
        if len(rows) >= batch_reads:
            flush()
    flush()
    if tables is None:
        raise StopIteration  # the reference's next(bamfileobj) on an empty file (:70)
    pe, pt, de, dt = tables
    meanq, rg_e, rg_t, q_e, q_t = _native.marginals_host(pe, pt)
    return meanq, rg_e, rg_t, q_e, q_t, pe, pt, de, dt


def bam_to_report(bamfileobj, fastafilename, var_pos):
    """reference: kbbq/gatk/bqsr.py:368-371."""
    rgs = list(utils.get_rg_to_pu(bamfileobj).values())
    vectors = bam_to_bqsr_covariates(bamfileobj, fastafilename, var_pos)
    return vectors_to_report(*vectors, rgs)

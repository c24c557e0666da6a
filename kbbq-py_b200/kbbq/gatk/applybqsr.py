"""Mirror of the hot-path part of the reference's kbbq/gatk/applybqsr.py: get_delta_qs (:80-103)
and the report reader table_to_vectors (:14-44).

BAM side: bamread_cycle_covariates, bamread_dinuc_covariates,
recalibrate_bamread (:46-78) for one read, recalibrate_bam_arrays for a packed batch; the gather
runs on the GPU (kbbq_apply_bam).
"""
import numpy as np

from .. import _native
from .. import compare_reads as utils


def table_to_vectors(table, rg_order, maxscore=42):
    """The nine recalibration vectors of a RecalibrationReport (kbbq/gatk/applybqsr.py:14-44).

    -> (meanq float64 [rg] = EstimatedQReported, global_errs, global_total [rg], q_errs, q_total
    [rg, maxscore + 1], pos_errs, pos_total [rg, maxscore + 1, 2 * seqlen], dinuc_errs, dinuc_total
    [rg, maxscore + 1, 16]), counts int64, cells the report does not list are 0.  seqlen is the
    largest cycle in the report; cycle c > 0 sits at c - 1 and cycle -c at 2 * seqlen - c, the layout
    of the table build.  Rows of read groups outside `rg_order` or qualities above maxscore are
    ignored.  The reference does this with pandas reindex / fillna over the full index product;
    here the listed rows are scattered into zero arrays.
    """
    rg_order = list(rg_order)
    R, nq = len(rg_order), maxscore + 1
    rg_index = {str(rg): i for i, rg in enumerate(rg_order)}

    def rg_codes(values):
        return np.array([rg_index.get(str(v), -1) for v in values], dtype=np.int64)

    t2 = table.tables[2].data.reset_index()
    meanq = np.full(R, np.nan, dtype=np.float64)
    global_errs, global_total = np.zeros(R, np.int64), np.zeros(R, np.int64)
    g = rg_codes(t2['ReadGroup'].to_numpy())
    missing = sorted(set(range(R)) - set(g[g >= 0].tolist()))
    if missing:
        # the reference fails converting the NaN of a missing row to int64 (:19)
        raise ValueError("read group %r is not in the report" % (rg_order[missing[0]],))
    ok = g >= 0
    meanq[g[ok]] = t2['EstimatedQReported'].to_numpy(dtype=np.float64)[ok]
    global_errs[g[ok]] = t2['Errors'].to_numpy(dtype=np.float64)[ok].astype(np.int64)
    global_total[g[ok]] = t2['Observations'].to_numpy(dtype=np.int64)[ok]

    t3 = table.tables[3].data.reset_index()
    q_errs, q_total = np.zeros((R, nq), np.int64), np.zeros((R, nq), np.int64)
    g, q = rg_codes(t3['ReadGroup'].to_numpy()), t3['QualityScore'].to_numpy(dtype=np.int64)
    ok = (g >= 0) & (q >= 0) & (q < nq)
    q_errs[g[ok], q[ok]] = t3['Errors'].to_numpy(dtype=np.float64)[ok].astype(np.int64)
    q_total[g[ok], q[ok]] = t3['Observations'].to_numpy(dtype=np.int64)[ok]

    t4 = table.tables[4].data.reset_index()
    g, q = rg_codes(t4['ReadGroup'].to_numpy()), t4['QualityScore'].to_numpy(dtype=np.int64)
    name = np.asarray(t4['CovariateName'].to_numpy(), dtype=str)
    value = np.asarray(t4['CovariateValue'].to_numpy(), dtype=str)
    errs = t4['Errors'].to_numpy(dtype=np.float64).astype(np.int64)
    obs = t4['Observations'].to_numpy(dtype=np.int64)
    ok = (g >= 0) & (q >= 0) & (q < nq)

    cyc = ok & (name == 'Cycle')
    if not np.any(cyc):
        raise KeyError('Cycle')  # .loc[..., 'Cycle'] in the reference (:28)
    c = value[cyc].astype(np.int64)
    # the reference takes the largest cycle (:30), which only works when read-1 cycles reach the full
    # length; the largest magnitude is the same number there and also right for read-2-only reports
    seqlen = int(np.abs(c).max())
    if np.any(c == 0):
        raise ValueError("malformed cycle values in the report")
    pos_errs, pos_total = np.zeros((R, nq, 2 * seqlen), np.int64), np.zeros((R, nq, 2 * seqlen), np.int64)
    slot = np.where(c > 0, c - 1, 2 * seqlen + c)
    pos_errs[g[cyc], q[cyc], slot] = errs[cyc]
    pos_total[g[cyc], q[cyc], slot] = obs[cyc]

    dinuc_to_int = utils.Dinucleotide.dinuc_to_int
    ndin = len(dinuc_to_int)
    dinuc_errs, dinuc_total = np.zeros((R, nq, ndin), np.int64), np.zeros((R, nq, ndin), np.int64)
    ctx = ok & (name == 'Context')
    d = np.array([dinuc_to_int.get(v, -1) for v in value[ctx]], dtype=np.int64)
    known = d >= 0  # contexts with an N are not part of the index product the reference reindexes to (:38)
    dinuc_errs[g[ctx][known], q[ctx][known], d[known]] = errs[ctx][known]
    dinuc_total[g[ctx][known], q[ctx][known], d[known]] = obs[ctx][known]

    return meanq, global_errs, global_total, q_errs, q_total, pos_errs, pos_total, dinuc_errs, dinuc_total


def get_delta_qs(meanq, rg_errs, rg_total, q_errs, q_total, pos_errs, pos_total, dinuc_errs, dinuc_total):
    """Hierarchical delta-Q tables; runs on the GPU (kbbq_get_delta_qs).

    Shapes as in the reference: [rg], [rg, q], [rg, q, covariate]; returns
    (rgdeltaq [rg], qscoredeltaq [rg, q], positiondeltaq [rg, q, cycle], dinucdq [rg, q, dinuc + 1])
    as fresh int arrays, the last dinuc column being the zero pad that index -1 gathers.
    """
    meanq = np.asarray(meanq)
    q_total, pos_total, dinuc_total = np.asarray(q_total), np.asarray(pos_total), np.asarray(dinuc_total)
    assert meanq.ndim == 1 and q_total.ndim == 2 and pos_total.ndim == 3 and dinuc_total.ndim == 3
    assert q_total.shape[0] == meanq.shape[0]
    assert pos_total.shape[:2] == q_total.shape and dinuc_total.shape[:2] == q_total.shape
    return _native.get_delta_qs_host(meanq, rg_errs, rg_total, q_errs, q_total, pos_errs, pos_total,
                                     dinuc_errs, dinuc_total)


# ---- BAM side ------------------------------------------------------------------------------------

def bamread_cycle_covariates(read):
    """Whole-read cycles, flipped on the reverse strand (reference: kbbq/gatk/applybqsr.py:46-50)."""
    cycle = utils.generic_cycle_covariate(read.query_length, read.is_read2)
    return cycle[::-1].copy() if read.is_reverse else cycle


def bamread_dinuc_covariates(read, use_oq=True, minscore=6):
    """Whole-read dinucleotides, from the reverse complement on the reverse strand
    (reference: kbbq/gatk/applybqsr.py:52-63)."""
    from . import bqsr
    quals = utils.bamread_get_oq(read) if use_oq else np.array(read.query_qualities, dtype=int)
    return bqsr._strand_dinuc(read.query_sequence, quals, read.is_reverse, minscore)


def recalibrate_bam_arrays(seq, qual, rg, is_read2, is_reverse, meanq, globaldeltaq, qscoredeltaq, positiondeltaq,
                           dinucdeltaq, minscore=6):
    """recalibrate_bamread over a packed batch: seq (bytes), qual [N, L]; rg, is_read2, is_reverse [N].
    -> int array [N, L].  Runs on the GPU (kbbq_apply_bam)."""
    seq = np.ascontiguousarray(seq, dtype=np.uint8)
    N, L = seq.shape
    flags = (np.asarray(is_read2).astype(np.uint8) & 1) | ((np.asarray(is_reverse).astype(np.uint8) & 1) << 1)
    R = np.asarray(meanq).shape[0]
    out = _native.apply_bam_host(seq, qual, rg, flags, L, R, meanq, globaldeltaq, qscoredeltaq, positiondeltaq,
                                 dinucdeltaq, minscore=minscore)
    return out.astype(np.int8).astype(int)  # the device keeps the low 8 bits of the sum


def recalibrate_bamread(read, meanq, globaldeltaq, qscoredeltaq, positiondeltaq, dinucdeltaq, rg_to_int,
                        use_oq=True, minscore=6):
    """Recalibrated qualities of ONE aligned read (reference: kbbq/gatk/applybqsr.py:65-78).

    Like the reference, the dinucleotide covariate is gated by the OQ qualities whatever `use_oq` says
    (:74 calls bamread_dinuc_covariates with its default); with use_oq = False the two quality arrays
    differ and the rare sites where they disagree about minscore are patched on the host.
    """
    original = utils.bamread_get_oq(read) if use_oq else np.array(read.query_qualities, dtype=int)
    if np.any(original < 0) or np.any(original > 255):
        raise IndexError("quality out of range")
    seq = np.frombuffer(read.query_sequence.encode(), dtype=np.uint8).reshape(1, -1)
    rg = rg_to_int[read.get_tag('RG')]
    meanq, globaldeltaq = np.atleast_1d(np.asarray(meanq)), np.atleast_1d(np.asarray(globaldeltaq))
    qscoredeltaq, positiondeltaq, dinucdeltaq = (np.asarray(a) for a in (qscoredeltaq, positiondeltaq, dinucdeltaq))
    if positiondeltaq.shape[2] != 2 * seq.shape[1]:
        raise IndexError("positiondeltaq cycle axis must have length 2 * len(read)")
    out = recalibrate_bam_arrays(seq, original.reshape(1, -1), np.zeros(1, np.uint16), [read.is_read2], [read.is_reverse],
                                 meanq[rg:rg + 1], globaldeltaq[rg:rg + 1], qscoredeltaq[rg:rg + 1],
                                 positiondeltaq[rg:rg + 1], dinucdeltaq[rg:rg + 1], minscore=minscore)[0]
    if not use_oq:
        # dinucleotides gated by OQ, everything else by the BAM qualities: redo the (few) affected sites
        din_oq = bamread_dinuc_covariates(read, True, minscore)
        din_q = bamread_dinuc_covariates(read, False, minscore)
        cyc = bamread_cycle_covariates(read)
        for i in np.flatnonzero((din_oq != din_q) & (original >= minscore)):
            q = original[i]
            out[i] = (meanq[rg] + globaldeltaq[rg] + qscoredeltaq[rg, q] + dinucdeltaq[rg, q, din_oq[i]] +
                      positiondeltaq[rg, q, cyc[i]])
    return out

"""Recalibration-report creation: mirror of the report half of the
reference's kbbq/gatk/bqsr.py (`quantize` :214-225, `vectors_to_report` :227-366).

`vectors_to_report` turns the nine count / quality vectors of the table build into a
:class:`kbbq.recaltable.RecalibrationReport` whose text is byte-identical to the reference's
(pinned by tests/golden/report_*.txt).  The EmpiricalQuality columns are MAP qualities: they come
from the delta-Q kernels of the C ABI (kbbq_delta_q, kbbq_posterior_q_real), one launch per table
over every cell at once; rows are then selected and ordered with NumPy (the reference builds
pandas frames over all cells, concatenates and re-sorts them through a MultiIndex).

The BAM half of the reference module (bam_to_bqsr_covariates, bamread_bqsr_*, trim_bamread,
bam_to_report) needs pysam and is out of scope.
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


def _needs_pysam(name):
    def f(*args, **kwargs):
        raise NotImplementedError("kbbq.gatk.bqsr.%s reads BAM files through pysam; out of scope of the B200 path" % name)
    f.__name__ = name
    return f


for _n in ("bam_to_bqsr_covariates", "bamread_bqsr_cycle", "bamread_bqsr_dinuc", "bamread_adaptor_boundary",
           "trim_bamread", "bam_to_report"):
    globals()[_n] = _needs_pysam(_n)
del _n

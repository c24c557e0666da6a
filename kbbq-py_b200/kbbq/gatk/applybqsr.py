"""Mirror of the hot-path part of the reference's kbbq/gatk/applybqsr.py: get_delta_qs (:80-103).

Out of scope (BAM input / GATK report): table_to_vectors, bamread_cycle_covariates,
bamread_dinuc_covariates, recalibrate_bamread.
"""
import numpy as np

from .. import _native


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

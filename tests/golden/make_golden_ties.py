#!/usr/bin/env python3
"""Near-tie cells of gatk_delta_q run through the UNMODIFIED reference (kbbq/compare_reads.py:235-260).

    python tests/golden/make_golden_ties.py      (build container only: needs /root/reference)

oracle.near_tie_cells constructs cells whose two best candidates are as close as integer counts allow; the
reference's answers on them pin the tie-breaking of the restatement (first maximum of the long-double sums).
Two sets:
  delta_near_ties.npz   ~50 k cells with 1e3 .. 1e10 observations: the reference's delta for each
  and, printed only, the census of a 2 M-cell search at 1e10 .. 1e11 observations: there the reference's
  scipy.stats.binom.logpmf adds the candidate-independent term gammaln(n+1) - gammaln(k+1) - gammaln(n-k+1)
  BEFORE rounding to fp64, so two candidates closer than the rounding noise of a ~1e10-sized number can come out
  in either order; the restatement (oracle and kernel) leaves that constant out.  The cells
  where that changed the answer are stored in delta_near_ties.npz under known_diff_* (documentation, not a test).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_shim  # noqa: E402

rc, cr, ab = ref_shim.load()
import oracle  # noqa: E402

oracle.build()


def reference(pq, errs, tot):
    return np.concatenate([cr.gatk_delta_q(pq[i:i + 4096], errs[i:i + 4096], tot[i:i + 4096])
                           for i in range(0, len(pq), 4096)])


if __name__ == "__main__":
    pq, errs, tot = oracle.near_tie_cells(75_000, 31, 3.0, 10.0)
    want = reference(pq, errs, tot)
    got = oracle.gatk_delta_q(pq, errs, tot)
    gap = oracle.delta_q_top2_gap(pq, errs, tot)
    print("set 1: %d cells, oracle != reference on %d; relative top-2 gap < 2^-40: %d, < 2^-50: %d" % (
        len(pq), int((want != got).sum()), int((gap < 2.0 ** -40).sum()), int((gap < 2.0 ** -50).sum())))
    assert (want == got).all()
    n2 = int(os.environ.get("KBBQ_TIES_SEARCH", "2000000"))
    p2, e2, t2 = oracle.near_tie_cells(int(n2 * 1.5), 32, 10.0, 11.0)
    w2 = reference(p2, e2, t2)
    g2 = oracle.gatk_delta_q(p2, e2, t2)
    gap2 = oracle.delta_q_top2_gap(p2, e2, t2)
    bad = np.nonzero(w2 != g2)[0]
    print("set 2: %d cells at 1e10 .. 1e11 observations: relative top-2 gap < 2^-60: %d (exact fp80 ties: %d); "
          "reference != restatement on %d" % (len(p2), int((gap2 < 2.0 ** -60).sum()), int((gap2 == 0).sum()), len(bad)))
    np.savez_compressed(os.path.join(HERE, "delta_near_ties.npz"), prior=pq.astype(np.int8), errs=errs, total=tot,
                        dq=want.astype(np.int8), known_diff_prior=p2[bad], known_diff_errs=e2[bad], known_diff_total=t2[bad],
                        known_diff_reference=w2[bad], known_diff_restatement=g2[bad], known_diff_searched=len(p2),
                        known_diff_gap=gap2[bad])

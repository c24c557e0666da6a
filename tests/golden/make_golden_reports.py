#!/usr/bin/env python3
"""Golden GATK recalibration reports, written by the UNMODIFIED Python reference.

Run in the build container only (needs /root/reference; imported through oracle/ref_shim, which
re-adds the pandas / NumPy aliases the reference's pinned stack had):

    python tests/golden/make_golden_reports.py

For the count tables of a few golden cases (tests/golden/<case>.npz, themselves produced by the
reference) it calls kbbq.gatk.bqsr.vectors_to_report (kbbq/gatk/bqsr.py:227-366) and stores
str(report) -- what RecalibrationReport.write() puts on disk (kbbq/recaltable.py:86-99,481-491) --
as tests/golden/report_<case>.txt.  `sparse_r3` is a hand-made table set with an unobserved read
group, unobserved qualities and counts up to 1e9.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_shim  # noqa: E402

ref_shim.load()
from kbbq.gatk import bqsr  # noqa: E402  (the reference's)

KEYS = ("meanq", "rg_errs", "rg_total", "q_errs", "q_total", "pos_errs", "pos_total", "dinuc_errs", "dinuc_total")
CASES = {"tiny_r2": ["lane1.AAGG", "lane2.CCTT"], "tails_r2_second": ["b_second", "a_first"]}


def sparse_case():
    rng = np.random.default_rng(77)
    R, L = 3, 5
    pos_total = np.zeros((R, 43, 2 * L), np.int64)
    pos_errs = np.zeros_like(pos_total)
    din_total = np.zeros((R, 43, 16), np.int64)
    din_errs = np.zeros_like(din_total)
    for rg in (0, 2):  # read group 1 is never observed
        for q in (6, 7, 20, 33, 41, 42):
            t = rng.integers(0, 10 ** rng.integers(1, 10), size=2 * L)
            t[rng.random(2 * L) < 0.3] = 0
            pos_total[rg, q] = t
            pos_errs[rg, q] = (t * rng.random(2 * L) * 10 ** (-q / 10.0) * 2).astype(np.int64)
            d = rng.integers(0, 10 ** rng.integers(1, 9), size=16)
            d[rng.random(16) < 0.3] = 0
            din_total[rg, q] = d
            din_errs[rg, q] = (d * rng.random(16) * 10 ** (-q / 10.0) * 2).astype(np.int64)
    q_total, q_errs = pos_total.sum(axis=2), pos_errs.sum(axis=2)
    rg_total, rg_errs = q_total.sum(axis=1), q_errs.sum(axis=1)
    meanq = np.array([30, 0, 25], np.int64)
    return dict(zip(KEYS, (meanq, rg_errs, rg_total, q_errs, q_total, pos_errs, pos_total, din_errs, din_total)))


def main():
    for case, rgs in CASES.items():
        d = np.load(os.path.join(HERE, case + ".npz"), allow_pickle=True)
        rep = bqsr.vectors_to_report(*[d[k] for k in KEYS], rgs)
        with open(os.path.join(HERE, "report_%s.txt" % case), "w") as fh:
            fh.write(str(rep))
    d = sparse_case()
    np.savez_compressed(os.path.join(HERE, "report_sparse_r3.npz"), **d)
    rep = bqsr.vectors_to_report(*[d[k] for k in KEYS], ["rgA", "rgB", "rgC"])
    with open(os.path.join(HERE, "report_sparse_r3.txt"), "w") as fh:
        fh.write(str(rep))


if __name__ == "__main__":
    main()

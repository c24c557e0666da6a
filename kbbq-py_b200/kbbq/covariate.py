"""Mirror of the reference's kbbq/covariate.py: object-shaped covariate count tables.

Same class and method names as the reference (kbbq/covariate.py:23-465): every covariate holds an
`errors` and a `total` array indexed by the covariate values; tables grow on demand.  The per-read
methods (`consume_read`) are host-side bookkeeping with the semantics the reference documents --
including the two index-tuple slips at kbbq/covariate.py:276-280 and :425, which are
NOT reproduced: errors are counted at (rg-of-errors, q-of-errors, ...) and observations at
(rg-of-valid, q-of-valid, ...), each tuple of equal length, which is what the unit tests of the
reference assert on their toy reads.

Bulk tallying of packed batches goes to the GPU: :meth:`CovariateData.consume_packed` calls
kbbq_build (csrc/build.cuh) and adds the resulting int64 tables, with the raw path's rule that a
base below minscore is never tallied (kbbq/recalibrate.py:96-101).
"""
import numpy as np

from . import _native
from . import compare_reads


def pad_axis(array, axis, n):
    """`array` with n zeros appended along `axis` (reference: kbbq/covariate.py:23-32).

    The reference's np.append with float zeros silently turns integer tables into float64; here the
    dtype of `array` is kept.
    """
    pad = np.zeros(array.shape[0:axis] + (n,) + array.shape[axis + 1:], dtype=array.dtype)
    return np.append(array, pad, axis=axis)


class Covariate:
    """Two equally shaped int64 arrays: `errors` and `total` (reference: kbbq/covariate.py:34-191)."""

    def __init__(self, shape=0):
        self.errors = np.zeros(shape, dtype=np.int64)
        self.total = np.zeros(shape, dtype=np.int64)

    def pad_axis(self, axis, n=1):
        self.errors = pad_axis(self.errors, axis=axis, n=n)
        self.total = pad_axis(self.total, axis=axis, n=n)

    def pad_axis_to_fit(self, axis, idx):
        """Grow `axis` until the scalar index `idx` (negative allowed) is valid."""
        axislen = self.shape()[axis]
        if idx < -axislen or idx >= axislen:
            self.pad_axis(axis=axis, n=(-idx - axislen) if idx < 0 else (idx - axislen + 1))

    def increment(self, idx, value=(1, 1)):
        """np.add.at on both arrays: idx[0] indexes `errors`, idx[1] indexes `total`."""
        np.add.at(self.errors, idx[0], value[0])
        np.add.at(self.total, idx[1], value[1])

    def shape(self):
        assert self.total.shape == self.errors.shape
        return self.total.shape

    def __getitem__(self, key):
        return (self.errors[key], self.total[key])

    def __setitem__(self, key, value):
        self.errors[key] = value[0]
        self.total[key] = value[1]


class RGCovariate(Covariate):
    """1-D, read group (reference: kbbq/covariate.py:193-234)."""

    def __init__(self):
        super().__init__(shape=0)

    def consume_read(self, read):
        rge, rgv = read.get_rg_errors()
        self.pad_axis_to_fit(axis=0, idx=read.get_rg_int())
        self.increment((rge, rgv))
        return rge, rgv

    def num_rgs(self):
        return self.shape()[0]


class QCovariate(Covariate):
    """2-D, read group x quality; owns the RGCovariate (reference: kbbq/covariate.py:236-290)."""

    def __init__(self):
        self.rgcov = RGCovariate()
        super().__init__(shape=(0, 0))

    def consume_read(self, read):
        rge, rgv = self.rgcov.consume_read(read)
        self.pad_axis_to_fit(axis=0, idx=self.rgcov.num_rgs() - 1)
        qe, qv = read.get_q_errors()
        if qv.size:
            self.pad_axis_to_fit(axis=1, idx=int(np.amax(qv)))
        self.increment(idx=((rge, qe), (rgv, qv)))
        return (rge, rgv), (qe, qv)

    def num_qs(self):
        return self.shape()[1]


class CycleCovariate(Covariate):
    """3-D, read group x quality x cycle; the cycle axis has length 2L: read-1 cycles from the
    front, read-2 cycles (negative indices) from the back, so growing it must keep both halves in
    place (reference: kbbq/covariate.py:292-354)."""

    def __init__(self):
        super().__init__(shape=(0, 0, 0))

    def pad_axis(self, axis, n=1):
        if not (axis == 2 or axis == -1):
            super().pad_axis(axis=axis, n=n)
            return
        if n % 2 != 0:
            raise ValueError('n should be even for the 2nd axis of a CycleCovariate. n = {} was given.'.format(n))
        oldlen = self.shape()[2]
        if oldlen == 0:
            super().pad_axis(axis=axis, n=n)
            return
        half = oldlen // 2
        grown = []
        for old in (self.errors, self.total):
            new = np.zeros(old.shape[0:2] + (oldlen + n,), dtype=old.dtype)
            new[..., 0:half] = old[..., 0:half]
            new[..., -half:] = old[..., -half:]
            grown.append(new)
        self.errors, self.total = grown

    def num_cycles(self):
        return self.shape()[-1] / 2


class DinucCovariate(Covariate):
    """3-D, read group x quality x 16 dinucleotides (reference: kbbq/covariate.py:356-373)."""

    def __init__(self):
        super().__init__(shape=(0, 0, len(compare_reads.Dinucleotide.dinucs)))

    def num_dinucs(self):
        return self.shape()[-1]


class CovariateData:
    """The three tables of one data set (reference: kbbq/covariate.py:375-465)."""

    def __init__(self):
        self.qcov = QCovariate()
        self.cyclecov = CycleCovariate()
        self.dinuccov = DinucCovariate()

    def _fit(self, num_rgs, num_qs, readlen):
        for cov in (self.cyclecov, self.dinuccov):
            if num_rgs:
                cov.pad_axis_to_fit(axis=0, idx=num_rgs - 1)
            if num_qs:
                cov.pad_axis_to_fit(axis=1, idx=num_qs - 1)
        if readlen:
            have = self.cyclecov.shape()[2]
            if have < 2 * readlen:
                self.cyclecov.pad_axis(axis=2, n=2 * readlen - have)

    def consume_read(self, read):
        """Tally one :class:`kbbq.read.ReadData` on the host."""
        (rge, rgv), (qe, qv) = self.qcov.consume_read(read)
        ce, cv = read.get_cycle_errors()
        dinuc = read.get_dinucleotide_array()
        skips, errors = np.asarray(read.skips, dtype=bool), np.asarray(read.errors, dtype=bool)
        dvalid = np.logical_and(dinuc != -1, ~skips)
        derr = np.logical_and(dvalid, errors)
        self._fit(self.get_num_rgs(), self.get_num_qs(), len(read))
        self.cyclecov.increment(idx=((rge, qe, ce), (rgv, qv, cv)))
        rg = read.get_rg_int()
        qual = np.asarray(read.qual)
        self.dinuccov.increment(idx=((np.full(int(derr.sum()), rg), qual[derr], dinuc[derr]),
                                     (np.full(int(dvalid.sum()), rg), qual[dvalid], dinuc[dvalid])))

    def consume_packed(self, seq, qual, corr, rg, second, num_rgs=None, minscore=6, device=None):
        """Tally a packed batch (see :func:`kbbq.read.pack_reads`) on the GPU and add it in.

        seq/qual/corr u8[N, L], rg u16[N] (or None), second u8[N] (or None).  Raw-path rule: bases
        below `minscore` are not tallied.  rg / q tables are the exact marginals of the cycle table.
        """
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        if seq.ndim != 2 or seq.shape[0] == 0:
            return
        L = seq.shape[1]
        R = int(num_rgs if num_rgs is not None else (int(np.max(rg)) + 1 if rg is not None and len(rg) else 1))
        pe, pt, de, dt = _native.build_host(seq, qual, corr, rg if R > 1 or rg is not None else None, second, L, R,
                                            minscore=minscore, device=device)
        self.qcov.rgcov.pad_axis_to_fit(axis=0, idx=R - 1)
        self.qcov.pad_axis_to_fit(axis=0, idx=R - 1)
        self.qcov.pad_axis_to_fit(axis=1, idx=_native.NQ - 1)
        self._fit(max(R, self.get_num_rgs()), self.get_num_qs(), L)
        twoL = self.cyclecov.shape()[2]
        if twoL != 2 * L:
            raise ValueError("packed batches must have the read length of the table (%d)" % (twoL // 2))
        self.cyclecov.errors[:R, :_native.NQ] += pe
        self.cyclecov.total[:R, :_native.NQ] += pt
        self.dinuccov.errors[:R, :_native.NQ] += de
        self.dinuccov.total[:R, :_native.NQ] += dt
        self.qcov.errors[:R, :_native.NQ] += pe.sum(axis=2)
        self.qcov.total[:R, :_native.NQ] += pt.sum(axis=2)
        self.qcov.rgcov.errors[:R] += pe.sum(axis=(1, 2))
        self.qcov.rgcov.total[:R] += pt.sum(axis=(1, 2))

    def get_num_rgs(self):
        return self.qcov.rgcov.num_rgs()

    def get_num_qs(self):
        return self.qcov.num_qs()

    def get_num_cycles(self):
        return self.cyclecov.num_cycles()

    def get_num_dinucs(self):
        return self.dinuccov.num_dinucs()

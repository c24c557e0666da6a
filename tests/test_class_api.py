"""Class API of the path: kbbq.read.ReadData and kbbq.covariate.*.

The known answers are the ones the reference's own unit tests hold for its toy read
(tests/conftest.py:204-218: seq ATG, qual [6,10,3], skips [F,F,T], errors [F,T,T]):
tests/test_read.py:87-136 and tests/test_covariate.py:6-200.  The GPU test checks that the packed
bulk path (CovariateData.consume_packed -> kbbq_build) gives the tables the per-read host methods
give on reads where both rules coincide (skips == quality below minscore).
"""
import numpy as np
import pytest

from kbbq import covariate, read


@pytest.fixture
def exreaddata():
    yield read.ReadData(seq=np.array(['A', 'T', 'G']), qual=np.array([6, 10, 3]),
                        skips=np.array([False, False, True]), name='read01', rg=0, second=False,
                        errors=np.array([False, True, True]))
    read.ReadData.rg_to_pu = dict()
    read.ReadData.rg_to_int = dict()
    read.ReadData.numrgs = 0


class FakeFastq:
    def __init__(self, name, sequence, quality):
        self.name, self.sequence, self.quality = name, sequence, quality

    def get_quality_array(self, offset=33):
        return [ord(c) - offset for c in self.quality]


# ---- read.ReadData ---------------------------------------------------------------------------

def test_readdata_registry_and_accessors(exreaddata):
    assert read.ReadData.rg_to_pu[0] == 0 and read.ReadData.rg_to_int[0] == 0 and read.ReadData.numrgs == 1
    assert exreaddata.str_qual() == ["'", "+", "$"]
    assert exreaddata.canonical_name() == 'read01/1'
    assert exreaddata.get_rg_int() == 0 and exreaddata.get_pu() == 0
    assert len(exreaddata) == 3
    assert np.array_equal(exreaddata.not_skipped_errors(), [False, True, False])


def test_readdata_extractors(exreaddata):
    e, v = exreaddata.get_rg_errors()
    assert np.array_equal(e, [0]) and np.array_equal(v, [0, 0])
    e, v = exreaddata.get_q_errors()
    assert np.array_equal(e, [10]) and np.array_equal(v, [6, 10])
    assert np.array_equal(exreaddata.get_cycle_array(), [0, 1, 2])
    e, v = exreaddata.get_cycle_errors()
    assert np.array_equal(e, [1]) and np.array_equal(v, [0, 1])
    assert np.array_equal(exreaddata.get_dinucleotide_array(), [-1, 1, -1])
    e, v = exreaddata.get_dinuc_errors()
    assert np.array_equal(e, [1]) and np.array_equal(v, [1])
    exreaddata.second = True
    assert np.array_equal(exreaddata.get_cycle_array(), [-1, -2, -3])


def test_readdata_from_fastq():
    fq = FakeFastq('r001', 'TTAGATAAAGGATACTG', '==99=?<*+/5:@A99:')
    try:
        r = read.ReadData.from_fastq(fq, rg='foo', second=True)
        assert r.name == 'r001' and read.ReadData.rg_to_int['foo'] == 0 and r.second is True
        assert np.array_equal(r.qual, [ord(c) - 33 for c in fq.quality])
        fq.name = 'r001/1'
        r = read.ReadData.from_fastq(fq)
        assert r.rg is None and r.second is False and r.name == 'r001'
        fq.name = 'r001/2_RG:Z:foo'
        r = read.ReadData.from_fastq(fq)
        assert r.name == 'r001' and r.rg == 'foo' and r.second is True
        assert read.ReadData.rg_to_int['foo'] == 0 and read.ReadData.rg_to_int[None] == 1
    finally:
        read.ReadData.rg_to_pu, read.ReadData.rg_to_int, read.ReadData.numrgs = dict(), dict(), 0


def test_readdata_from_bamread_duck_typed():
    class Bam:
        query_sequence, query_name, is_reverse, is_read2 = 'TTAGATAAAGGATACTG', 'r001', False, False
        query_qualities = [28, 28, 24, 24, 28, 30, 27, 9, 10, 14, 20, 25, 31, 32, 24, 24, 25]
        tags = {}

        def has_tag(self, t):
            return t in self.tags

        def get_tag(self, t):
            return self.tags[t]
    try:
        b = Bam()
        r = read.ReadData.from_bamread(b)
        assert r.rg is None and np.array_equal(r.qual, b.query_qualities)
        b.tags = {'OQ': '(' * 17, 'RG': 'foo'}
        r = read.ReadData.from_bamread(b, use_oq=True)
        assert np.array_equal(r.qual, [7] * 17) and r.rg == 'foo'
        b.is_reverse = True
        r = read.ReadData.from_bamread(b)
        assert np.array_equal(r.qual, b.query_qualities[::-1])
        assert np.array_equal(r.seq, list('CAGTATCCTTTATCTAA'))
    finally:
        read.ReadData.rg_to_pu, read.ReadData.rg_to_int, read.ReadData.numrgs = dict(), dict(), 0


# ---- covariate.* -----------------------------------------------------------------------------

def test_pad_axis_keeps_dtype():
    assert np.array_equal(covariate.pad_axis(np.array([1]), 0, 2), [1, 0, 0])
    assert np.array_equal(covariate.pad_axis(np.array([[1]]), 0, 2), [[1], [0], [0]])
    out = covariate.pad_axis(np.array([[1]]), 1, 2)
    assert np.array_equal(out, [[1, 0, 0]]) and out.dtype == np.array([[1]]).dtype


def test_covariate_base_class():
    c = covariate.Covariate()
    assert c.errors.shape == (0,) and c.total.shape == (0,) and c.shape() == (0,)
    c.pad_axis(0)
    assert c.shape() == (1,)
    c = covariate.Covariate((3, 4))
    c.pad_axis(1, 2)
    assert c.shape() == (3, 6)
    c = covariate.Covariate()
    c.pad_axis_to_fit(0, 99)
    assert c.shape() == (100,)
    c = covariate.Covariate((1, 2))
    c.pad_axis_to_fit(1, -10)
    assert c.shape() == (1, 10)
    c.pad_axis_to_fit(1, 0)
    assert c.shape() == (1, 10)
    c = covariate.Covariate((10,))
    c.increment((0, 0), (0, 1))
    assert not c.errors.any() and np.array_equal(c.total, [1] + [0] * 9) and c[0] == (0, 1)
    c = covariate.Covariate((1,))
    c[0] = (0, 1)
    c.increment((0, 0), (0, 1))
    assert c[0] == (0, 2)


def test_rg_and_q_covariates(exreaddata):
    rgc = covariate.RGCovariate()
    assert rgc.shape() == (0,) and rgc.num_rgs() == 0
    e, v = rgc.consume_read(exreaddata)
    assert np.array_equal(e, [0]) and np.array_equal(v, [0, 0]) and rgc[0] == (1, 2) and rgc.num_rgs() == 1
    qc = covariate.QCovariate()
    assert qc.shape() == (0, 0) and qc.num_qs() == 0
    (rge, rgv), (qe, qv) = qc.consume_read(exreaddata)
    assert np.array_equal(rge, [0]) and np.array_equal(rgv, [0, 0])
    assert np.array_equal(qe, [10]) and np.array_equal(qv, [6, 10])
    assert qc[(0, 10)] == (1, 1) and qc[(0, 6)] == (0, 1) and qc.num_qs() == 11


def test_cycle_covariate_growth_keeps_both_halves():
    c = covariate.CycleCovariate()
    assert c.shape() == (0, 0, 0)
    c.pad_axis(axis=0, n=1)
    assert c.shape() == (1, 0, 0)
    with pytest.raises(ValueError):
        c.pad_axis(2, 1)
    c.pad_axis(1, 1)
    c.pad_axis(2, 2)
    assert c.shape() == (1, 1, 2) and c.num_cycles() == 1
    c[(0, 0, 0)] = (1, 1)
    c[(0, 0, -1)] = (2, 2)
    c.pad_axis(2, 4)
    assert c.shape() == (1, 1, 6) and c[0, 0, 0] == (1, 1) and c[0, 0, -1] == (2, 2)
    assert c.total.sum() == 3


def test_dinuc_covariate_and_container(exreaddata):
    d = covariate.DinucCovariate()
    assert d.shape() == (0, 0, 16) and d.num_dinucs() == 16
    data = covariate.CovariateData()
    assert data.qcov.shape() == (0, 0) and data.cyclecov.shape() == (0, 0, 0) and data.dinuccov.shape() == (0, 0, 16)
    assert (data.get_num_rgs(), data.get_num_qs(), data.get_num_cycles(), data.get_num_dinucs()) == (0, 0, 0, 16)
    data.consume_read(exreaddata)
    assert data.qcov.rgcov[0] == (1, 2)
    assert data.qcov[0, 10] == (1, 1)
    assert data.cyclecov[0, 6, 0] == (0, 1)
    assert data.cyclecov[0, 10, 1] == (1, 1)
    assert data.dinuccov[0, 10, 1] == (1, 1)
    assert data.cyclecov.shape() == (1, 11, 6) and data.get_num_cycles() == 3


def test_consume_read_on_real_length_reads_does_not_raise():
    """The reference's CovariateData.consume_read raises IndexError on any realistic read;
    the intended semantics must hold: totals equal the number of non-skipped bases."""
    rng = np.random.default_rng(5)
    try:
        data = covariate.CovariateData()
        n_valid = 0
        for i in range(6):
            L = 40
            seq = np.array(list(rng.choice(list('ACGTN'), size=L, p=[.24, .24, .24, .24, .04])))
            qual = rng.integers(2, 42, size=L)
            r = read.ReadData(seq=seq, qual=qual, skips=qual < 6, name='r%d' % i, rg='g%d' % (i % 2),
                              second=bool(i & 1), errors=rng.random(L) < 0.1)
            data.consume_read(r)
            n_valid += int((qual >= 6).sum())
        assert data.qcov.rgcov.total.sum() == n_valid == data.qcov.total.sum() == data.cyclecov.total.sum()
        assert data.get_num_rgs() == 2 and data.cyclecov.shape()[2] == 80
    finally:
        read.ReadData.rg_to_pu, read.ReadData.rg_to_int, read.ReadData.numrgs = dict(), dict(), 0


@pytest.mark.gpu
def test_consume_packed_matches_per_read_host_tally():
    rng = np.random.default_rng(11)
    try:
        reads = []
        for i in range(257):
            L = 50
            seq = np.array(list(rng.choice(list('ACGTN'), size=L, p=[.245, .245, .245, .245, .02])))
            qual = rng.integers(2, 43, size=L)
            reads.append(read.ReadData(seq=seq, qual=qual, skips=qual < 6, name='r%d' % i, rg='g%d' % (i % 3),
                                       second=bool(i & 1), errors=rng.random(L) < 0.05))
        host = covariate.CovariateData()
        for r in reads:
            host.consume_read(r)
        dev = covariate.CovariateData()
        packed = read.pack_reads(reads)
        dev.consume_packed(*packed, num_rgs=3)
        nq = host.get_num_qs()
        for a, b in ((host.qcov.rgcov, dev.qcov.rgcov), (host.qcov, dev.qcov), (host.cyclecov, dev.cyclecov),
                     (host.dinuccov, dev.dinuccov)):
            for x, y in ((a.errors, b.errors), (a.total, b.total)):
                if x.ndim == 1:
                    assert np.array_equal(x, y)
                else:
                    assert np.array_equal(x, y[:, :nq]) and not y[:, nq:].any()
    finally:
        read.ReadData.rg_to_pu, read.ReadData.rg_to_int, read.ReadData.numrgs = dict(), dict(), 0

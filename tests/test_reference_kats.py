"""The reference's own known-answer tests, pointed at the drop-in package (B200 only).

Carried over from tests/test_recalibrate.py:19-135, tests/test_compare_reads.py:141-151,219-233 and
tests/test_gatk_applybqsr.py:105-121 of adamjorr/kbbq-py (same fixtures, same expected values)."""
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture()
def fastx_mod():
    from kbbq import fastx
    return fastx


@pytest.fixture()
def uncorr_and_corr_fastq_files(tmp_path, fastx_mod):
    r = fastx_mod.FastxRecord(name='foo', sequence='ATG', quality='((#')  # 7, 7, 2
    r2 = fastx_mod.FastxRecord(name=r.name, sequence='ACG', quality=r.quality)
    fu, fc = tmp_path / 'uncorr.fq', tmp_path / 'corr.fq'
    fu.write_text(str(r))
    fc.write_text(str(r2))
    return str(fu), str(fc)


@pytest.fixture()
def uncorr_and_corr_with_rg(tmp_path, fastx_mod):
    r = fastx_mod.FastxRecord(name='foo/1_RG:Z:bar', sequence='ATG', quality='((#')
    r2 = fastx_mod.FastxRecord(name=r.name, sequence='ACG', quality=r.quality)
    fu, fc = tmp_path / 'uncorr_withrg.fq', tmp_path / 'corr_withrg.fq'
    fu.write_text(str(r))
    fc.write_text(str(r2))
    return str(fu), str(fc)


def test_find_corrected_sites(fastx_mod):
    from kbbq import recalibrate
    r = fastx_mod.FastxRecord(name='r001/1', sequence='TTAGATAAAGGATACTG', quality='==99=?<*+/5:@A99:')
    r2 = fastx_mod.FastxRecord(name=r.name, sequence='TTAGACAAAGGATACTG', quality=r.quality)
    correct = np.zeros(17, dtype=bool)
    correct[5] = True
    assert np.array_equal(recalibrate.find_corrected_sites(r, r2), correct)
    with pytest.raises(AssertionError):
        recalibrate.find_corrected_sites(r, fastx_mod.FastxRecord(name='zzz', sequence=r.sequence, quality=r.quality))


def test_fastq_to_covariate_arrays(uncorr_and_corr_fastq_files, uncorr_and_corr_with_rg):
    from kbbq import compare_reads, recalibrate
    correct_pos_errs = np.zeros((1, 43, 6))
    correct_pos_total = np.zeros((1, 43, 6))
    correct_pos_errs[0, 7, 1] = 1
    correct_pos_total[0, 7, 0] = 1
    correct_pos_total[0, 7, 1] = 1
    correct_dinuc_errs = np.zeros((1, 43, 16))
    correct_dinuc_total = np.zeros((1, 43, 16))
    correct_dinuc_errs[0, 7, compare_reads.Dinucleotide.dinuc_to_int['AT']] = 1
    correct_dinuc_total[0, 7, compare_reads.Dinucleotide.dinuc_to_int['AT']] = 1
    correct_vectors = [np.array([6]), np.array([1]), np.array([2]),
                       np.array([[0, 0, 0, 0, 0, 0, 0, 1] + [0] * 35]),
                       np.array([[0, 0, 0, 0, 0, 0, 0, 2] + [0] * 35]),
                       correct_pos_errs, correct_pos_total, correct_dinuc_errs, correct_dinuc_total]
    for a, b in zip(correct_vectors, recalibrate.fastq_to_covariate_arrays(uncorr_and_corr_fastq_files)):
        assert np.array_equal(a, b)
    for a, b in zip(correct_vectors, recalibrate.fastq_to_covariate_arrays(uncorr_and_corr_with_rg, infer_rg=True)):
        assert np.array_equal(a, b)


def test_recalibrate_fastq_driver(uncorr_and_corr_fastq_files, uncorr_and_corr_with_rg, capfd, fastx_mod):
    from kbbq import recalibrate
    correct_read = fastx_mod.FastxRecord(name='foo', sequence='ATG', quality='\'\'#')  # 6, 6, 2
    correct_read_with_rg = fastx_mod.FastxRecord(name='foo/1_RG:Z:bar', sequence='ATG', quality='\'\'#')
    recalibrate.recalibrate_fastq(uncorr_and_corr_fastq_files)
    assert capfd.readouterr().out == str(correct_read) + '\n'
    recalibrate.recalibrate_fastq(uncorr_and_corr_with_rg, infer_rg=True)
    assert capfd.readouterr().out == str(correct_read_with_rg) + '\n'


def test_recalibrate_dispatch_and_main(uncorr_and_corr_fastq_files, capfd, monkeypatch):
    import kbbq.main
    from kbbq import recalibrate
    correct = '@foo\nATG\n+\n\'\'#\n'
    recalibrate.recalibrate(bam=None, fastq=uncorr_and_corr_fastq_files)
    assert capfd.readouterr().out == correct
    with pytest.raises(NotImplementedError):
        recalibrate.recalibrate_bam(None)
    with pytest.raises(NotImplementedError):
        recalibrate.recalibrate(fastq=None, bam='foo')
    with pytest.raises(NotImplementedError):
        recalibrate.recalibrate(fastq=None, bam=None, gatkreport='foo')
    with pytest.raises(ValueError):
        recalibrate.recalibrate(fastq=None, bam=None, gatkreport=None)
    with monkeypatch.context() as m:
        m.setattr(sys, 'argv', [sys.argv[0]] + ["recalibrate", '-f'] + list(uncorr_and_corr_fastq_files))
        kbbq.main.main()
    assert capfd.readouterr().out == correct
    with pytest.raises(NotImplementedError), monkeypatch.context() as m:
        m.setattr(sys, 'argv', [sys.argv[0]] + ["recalibrate", '-b', 'foo'])
        kbbq.main.main()
    with pytest.raises(NotImplementedError), monkeypatch.context() as m:
        m.setattr(sys, 'argv', [sys.argv[0]] + ["recalibrate", '-b', 'foo', '-g', 'bar'])
        kbbq.main.main()


def test_gatk_delta_q():
    from kbbq import compare_reads
    prior_q = np.array([10, 20, 30])
    numerrs = np.array([10, 200, 0])
    numtotal = np.array([1000, 1000, 50000])
    dq = compare_reads.gatk_delta_q(prior_q, numerrs, numtotal)
    assert dq.shape == prior_q.shape
    assert dq[0] > 0 and dq[1] < 0 and dq[2] > 0
    assert np.all(dq + prior_q <= 42) and np.all(dq + prior_q > 0)
    assert dq.tolist() == [3, -8, 2]  # the reference's actual values
    # n-d inputs keep their shape
    assert compare_reads.gatk_delta_q(prior_q.reshape(3, 1), numerrs.reshape(3, 1), numtotal.reshape(3, 1)).shape == (3, 1)


def test_recalibrate_fastq_single_read(fastx_mod):
    from kbbq import compare_reads
    read = fastx_mod.FastxRecord(name='foo', sequence='ATG', quality='((#')  # 7, 7, 2
    meanq = np.array([10])
    globaldeltaq = np.array([1])
    qscoredeltaq = np.array([[2, 2, 2, 2, 2, 2, 2, 2]])
    positiondeltaq = np.zeros((1, 8, 6))
    positiondeltaq[0, 7, :] = 3
    dinucdeltaq = np.zeros([1, 8, 16])
    dinucdeltaq[0, 7, :] = 5
    assert np.array_equal(compare_reads.recalibrate_fastq(read, meanq, globaldeltaq, qscoredeltaq, positiondeltaq,
                                                          dinucdeltaq, np.array([0]),
                                                          compare_reads.Dinucleotide.dinuc_to_int),
                          np.array([21, 21, 2]))


def test_get_delta_qs():
    from kbbq.gatk import applybqsr
    rgdeltaq, qscoredeltaq, positiondeltaq, dinucdeltaq = applybqsr.get_delta_qs(
        np.array([10]), np.array([0]), np.array([1000]), np.array([[0]]), np.array([[1000]]),
        np.array([[[0]]]), np.array([[[1000]]]), np.array([[[0]]]), np.array([[[1000]]]))
    assert np.array_equal(rgdeltaq, np.array([3]))
    assert np.array_equal(qscoredeltaq, np.array([[2]]))
    assert np.array_equal(positiondeltaq, np.array([[[1]]]))
    assert np.array_equal(dinucdeltaq, np.array([[[1, 0]]]))

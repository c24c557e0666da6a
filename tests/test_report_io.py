"""GATK recalibration report I/O.

CPU part: the reference's own known-answer tests for kbbq/recaltable.py and
applybqsr.table_to_vectors (tests/test_recaltable.py:12-164, tests/test_gatk_applybqsr.py:11-63,
restated inline; the ones that need the undownloadable tests/data report use reports the reference
wrote here, tests/golden/report_*.txt made by tests/golden/make_golden_reports.py) -- parsing,
printing (byte-identical) and the report -> vectors direction are host text / index work.
GPU part: vectors -> report, whose EmpiricalQuality columns come from the delta-Q kernels.
"""
import os

import numpy as np
import pandas as pd
import pytest

from conftest import GOLDEN, TABLE_KEYS, load_case

from kbbq import recaltable
from kbbq.gatk import applybqsr

REPORTS = {
    "tiny_r2": ("tiny_r2", ["lane1.AAGG", "lane2.CCTT"]),
    "tails_r2_second": ("tails_r2_second", ["b_second", "a_first"]),
    "sparse_r3": ("report_sparse_r3", ["rgA", "rgB", "rgC"]),
}

EXTABLE = '''#:GATKTable:6:2:%s:%s:%.4f:%.4f:%d:%.2f:;
#:GATKTable:RecalTable0:
ReadGroup                   EventType  EmpiricalQuality  EstimatedQReported  Observations  Errors 
HJCMTCCXX160113.5.AAGGATGT  M                   22.0000             24.3199        210398  1382.00
HK2WYCCXX160124.1.AAGGATGT  M                   22.0000             24.3994        196298  1391.00'''

SMALL_REPORT = """#:GATKReport.v1.1:5
#:GATKTable:2:17:%s:%s:;
#:GATKTable:Arguments:Recalibration argument collection values used in this run
Argument                    Value                                                                   

#:GATKTable:3:94:%d:%d:%d:;
#:GATKTable:Quantized:Quality quantization map
QualityScore  Count    QuantizedScore

#:GATKTable:6:1:%s:%s:%.4f:%.4f:%d:%.2f:;
#:GATKTable:RecalTable0:
ReadGroup  EventType  EmpiricalQuality  EstimatedQReported  Observations  Errors 
1          M                   23.0000              7.0000        200000  1000.00

#:GATKTable:6:1:%s:%d:%s:%.4f:%d:%.2f:;
#:GATKTable:RecalTable1:
ReadGroup  QualityScore  EventType  EmpiricalQuality  Observations  Errors 
1                     7  M                   23.0000        200000  1000.00

#:GATKTable:8:50763:%s:%d:%s:%s:%s:%.4f:%d:%.2f:;
#:GATKTable:RecalTable2:
ReadGroup  QualityScore  CovariateValue  CovariateName  EventType  EmpiricalQuality  Observations  Errors 
1                     7  1               Cycle          M                   23.0000        200000  1000.00
1                     7  AC              Context        M                   23.0000        200000  1000.00

"""


def report_path(case):
    return os.path.join(GOLDEN, "report_%s.txt" % case)


# ---- GATKReport (reference tests/test_recaltable.py:12-63) ---------------------------------------

def test_report_init_header_repr_eq():
    assert recaltable.GATKReport([]).tables == []
    r = recaltable.GATKReport([])
    assert r.get_headerstring() == '#:GATKReport.v1.1:0'
    assert repr(r) == '#:GATKReport.v1.1:0\n\n'
    r.tables = [0, 1, 2]
    assert r.get_headerstring() == '#:GATKReport.v1.1:3'
    assert repr(r) == '#:GATKReport.v1.1:3\n0\n1\n2\n'
    report = recaltable.GATKReport([0, 1, 2])
    assert report == recaltable.GATKReport([0, 1, 2])
    assert not report == recaltable.GATKReport([4, 5, 6])
    assert not report == recaltable.GATKReport([])
    assert not report == recaltable.GATKReport([], version='0.0')
    assert not report == 3


@pytest.mark.parametrize("case", sorted(REPORTS))
def test_report_fromfile_write_roundtrip(case, tmp_path):
    path = report_path(case)
    r = recaltable.GATKReport.fromfile(path)
    assert len(r.tables) == 5
    out = tmp_path / "report.txt"
    recaltable.RecalibrationReport.fromfile(path).write(out)
    assert out.read_bytes() == open(path, "rb").read()  # byte-identical to what the reference wrote
    # truncated file: the header promises more tables than there are
    lines = open(path).read().splitlines(keepends=True)
    cut = tmp_path / "truncated.txt"
    cut.write_text(''.join(lines[:130]))
    with pytest.raises(ValueError):
        recaltable.GATKReport.fromfile(cut)


# ---- GATKTable (reference tests/test_recaltable.py:69-164) ---------------------------------------

def test_table_fromstring_and_formatting(tmp_path):
    assert recaltable.GATKTable(title='foo', description='bar', data='baz').data == 'baz'
    table = recaltable.GATKTable.fromstring(EXTABLE)
    assert table.title == 'RecalTable0' and table.description == ''
    assert table.data.shape == (2, 6)
    assert table.get_fmtstring() == '#:GATKTable:6:2:%s:%s:%.4f:%.4f:%d:%.2f:;'
    assert table.get_colfmts() == ['%s', '%s', '%.4f', '%.4f', '%d', '%.2f']
    assert table.get_datastring() == '\n'.join(EXTABLE.splitlines()[2:])
    assert table.get_nrows() == 2 and table.get_ncols() == 6
    assert str(table) == EXTABLE
    path = tmp_path / 'table.txt'
    with path.open('w') as f:
        table.write(f)
    assert path.read_text() == EXTABLE + '\n'
    lines = EXTABLE.splitlines(keepends=True)[0:3]
    emptystr = ''.join(lines[0:2] + ['  '.join(lines[2].split())])
    empty = recaltable.GATKTable.fromstring(emptystr)
    assert empty.get_datastring() == '\n'.join(emptystr.splitlines()[2:])
    assert recaltable.GATKTable(title='foo', description='bar', data='').get_titlestring() == '#:GATKTable:foo:bar'


def test_table_parse_fmtstring():
    typedict = recaltable.GATKTable.parse_fmtstring(header=['foo', 'bar', 'baz', 'test'],
                                                    fmtstring='#:GATKTable:0:4:%d:%.4f:%s:%x:;')
    assert typedict == {'foo': np.int64, 'bar': np.float64, 'baz': str}


def test_table_repr_and_eq():
    full = recaltable.GATKTable(title='foo', description='bar', data=pd.DataFrame({'spam': ['eggs']}))
    assert repr(full) == "#:GATKTable:1:1:%s:;\n#:GATKTable:foo:bar\n   spam\n0  eggs"
    extable = recaltable.GATKTable.fromstring(EXTABLE)
    notitle = recaltable.GATKTable(title='', description='bar', data=pd.DataFrame({'spam': ['eggs']}))
    nodesc = recaltable.GATKTable(title='foo', description='', data=pd.DataFrame({'spam': ['eggs']}))
    nodata = recaltable.GATKTable(title='foo', description='bar', data=pd.DataFrame())
    for t in (extable, full, notitle, nodesc, nodata):
        assert t == t
    assert extable != full
    for t in (notitle, nodesc, nodata):
        assert full != t and t != full
    assert extable != 5


# ---- RecalibrationReport (reference tests/test_recaltable.py:170-195) ----------------------------

@pytest.mark.parametrize("case", sorted(REPORTS))
def test_recalibration_report_init_and_str(case):
    with pytest.raises(ValueError):
        recaltable.RecalibrationReport([])
    r = recaltable.RecalibrationReport.fromfile(report_path(case))
    for table, title in zip(r.tables, ['Arguments', 'Quantized', 'RecalTable0', 'RecalTable1', 'RecalTable2']):
        assert table.title == title
        assert table.get_nrows() > 1 and table.get_ncols() > 1
    assert list(r.tables[0].data.index.names) == ['Argument']
    assert list(r.tables[1].data.index.names) == ['QualityScore']
    assert list(r.tables[2].data.index.names) == ['ReadGroup']
    assert list(r.tables[3].data.index.names) == ['ReadGroup', 'QualityScore']
    assert list(r.tables[4].data.index.names) == ['ReadGroup', 'QualityScore', 'CovariateName', 'CovariateValue']
    assert str(r) == open(report_path(case)).read()
    assert str(r) == open(report_path(case)).read()  # printing leaves the index order as it was


# ---- report -> vectors (reference tests/test_gatk_applybqsr.py:11-63) ----------------------------

def test_table_to_vectors_small_report(tmp_path):
    p = tmp_path / 'small_report.txt'
    p.write_text(SMALL_REPORT)
    small = recaltable.RecalibrationReport.fromfile(p)
    (meanq, global_errs, global_total, q_errs, q_total, pos_errs, pos_total,
     dinuc_errs, dinuc_total) = applybqsr.table_to_vectors(small, ["1"])
    assert np.array_equal(meanq, np.array([7], dtype=np.float64))
    assert np.array_equal(global_errs, np.array([1000], dtype=np.int64))
    assert np.array_equal(global_total, np.array([200000], dtype=np.int64))
    assert np.array_equal(q_errs, np.array([[0] * 7 + [1000] + [0] * 35], dtype=np.int64))
    assert np.array_equal(q_total, np.array([[0] * 7 + [200000] + [0] * 35], dtype=np.int64))
    want = np.zeros((1, 43, 2), dtype=np.int64)
    want[0, 7, 0] = 1000
    assert np.array_equal(pos_errs, want)
    want[0, 7, 0] = 200000
    assert np.array_equal(pos_total, want)
    want = np.zeros((1, 43, 16), dtype=np.int64)
    want[0, 7, 3] = 1000
    assert np.array_equal(dinuc_errs, want)
    want[0, 7, 3] = 200000
    assert np.array_equal(dinuc_total, want)
    for a in (global_errs, global_total, q_errs, q_total, pos_errs, pos_total, dinuc_errs, dinuc_total):
        assert a.dtype == np.int64
    with pytest.raises(ValueError):
        applybqsr.table_to_vectors(small, ["1", "not there"])


@pytest.mark.parametrize("case", sorted(REPORTS))
def test_table_to_vectors_inverts_the_reference_writer(case):
    """vectors -> (reference) report -> vectors gives the count tables back."""
    npz, rgs = REPORTS[case]
    want = load_case(npz)
    report = recaltable.RecalibrationReport.fromfile(report_path(case))
    observed = [i for i in range(len(rgs)) if want["rg_total"][i] != 0]
    got = applybqsr.table_to_vectors(report, [rgs[i] for i in observed])
    for key, a in zip(TABLE_KEYS[1:], got[1:]):
        assert a.dtype == np.int64
        assert np.array_equal(a, want[key][observed]), key
    # EstimatedQReported is printed with four decimals
    est = report.tables[2].data['EstimatedQReported'].to_numpy()
    assert np.array_equal(got[0], est)
    # a different read-group order permutes the first axis
    if len(observed) > 1:
        rev = applybqsr.table_to_vectors(report, [rgs[i] for i in observed[::-1]])
        assert np.array_equal(rev[6], want["pos_total"][observed[::-1]])


# ---- vectors -> report: EmpiricalQuality from the delta-Q kernels (B200) -------------------------

@pytest.mark.gpu
@pytest.mark.parametrize("case", sorted(REPORTS))
def test_vectors_to_report_matches_the_reference_text(case):
    from kbbq.gatk import bqsr
    npz, rgs = REPORTS[case]
    d = load_case(npz)
    report = bqsr.vectors_to_report(*[d[k] for k in TABLE_KEYS], rgs)
    assert str(report) == open(report_path(case)).read()
    again = applybqsr.table_to_vectors(report, [rg for i, rg in enumerate(rgs) if d["rg_total"][i] != 0])
    keep = d["rg_total"] != 0
    assert np.array_equal(again[6], d["pos_total"][keep])
    assert np.array_equal(again[7], d["dinuc_errs"][keep])


@pytest.mark.gpu
def test_report_checkpoint_resume_gives_the_same_quals():
    """tables -> report file -> tables -> deltas -> apply equals the direct path (checkpoint / resume)."""
    from kbbq import _native
    from kbbq.gatk import bqsr
    d = load_case("mixed_r3")
    L, R = int(d["L"]), int(d["R"])
    rgs = ["g%d" % i for i in range(R)]
    report = bqsr.vectors_to_report(*[d[k] for k in TABLE_KEYS], rgs)
    back = applybqsr.table_to_vectors(recaltable.RecalibrationReport.fromfile(_write(report)), rgs)
    deltas = applybqsr.get_delta_qs(d["meanq"], *back[1:])
    out = _native.apply_host(d["seq"], d["qual"], d["rg"], d["second"], L, R, d["meanq"], *deltas)
    assert np.array_equal(out.astype(np.int64), d["outq"].reshape(out.shape))


def _write(report):
    import tempfile
    fd, path = tempfile.mkstemp(suffix=".recal.txt")
    os.close(fd)
    report.write(path)
    return path

"""GATK recalibration report I/O: mirror of the reference's kbbq/recaltable.py.

Same classes, attributes and method names -- `GATKReport`, `GATKTable`, `RecalibrationReport`, with
`.tables` and a pandas DataFrame in `.data` -- and byte-identical text (pinned by reports the
reference itself wrote, tests/golden/report_*.txt).  The report is the on-disk form of the count
tables the build kernel produces, so it doubles as checkpoint / resume of a recalibration
(`kbbq.gatk.bqsr.vectors_to_report` / `kbbq.gatk.applybqsr.table_to_vectors`).

Not a transcription: the reference formats row by row through pandas `apply` / `itertuples`
(kbbq/recaltable.py:305-336, minutes for the 1.4 M rows of 32 read groups x 250 bp); here every
column is formatted and padded as one NumPy string array and rows are joined column-wise.

File format (kbbq/recaltable.py:191-207): `#:GATKReport.v1.1:<ntables>`, then per table
`#:GATKTable:<ncol>:<nrow>:<fmt>:...:;`, `#:GATKTable:<title>:<description>`, a header line and the
rows, columns separated by two blanks and padded to the widest cell (strings and headers left
justified, numbers right justified), tables separated by an empty line.
"""
import numpy as np
import pandas as pd

_PRECISION = {'EmpiricalQuality': '.4', 'EstimatedQReported': '.4', 'Errors': '.2'}


def _kind_char(dtype):
    """printf conversion of a column: d for integers, f for floats, s for everything else."""
    kind = getattr(dtype, "kind", "O")
    if kind in "iu":
        return 'd'
    if kind == 'f':
        return 'f'
    return 's'


class GATKReport:
    """A GATK report: a version and a list of :class:`GATKTable` (kbbq/recaltable.py:9-124)."""

    def __init__(self, tables, version='1.1'):
        self.tables = tables
        self.version = version

    @classmethod
    def fromfile(cls, filename):
        with open(filename) as fh:
            fullheader = fh.readline()
            _, version, ntables = fullheader.strip().split(':')
            version = version.split(sep='v', maxsplit=1)[-1]
            chunks = [c for c in fh.read().split('\n\n') if c != '']
        tables = [GATKTable.fromstring(c) for c in chunks]
        if len(tables) != int(ntables):
            raise ValueError("Malformed or truncated file %s. The header (%s) implies there should be %s tables "
                             "in this report, but we only found %d." % (filename, fullheader.strip(), ntables, len(tables)))
        return cls(tables, version)

    def get_headerstring(self):
        return '#:GATKReport.v' + self.version + ':' + str(len(self.tables))

    def write(self, filename):
        with open(filename, 'w') as fh:
            fh.write(str(self))

    def __str__(self):
        return self.get_headerstring() + '\n' + '\n\n'.join([str(t) for t in self.tables] + [''])

    def __repr__(self):
        return self.get_headerstring() + '\n' + '\n'.join([repr(t) for t in self.tables]) + '\n'

    def __eq__(self, other):
        if type(other) is not type(self):
            return NotImplemented
        if self.version != other.version or len(self.tables) != len(other.tables):
            return False
        return all(s == o for s, o in zip(self.tables, other.tables))


class GATKTable:
    """One table of a report; `.data` is a pandas DataFrame (kbbq/recaltable.py:126-400)."""

    def __init__(self, title, description, data):
        self.title = title
        self.description = description
        self.data = data
        self.typemap = {np.dtype(np.int64): 'd', np.dtype(np.float64): 'f', str: 's', np.dtype(object): 's'}
        self.precisionmap = dict(_PRECISION)

    @classmethod
    def fromstring(cls, tablestring):
        rows = tablestring.splitlines()
        title, description = rows[1].split(':')[2:4]
        header = rows[2].split()
        typedict = cls.parse_fmtstring(header, rows[0])
        body = [r.split() for r in rows[3:]]
        ncol = len(header)
        for r in body:
            if len(r) != ncol:
                raise ValueError("table %s: a row has %d fields, the header %d" % (title, len(r), ncol))
        cols = {}
        for j, h in enumerate(header):
            cell = [r[j] for r in body]
            t = typedict.get(h)
            if t is np.int64:
                cols[h] = np.array(cell, dtype=np.int64) if cell else np.zeros(0, np.int64)
            elif t is np.float64:
                cols[h] = np.array(cell, dtype=np.float64) if cell else np.zeros(0, np.float64)
            else:
                cols[h] = np.array(cell, dtype=object)
        return cls(title, description, pd.DataFrame(cols, columns=header))

    @staticmethod
    def parse_fmtstring(header, fmtstring):
        """{column title: type} from `#:GATKTable:ncol:nrow:%s:%d:%.4f:;` (kbbq/recaltable.py:221-246)."""
        fmts = fmtstring.split(':')[4:-1]
        out = {}
        for i, h in enumerate(header):
            f = fmts[i]
            if f.endswith('d'):
                out[h] = np.int64
            elif f.endswith('f'):
                out[h] = np.float64
            elif f.endswith('s'):
                out[h] = str
        return out

    def get_unindexed(self):
        """The frame with a named index turned back into columns (kbbq/recaltable.py:267-279)."""
        if list(self.data.index.names) != [None]:
            return self.data.reset_index()
        return self.data.copy()

    def get_colfmts(self):
        frame = self.get_unindexed()
        return ['%' + self.precisionmap.get(h, '') + _kind_char(t) for t, h in zip(frame.dtypes, frame.columns)]

    def get_fmtstring(self):
        return ':'.join(['#', 'GATKTable', str(self.get_ncols()), str(self.get_nrows())] + self.get_colfmts() + [';'])

    def get_titlestring(self):
        return ':'.join(['#', 'GATKTable', self.title, self.description])

    def get_datastring(self):
        """Header line and rows (kbbq/recaltable.py:305-336), formatted column by column."""
        frame = self.get_unindexed()
        fmts = self.get_colfmts()
        header = [str(h) for h in frame.columns]
        n = frame.shape[0]
        cells, widths = [], []
        for h, f in zip(header, fmts):
            col = frame[h].to_numpy()
            if f == '%s':
                txt = np.asarray(col, dtype=str) if n else np.zeros(0, dtype='U1')
            elif f.endswith('d'):
                txt = np.char.mod('%d', col.astype(np.int64)) if n else np.zeros(0, dtype='U1')
            else:
                txt = np.char.mod(f, col.astype(np.float64)) if n else np.zeros(0, dtype='U1')
            # Width quirk of the reference, kept because the text must match: the widths of numeric
            # columns are measured on the values printed with the LAST column's format (the formatting
            # closures at kbbq/recaltable.py:319 all bind the last loop variable), the cells themselves
            # use their own format (:331-335).  Visible once Observations has ten digits.
            measured = txt
            if f != '%s' and n and fmts[-1] != f:
                last = fmts[-1]
                if last == '%s':
                    measured = np.asarray(col, dtype=str)
                elif last.endswith('d'):
                    measured = np.char.mod('%d', col.astype(np.int64))
                else:
                    measured = np.char.mod(last, col.astype(np.float64))
            w = max(len(h), int(np.char.str_len(measured).max()) if n else 0)
            if n:
                cells.append(np.char.ljust(txt, w) if f == '%s' else np.char.rjust(txt, w))
            widths.append(w)
        lines = ['  '.join(h.ljust(w) for h, w in zip(header, widths))]
        if n:
            acc = cells[0]
            for c in cells[1:]:
                acc = np.char.add(np.char.add(acc, '  '), c)
            lines.extend(acc.tolist())
        return '\n'.join(lines)

    def get_nrows(self):
        return self.get_unindexed().shape[0]

    def get_ncols(self):
        return self.get_unindexed().shape[1]

    def write(self, filehandle):
        return filehandle.write(str(self) + '\n')

    def __str__(self):
        return self.get_fmtstring() + '\n' + self.get_titlestring() + '\n' + self.get_datastring()

    def __repr__(self):
        return self.get_fmtstring() + '\n' + self.get_titlestring() + '\n' + repr(self.data)

    def __eq__(self, other):
        if type(other) is not type(self):
            return NotImplemented
        return self.title == other.title and self.description == other.description and self.data.equals(other.data)


class RecalibrationReport(GATKReport):
    """The five tables of a base-quality recalibration report (kbbq/recaltable.py:402-491):

    0 Arguments (Argument -> Value), 1 Quantized (QualityScore -> Count, QuantizedScore),
    2 RecalTable0 by ReadGroup, 3 RecalTable1 by (ReadGroup, QualityScore),
    4 RecalTable2 by (ReadGroup, QualityScore, CovariateName, CovariateValue).
    The constructor sets these indices and column types; printing restores GATK's column order
    (CovariateValue before CovariateName).
    """

    TITLES = ['Arguments', 'Quantized', 'RecalTable0', 'RecalTable1', 'RecalTable2']

    def __init__(self, tables, version='1.1'):
        super().__init__(tables, version)
        if len(self.tables) != 5:
            raise ValueError("A RecalibrationReport should have 5 tables. This report contains %d." % len(self.tables))
        for t, title in zip(self.tables, self.TITLES):
            assert t.title == title
        t = self.tables
        t[0].data = t[0].data.set_index('Argument')
        t[1].data = t[1].data.astype({'QualityScore': np.int64, 'Count': np.int64, 'QuantizedScore': np.int64})
        t[1].data = t[1].data.set_index('QualityScore')
        t[2].data = t[2].data.set_index('ReadGroup')
        t[3].data = self._as_text(t[3].data, ['ReadGroup']).astype({'QualityScore': np.int64})
        t[3].data = t[3].data.set_index(['ReadGroup', 'QualityScore'])
        t[4].data = self._as_text(t[4].data, ['ReadGroup', 'CovariateName', 'CovariateValue']).astype({'QualityScore': np.int64})
        t[4].data = t[4].data.set_index(['ReadGroup', 'QualityScore', 'CovariateName', 'CovariateValue'])

    @staticmethod
    def _as_text(frame, columns):
        frame = frame.copy()
        for c in columns:
            frame[c] = np.asarray(frame[c].to_numpy(), dtype=str).astype(object)
        return frame

    def __str__(self):
        cov = self.tables[4]
        kept = cov.data
        cov.data = kept.swaplevel('CovariateValue', 'CovariateName')
        try:
            return super().__str__()
        finally:
            cov.data = kept

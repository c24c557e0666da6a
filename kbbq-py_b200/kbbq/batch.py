"""Host-side batching: FASTQ files -> the packed SoA buffers the C ABI takes.

Covers what the reference does read by read at the top of its two loops: read-group inference
and first-seen numbering (kbbq/recalibrate.py:27-31,59-64,143-148; kbbq/compare_reads.py:308-318),
second-in-pair inference (kbbq/compare_reads.py:304-306) and the name check of
find_corrected_sites (kbbq/recalibrate.py:17).
"""
import numpy as np

from . import fastx


class _Name:
    __slots__ = ("name",)

    def __init__(self, name):
        self.name = name


def infer_second(names):
    """fastq_infer_secondinpair over a list of names -> u8[N]."""
    return np.fromiter((n.split('_')[0][-2:] == '/2' for n in names), dtype=np.uint8, count=len(names))


def infer_rg(names, infer):
    """Read-group ints in first-seen order -> (u16[N], list of rg keys).

    infer False: every read is group 0 (kbbq/recalibrate.py:27-28).  infer True: fastq_infer_rg,
    i.e. the text after the last ':' of the second '_'-separated field, which must start with
    'RG' (AssertionError otherwise; IndexError when there is no such field).
    """
    n = len(names)
    if not infer:
        return np.zeros(n, np.uint16), [0]
    seen = {}
    out = np.empty(n, np.uint16)
    for i, name in enumerate(names):
        rgstr = name.split('_')[1]
        assert rgstr[0:2] == 'RG'
        key = rgstr.split(':')[-1]
        k = seen.get(key)
        if k is None:
            k = len(seen)
            if k > 65534:
                raise ValueError("more than 65535 read groups")
            seen[key] = k
        out[i] = k
    return out, list(seen)


class ReadBatch:
    """Packed reads: seq/qual/corr u8[N, L], rg u16[N], second u8[N], R, L; names on demand.

    Built by the native tokenizer (kbbq.fastx.NativeFastq); `source` keeps the indexed file of the
    reads so that names are only materialised when somebody asks and the recalibrated FASTQ can be
    written by the native formatter.
    """

    def __init__(self, names, seq, qual, corr, rg, second, rg_keys, source=None):
        self._names, self.seq, self.qual, self.corr = names, seq, qual, corr
        self.rg, self.second, self.rg_keys = rg, second, rg_keys
        self.N, self.L = seq.shape if seq.ndim == 2 else (0, 0)
        self.R = max(1, len(rg_keys))
        self.source = source
        self.N_all = self.N   # reads of fastq[0] (>= N when the corrected file is shorter)

    @property
    def names(self):
        if self._names is None:
            self._names = [self.source.name(i) for i in range(self.N)]
        return self._names

    @classmethod
    def from_fastq(cls, fastq, infer_rg_flag=False, need_corrected=True):
        reads = fastx.NativeFastq(fastq[0])
        n = n_all = reads.N
        corr = None
        if need_corrected:
            fixed = fastx.NativeFastq(fastq[1])
            n = min(n, fixed.N)  # zip() in the reference stops at the shorter file
            reads.check_names(fixed, n)  # find_corrected_sites, kbbq/recalibrate.py:17
        if n == 0:
            z = np.zeros((0, 0), np.uint8)
            return cls([], z, z.copy(), z.copy() if need_corrected else None, np.zeros(0, np.uint16),
                       np.zeros(0, np.uint8), [0], reads)
        rg, second, keys = reads.infer(infer_rg_flag)
        seq, qual = reads.pack(0, n)
        if need_corrected:
            if fixed.L != reads.L:
                raise ValueError("operands could not be broadcast together: corrected reads differ in length")
            corr, _ = fixed.pack(0, n)
            fixed.close()
        b = cls(None, seq, qual, corr, rg[:n], second[:n], keys, reads)
        b.N_all = n_all
        return b

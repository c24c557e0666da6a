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
    """Packed reads: names, seq/qual/corr u8[N, L], rg u16[N], second u8[N], R, L."""

    def __init__(self, names, seq, qual, corr, rg, second, rg_keys):
        self.names, self.seq, self.qual, self.corr = names, seq, qual, corr
        self.rg, self.second, self.rg_keys = rg, second, rg_keys
        self.N, self.L = seq.shape if seq.ndim == 2 else (0, 0)
        self.R = max(1, len(rg_keys))

    @classmethod
    def from_fastq(cls, fastq, infer_rg_flag=False, need_corrected=True):
        names, seq, qual = fastx.read_packed(fastq[0])
        corr = None
        if need_corrected:
            cnames, corr, _ = fastx.read_packed(fastq[1])
            n = min(len(names), len(cnames))  # zip() in the reference stops at the shorter file
            names, seq, qual, cnames, corr = names[:n], seq[:n], qual[:n], cnames[:n], corr[:n]
            for a, b in zip(names, cnames):
                assert b.startswith(a)  # find_corrected_sites, kbbq/recalibrate.py:17
            if n and corr.shape != seq.shape:
                raise ValueError("operands could not be broadcast together: corrected reads differ in length")
        rg, keys = infer_rg(names, infer_rg_flag)
        return cls(names, seq, qual, corr, rg, infer_second(names), keys)

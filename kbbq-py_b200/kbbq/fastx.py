"""FASTQ tokenising for the host side (the reference uses pysam.FastxFile, kbbq/recalibrate.py:56,141).

`FastxRecord` / `FastxFile` keep pysam's attribute names so code written against the reference
reads the same.  Whole files go through the native multithreaded tokenizer of libkbbq_b200.so
(csrc/fastq_io.cpp): `NativeFastq` wraps its handle, `read_packed` returns the packed SoA arrays the
C ABI takes.
"""
import ctypes as C
import gzip

import numpy as np

from . import _native

_ERRORS = {
    -5: lambda p: OSError("%s: cannot open / read / write" % p),
    -6: lambda p: ValueError("%s: malformed FASTQ (4-line records expected)" % p),
    -7: lambda p: ValueError("%s: reads of unequal length are not supported on this path" % p),
    -8: lambda p: IndexError("list index out of range"),      # name.split('_')[1], kbbq/compare_reads.py:316
    -9: lambda p: AssertionError("read group field must start with RG"),  # kbbq/compare_reads.py:317
    -10: lambda p: AssertionError("corrected read name does not start with the read name"),  # kbbq/recalibrate.py:17
}


def _check(rc, path=""):
    if rc == 0:
        return
    if rc in _ERRORS:
        raise _ERRORS[rc](path)
    _native.check(rc)


class NativeFastq:
    """One FASTQ file indexed by the native tokenizer (mmap; .gz is inflated into memory)."""

    def __init__(self, path, threads=0):
        self.path = str(path)
        self._lib = _native.lib()
        h = C.c_void_p()
        _check(self._lib.kbbq_fastq_open(self.path.encode(), threads, C.byref(h)), self.path)
        self._h = h
        self.threads = threads
        self.N = int(self._lib.kbbq_fastq_num_reads(h))
        self.L = int(self._lib.kbbq_fastq_read_len(h))

    def close(self):
        if self._h is not None:
            self._lib.kbbq_fastq_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def pack(self, first=0, n=None):
        """-> (seq u8[n, L], qual u8[n, L]) with qual = ASCII - 33."""
        n = self.N - first if n is None else n
        if self.L < 0:
            raise ValueError("%s: reads of unequal length are not supported on this path" % self.path)
        seq = np.empty((n, self.L), np.uint8)
        qual = np.empty((n, self.L), np.uint8)
        _check(self._lib.kbbq_fastq_pack(self._h, first, n, _native.ptr(seq), _native.ptr(qual), self.threads),
               self.path)
        return seq, qual

    def name(self, i):
        s, n = C.c_char_p(), C.c_int()
        p = C.c_void_p()
        rc = self._lib.kbbq_fastq_name(self._h, i, C.cast(C.byref(p), C.POINTER(C.c_char_p)), C.byref(n))
        _check(rc, self.path)
        return C.string_at(p.value, n.value).decode()

    def names(self):
        return [self.name(i) for i in range(self.N)]

    def infer(self, infer_rg):
        """-> (rg u16[N], second u8[N], list of read-group keys in first-seen order)."""
        rg = np.empty(self.N, np.uint16)
        second = np.empty(self.N, np.uint8)
        n_rg = C.c_int(0)
        _check(self._lib.kbbq_fastq_infer(self._h, 1 if infer_rg else 0, _native.ptr(rg), _native.ptr(second),
                                          C.byref(n_rg), self.threads), self.path)
        keys = [0]
        if infer_rg:
            keys = []
            for k in range(n_rg.value if self.N else 0):
                p, n = C.c_void_p(), C.c_int()
                _check(self._lib.kbbq_fastq_rg_key(self._h, k, C.cast(C.byref(p), C.POINTER(C.c_char_p)), C.byref(n)))
                keys.append(C.string_at(p.value, n.value).decode())
        return rg, second, keys

    def check_names(self, corrected, n):
        bad = C.c_int64(-1)
        _check(self._lib.kbbq_fastq_check_names(self._h, corrected._h, n, self.threads, C.byref(bad)), self.path)

    def write(self, fd, out_qual, first=0, n=None):
        """'@name / seq / + / qual' records of reads [first, first + n) with the given phred values."""
        n = self.N - first if n is None else n
        out_qual = np.ascontiguousarray(out_qual, dtype=np.uint8)
        assert out_qual.size == n * max(self.L, 0)
        _check(self._lib.kbbq_fastq_write(fd, self._h, first, n, _native.ptr(out_qual), self.threads), self.path)


class FastxRecord:
    def __init__(self, name=None, sequence=None, quality=None, comment=None):
        self.name = name
        self.sequence = sequence
        self.quality = quality
        self.comment = comment

    def get_quality_array(self, offset=33):
        return [ord(ch) - offset for ch in self.quality]

    def __str__(self):
        head = self.name if not self.comment else "%s %s" % (self.name, self.comment)
        return "@%s\n%s\n+\n%s" % (head, self.sequence, self.quality)


def _open(path):
    if str(path).endswith(".gz"):
        return gzip.open(path, "rt")
    return open(path)


class FastxFile:
    def __init__(self, filename, mode="r"):
        self._fh = _open(filename)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self._fh.close()

    def close(self):
        self._fh.close()

    def __iter__(self):
        fh = self._fh
        while True:
            header = fh.readline()
            if not header:
                return
            seq = fh.readline().rstrip("\r\n")
            fh.readline()
            qual = fh.readline().rstrip("\r\n")
            fields = header.rstrip("\r\n")[1:].split(None, 1)
            yield FastxRecord(fields[0] if fields else "", seq, qual, fields[1] if len(fields) > 1 else None)


def read_packed(path):
    """Parse a 4-line FASTQ file -> (names list[str], seq u8[N, L], qual u8[N, L] phred).

    All reads must have the same length: the reference's table code cannot handle anything else
    (ValueError / IndexError there); a ValueError is raised here.
    """
    f = NativeFastq(path)
    try:
        if f.N == 0:
            return [], np.zeros((0, 0), np.uint8), np.zeros((0, 0), np.uint8)
        seq, qual = f.pack()
        return f.names(), seq, qual
    finally:
        f.close()

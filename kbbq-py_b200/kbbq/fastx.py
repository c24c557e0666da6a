"""FASTQ tokenising for the host side (the reference uses pysam.FastxFile, kbbq/recalibrate.py:56,141).

`FastxRecord` / `FastxFile` keep pysam's attribute names so code written against the reference
reads the same; `read_packed` parses a whole file into the packed SoA arrays the C ABI takes.
"""
import gzip

import numpy as np


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
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "rb") as fh:
        data = fh.read()
    lines = data.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    if len(lines) % 4:
        raise ValueError("%s: truncated FASTQ (line count %d is not a multiple of 4)" % (path, len(lines)))
    n = len(lines) // 4
    names = [ln[1:].split(None, 1)[0].decode() if len(ln) > 1 else "" for ln in lines[0::4]]
    if n == 0:
        return names, np.zeros((0, 0), np.uint8), np.zeros((0, 0), np.uint8)
    seqs = [ln.rstrip(b"\r") for ln in lines[1::4]]
    quals = [ln.rstrip(b"\r") for ln in lines[3::4]]
    L = len(seqs[0])
    if any(len(s) != L for s in seqs) or any(len(q) != L for q in quals):
        raise ValueError("%s: reads of unequal length are not supported on this path" % path)
    seq = np.frombuffer(b"".join(seqs), dtype=np.uint8).reshape(n, L)
    qual = np.frombuffer(b"".join(quals), dtype=np.uint8).reshape(n, L) - np.uint8(33)
    return names, seq, qual

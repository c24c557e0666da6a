"""Stand-in for the parts of pysam the reference's FASTQ path touches (FastxFile / FastxRecord).

Test infrastructure only: lets `oracle/ref_shim` import the unmodified reference in a container
that has no pysam.  Never imported by the product package.
"""


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


class FastxFile:
    def __init__(self, filename, mode="r"):
        self._fh = open(filename)

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
            seq = fh.readline().rstrip("\n")
            fh.readline()
            qual = fh.readline().rstrip("\n")
            fields = header.rstrip("\n")[1:].split(None, 1)
            yield FastxRecord(fields[0], seq, qual, fields[1] if len(fields) > 1 else None)

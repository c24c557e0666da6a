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


# ---- in-memory stand-ins for the BAM side (tests/golden/make_golden_bam.py) -----------------------
# Only the attributes the reference's BQSR emulation reads (kbbq/gatk/bqsr.py:23-212,
# kbbq/gatk/applybqsr.py:46-78, kbbq/compare_reads.py:84-135,332-340) are provided.

class AlignedSegment:
    """A mapped read described by plain Python values."""

    def __init__(self, query_name, query_sequence, qualities, cigartuples, reference_name, reference_start,
                 is_reverse=False, is_read2=False, tags=None):
        self.query_name = query_name
        self.query_sequence = query_sequence
        self.query_qualities = list(qualities)
        self.cigartuples = list(cigartuples)
        self.reference_name = reference_name
        self.reference_start = reference_start
        self.is_reverse = is_reverse
        self.is_read2 = is_read2
        self.is_read1 = not is_read2
        self.is_paired = True
        self.is_unmapped = False
        self.mate_is_unmapped = False
        self.mate_is_reverse = not is_reverse
        self.tlen = 0                      # no adaptor boundary unless a test sets one
        self.template_length = 0
        self.next_reference_start = reference_start
        self._tags = dict(tags or {})

    @property
    def query_length(self):
        return len(self.query_sequence)

    def _clip(self, side):
        ops = self.cigartuples if side == 0 else self.cigartuples[::-1]
        n = 0
        for op, l in ops:
            if op == 4:
                n += l
            elif op != 5:
                break
        return n

    @property
    def query_alignment_start(self):
        return self._clip(0)

    @property
    def query_alignment_end(self):
        return self.query_length - self._clip(1)

    @property
    def query_alignment_length(self):
        return self.query_alignment_end - self.query_alignment_start

    @property
    def reference_end(self):
        return self.reference_start + sum(l for op, l in self.cigartuples if op in (0, 2, 3, 7, 8))

    def get_tag(self, tag):
        return self._tags[tag]

    def set_tag(self, tag, value):
        self._tags[tag] = value

    def get_aligned_pairs(self):
        pairs, q, r = [], 0, self.reference_start
        for op, l in self.cigartuples:
            if op in (0, 7, 8):
                pairs += [(q + i, r + i) for i in range(l)]
                q += l
                r += l
            elif op in (1, 4):
                pairs += [(q + i, None) for i in range(l)]
                q += l
            elif op in (2, 3):
                pairs += [(None, r + i) for i in range(l)]
                r += l
        return pairs


class _Header:
    def __init__(self, d):
        self._d = d

    def as_dict(self):
        return self._d


class _IndexStat:
    def __init__(self, total):
        self.total = total


class AlignmentFile:
    """An iterator over AlignedSegment objects with a header: AlignmentFile(reads=[...], header={...})."""

    def __init__(self, filename=None, mode="r", reads=None, header=None):
        self._reads = list(reads or [])
        self._it = iter(self._reads)
        self.header = _Header(header or {})

    def get_index_statistics(self):
        return [_IndexStat(len(self._reads))]

    def __iter__(self):
        return self

    def __next__(self):
        return next(self._it)


_FASTA_REGISTRY = {}


def register_fasta(name, contigs):
    """Make FastaFile(name) serve `contigs` = {contig: sequence string}."""
    _FASTA_REGISTRY[name] = dict(contigs)


class FastaFile:
    def __init__(self, filename):
        self._contigs = _FASTA_REGISTRY[filename]
        self.references = list(self._contigs)

    def fetch(self, reference=None, start=None, end=None):
        s = self._contigs[reference]
        return s[start:end]

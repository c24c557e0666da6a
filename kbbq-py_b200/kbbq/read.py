"""Mirror of the reference's kbbq/read.py: the :class:`ReadData` value class (host batching API).

Same constructor, attributes, class-level read-group registry and per-read extractor methods as
the reference (kbbq/read.py:21-378).  One ReadData is one read on the host; it is API surface and
bookkeeping -- the arithmetic of the hot path runs on packed batches of reads
(:func:`pack_reads` -> :meth:`kbbq.covariate.CovariateData.consume_packed` -> kbbq_build on the GPU).
BAM constructors need pysam objects and stay duck-typed (from_bamread, load_rgs_from_bamfile); the
BAM path itself is out of scope.
"""
import numpy as np

from . import compare_reads


class ReadData:
    """seq / qual / skips / errors arrays of one read plus name, read group and pair flag.

    Class attributes (a process-global registry, as in the reference, kbbq/read.py:70-94):
    rg_to_pu, rg_to_int (first-seen order), numrgs.
    """

    rg_to_pu = dict()
    rg_to_int = dict()
    numrgs = 0

    def __init__(self, seq, qual, skips, name, rg, second, errors):
        self.seq = seq
        self.qual = qual
        self.skips = skips
        self.name = name
        self.rg = rg
        cls = self.__class__
        if rg not in ReadData.rg_to_pu:
            cls.rg_to_pu[rg] = rg
            cls.rg_to_int[rg] = ReadData.numrgs
            cls.numrgs = ReadData.numrgs + 1
        self.second = second
        self.errors = errors

    # ---- constructors ---------------------------------------------------------------------------
    @classmethod
    def from_bamread(cls, bamread, use_oq=False):
        """reference: kbbq/read.py:100-141 (reverse-strand reads are flipped and complemented)."""
        seq = np.array(list(bamread.query_sequence), dtype=np.str_)
        qual = bamread_get_quals(bamread, use_oq)
        if bamread.is_reverse:
            seq = compare_reads.Dinucleotide.veccomplement(np.flip(seq), 'N')
            qual = np.flip(qual)
        n = len(seq)
        return cls(seq=seq, qual=qual, skips=np.zeros(n, dtype=bool), name=bamread.query_name,
                   rg=bamread.get_tag('RG') if bamread.has_tag('RG') else None,
                   second=bamread.is_read2, errors=np.zeros(n, dtype=bool))

    @classmethod
    def from_fastq(cls, fastqread, rg=None, second=None, namedelimiter='_'):
        """reference: kbbq/read.py:143-196.  RG = text after the last ':' of the last
        delimiter-separated field that starts with 'RG:'; second = the first field ends in '/2';
        the stored name is the first field without its /1 or /2."""
        seq = np.array(list(fastqread.sequence), dtype=np.str_)
        n = len(seq)
        fields = fastqread.name.split(sep=namedelimiter)
        if rg is None:
            tagged = [f.split(':')[-1] for f in fields if f[0:3] == 'RG:']
            if tagged:
                rg = tagged[-1]
        if second is None:
            second = (fields[0][-2:] == '/2')
        if fields[0].endswith(('/1', '/2')):
            fields[0] = fields[0][:-2]
        return cls(seq=seq, qual=np.array(fastqread.get_quality_array(), dtype=int),
                   skips=np.zeros(n, dtype=bool), name=fields[0], rg=rg, second=second,
                   errors=np.zeros(n, dtype=bool))

    @classmethod
    def load_rgs_from_bamfile(cls, bamfileobj):
        """reference: kbbq/read.py:198-213."""
        for rg in bamfileobj.header.as_dict()['RG']:
            cls.rg_to_pu[rg['ID']] = rg['PU']
            cls.rg_to_int[rg['ID']] = cls.numrgs
            cls.numrgs = cls.numrgs + 1

    # ---- small accessors ------------------------------------------------------------------------
    def str_qual(self, offset=33):
        return list((self.qual + offset).astype(np.uint32).view('U1'))

    def canonical_name(self):
        return self.name + ("/2" if self.second else "/1")

    def get_rg_int(self):
        return self.__class__.rg_to_int[self.rg]

    def get_pu(self):
        return self.__class__.rg_to_pu[self.rg]

    def not_skipped_errors(self):
        return np.logical_and(self.errors, ~self.skips)

    def __len__(self):
        return len(self.seq)

    # ---- covariate extractors: (values at non-skipped errors, values at non-skipped sites) ------
    def get_rg_errors(self):
        rg = np.broadcast_to(self.get_rg_int(), len(self))
        return rg[self.not_skipped_errors()], rg[~self.skips]

    def get_q_errors(self):
        return self.qual[self.not_skipped_errors()], self.qual[~self.skips]

    def get_cycle_array(self):
        """0..L-1 for read 1, -1..-L for read 2 (an index from the end of the 2L cycle axis)."""
        cycle = np.arange(len(self))
        return np.negative(cycle + 1) if self.second else cycle

    def get_cycle_errors(self):
        cycle = self.get_cycle_array()
        return cycle[self.not_skipped_errors()], cycle[~self.skips]

    def get_dinucleotide_array(self, minscore=6):
        """reference: kbbq/read.py:336-353; same arithmetic as compare_reads.generic_dinuc_covariate."""
        return compare_reads.generic_dinuc_covariate(np.asarray(self.seq, dtype=np.str_), np.asarray(self.qual),
                                                     minscore)

    def get_dinuc_errors(self, minscore=6):
        dinuc = self.get_dinucleotide_array(minscore)
        dvalid = np.logical_and(dinuc != -1, ~self.skips)
        return dinuc[np.logical_and(dvalid, self.errors)], dinuc[dvalid]


def bamread_get_oq(read, offset=33):
    """reference: kbbq/read.py:380-396."""
    oq = np.array(list(read.get_tag('OQ')), dtype=np.str_)
    return np.array(oq.view(np.uint32) - offset, dtype=np.uint32)


def bamread_get_quals(read, use_oq=False):
    """reference: kbbq/read.py:398-415."""
    if use_oq:
        return bamread_get_oq(read)
    return np.array(read.query_qualities, dtype=int)


def pack_reads(reads, minscore=6):
    """Pack ReadData objects of equal length into the SoA buffers of the C ABI.

    Returns (seq u8[N,L], qual u8[N,L], corr u8[N,L], rg u16[N], second u8[N]).  A base flagged in
    `errors` gets a corrected base that differs from the original (the kernels only compare the two
    bytes); a base flagged in `skips` gets quality 0, i.e. below `minscore`, which is how the packed
    path expresses "do not tally" (kbbq/recalibrate.py:96: skips = q < minscore).
    """
    reads = list(reads)
    if not reads:
        z = np.zeros((0, 0), np.uint8)
        return z, z.copy(), z.copy(), np.zeros(0, np.uint16), np.zeros(0, np.uint8)
    L = len(reads[0])
    if any(len(r) != L for r in reads):
        raise ValueError("the packed path needs reads of one length")
    if minscore < 1:
        raise ValueError("minscore must be at least 1 to express skipped bases")
    n = len(reads)
    seq = np.empty((n, L), np.uint8)
    qual = np.empty((n, L), np.uint8)
    for i, r in enumerate(reads):
        seq[i] = np.asarray(r.seq, dtype=np.str_).view(np.uint32).astype(np.uint8)
        q = np.asarray(r.qual).astype(np.int64)
        if np.any((q < 0) | (q > 255)):
            raise IndexError("quality out of range")
        qual[i] = np.where(np.asarray(r.skips, dtype=bool), 0, q)
    errors = np.stack([np.asarray(r.errors, dtype=bool) for r in reads])
    corr = np.where(errors, np.where(seq == ord('A'), ord('C'), ord('A')).astype(np.uint8), seq)
    rg = np.fromiter((r.get_rg_int() for r in reads), dtype=np.uint16, count=n)
    second = np.fromiter((1 if r.second else 0 for r in reads), dtype=np.uint8, count=n)
    return seq, qual, corr, rg, second

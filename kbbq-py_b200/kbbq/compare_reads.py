"""Mirror of the hot-path part of the reference's kbbq/compare_reads.py (lines 141-328).

Same names, argument meaning and error behaviour.  The arithmetic of the model
(:func:`gatk_delta_q`) and of the apply (:func:`recalibrate_fastq`) runs on the GPU through
the C ABI; the small covariate helpers that only produce index arrays for a single read stay
numpy on the host (they are API surface, the batched kernels fuse the same index arithmetic).
The BAM-side helpers find_read_errors, bamread_get_oq and get_rg_to_pu (host work on one pysam
record: CIGAR walk, tag decoding) are here for kbbq.gatk.bqsr / kbbq.gatk.applybqsr; they take any
object with pysam's AlignedSegment attributes and never import pysam themselves.
Out of scope (regression experiments, file readers): load_positions, get_var_sites,
train_regression, regression_recalibrate.
"""
import numpy as np

from . import _native


class RescaledNormal:
    """Cached prior of the Bayesian model (reference: kbbq/compare_reads.py:141-191).

    prior_dist[d] = log(.9 * exp(-((d/.5)**2)/2)), -inf once exp() underflows (d >= 19).
    """

    maxscore = 42
    possible_diffs = np.arange(maxscore + 1, dtype=np.int_)
    prior_dist = np.zeros(possible_diffs.shape[0], dtype=np.longdouble)
    _old = np.seterr(all='raise')
    for _i in range(possible_diffs.shape[0]):
        try:
            prior_dist[_i] = np.log(.9 * np.exp(-((possible_diffs[_i] / .5) ** 2) / 2))
        except FloatingPointError:
            prior_dist[_i] = -np.inf
    np.seterr(**_old)
    del _i, _old

    @classmethod
    def prior(cls, difference):
        return cls.prior_dist[difference]


class Dinucleotide:
    """Dinucleotide <-> int map (reference: kbbq/compare_reads.py:193-233)."""

    nucleotides = ['A', 'T', 'G', 'C']
    complement = {'A': 'T', 'T': 'A', 'G': 'C', 'C': 'G'}
    dinucs = [i + j for i in ['A', 'T', 'G', 'C'] for j in ['A', 'T', 'G', 'C']]
    dinuc_to_int = dict(zip(dinucs, range(len(dinucs))))

    vectorized_get = np.vectorize(dinuc_to_int.get, otypes=[int])
    vectorized_complement = np.vectorize(complement.get, otypes=[np.str_])

    @classmethod
    def vecget(cls, *args, **kwargs):
        return cls.vectorized_get(*args, **kwargs)

    @classmethod
    def veccomplement(cls, *args, **kwargs):
        return cls.vectorized_complement(*args, **kwargs)


def gatk_delta_q(prior_q, numerrs, numtotal, maxscore=42):
    """Shift of the MAP quality from the prior, per cell (reference: kbbq/compare_reads.py:235-260).

    Runs on the GPU (kbbq_delta_q).  Only maxscore == 42 is supported, as on the whole path.
    """
    prior_q, numerrs, numtotal = np.asarray(prior_q), np.asarray(numerrs), np.asarray(numtotal)
    assert prior_q.shape == numerrs.shape == numtotal.shape
    if maxscore != 42:
        raise NotImplementedError("only maxscore = 42 is supported")
    if np.any(prior_q < 0) or np.any(prior_q > maxscore):
        raise IndexError("prior_q outside 0..%d" % maxscore)
    return _native.delta_q_host(prior_q, numerrs, numtotal).reshape(prior_q.shape)


def p_to_q(p, maxscore=42):
    """reference: kbbq/compare_reads.py:262-267 (truncating, p == 0 -> maxscore, clipped)."""
    p = np.asarray(p)
    q = np.zeros(p.shape, dtype=int)
    nz = p != 0
    q[nz] = (-10.0 * np.log10(p[nz])).astype(int)
    q[~nz] = maxscore
    return np.clip(q, 0, maxscore).copy()


def q_to_p(q):
    """reference: kbbq/compare_reads.py:269-271."""
    return np.array(np.power(10.0, -(np.asarray(q) / 10.0)), dtype=np.longdouble, copy=True)


# ---- generic covariate functions (index arrays for one read; host-side API surface) -------------

def generic_cycle_covariate(sequencelen, secondinpair=False):
    """reference: kbbq/compare_reads.py:275-279."""
    cycle = np.arange(sequencelen)
    if secondinpair:
        cycle = np.negative(cycle + 1)
    return cycle


_CODE = np.full(256, -1, dtype=int)
for _k, _b in enumerate('ATGC'):
    _CODE[ord(_b)] = _k
del _k, _b


def generic_dinuc_covariate(sequences, quals, minscore=6):
    """reference: kbbq/compare_reads.py:281-293.  -1 for the first base, a base below minscore, or
    an N in either position; a non-ACGTN base at a valid site is a TypeError, as in the reference
    (Dinucleotide.vecget returns None there)."""
    sequences, quals = np.ascontiguousarray(sequences), np.asarray(quals)
    assert sequences.shape == quals.shape
    assert sequences.dtype == np.dtype('U1')
    codes = _CODE[np.clip(sequences.view(np.uint32).reshape(sequences.shape), 0, 255)]
    codes[sequences.view(np.uint32).reshape(sequences.shape) > 255] = -1
    dinuccov = np.zeros(sequences.shape, dtype=int)
    dinuccov[..., 0] = -1
    is_n = (sequences[..., 1:] == 'N')
    follows_n = (sequences[..., :-1] == 'N')
    invalid = np.logical_or(quals[..., 1:] < minscore, np.logical_or(is_n, follows_n))
    pair = 4 * codes[..., :-1] + codes[..., 1:]
    bad = np.logical_and(~invalid, np.logical_or(codes[..., :-1] < 0, codes[..., 1:] < 0))
    if np.any(bad):
        raise TypeError("int() argument must be a string, a bytes-like object or a number, not 'NoneType'")
    dinuccov[..., 1:] = np.where(invalid, -1, pair)
    return dinuccov


# ---- FASTQ reads -------------------------------------------------------------------------------

def fastq_cycle_covariates(read, secondinpair=False):
    return generic_cycle_covariate(len(read.sequence), secondinpair)


def fastq_dinuc_covariates(read, minscore=6):
    quals = np.array(read.get_quality_array(), dtype=int)
    return generic_dinuc_covariate(np.array(list(read.sequence)), quals, minscore)


def fastq_infer_secondinpair(read):
    """reference: kbbq/compare_reads.py:304-306."""
    namestr = read.name.split(sep='_')[0]
    return namestr[-2:] == '/2'


def fastq_infer_rg(read):
    """reference: kbbq/compare_reads.py:308-318."""
    rgstr = read.name.split(sep='_')[1]
    assert rgstr[0:2] == 'RG'
    return rgstr.split(':')[-1]


def recalibrate_fastq(read, meanq, globaldeltaq, qscoredeltaq, positiondeltaq, dinucdeltaq, rg,
                      dinuc_to_int, secondinpair=False, minscore=6, maxscore=42):
    """Recalibrated qualities of ONE read (reference: kbbq/compare_reads.py:320-328).

    The gather-sum runs on the GPU (kbbq_apply) on a batch of one.  `rg` may be an int or a
    one-element array (the reference's tests pass np.array([0])).  The table shapes are taken
    from the arguments, as Python indexing does in the reference: a quality beyond the q axis is
    an IndexError, dinuc -1 gathers the last dinuc column.
    """
    seq = np.frombuffer(read.sequence.encode(), dtype=np.uint8)
    qual = np.array(read.get_quality_array(), dtype=int)
    L = seq.size
    qscoredeltaq, positiondeltaq, dinucdeltaq = (np.asarray(a) for a in (qscoredeltaq, positiondeltaq, dinucdeltaq))
    meanq, globaldeltaq = np.atleast_1d(np.asarray(meanq)), np.atleast_1d(np.asarray(globaldeltaq))
    g = int(np.asarray(rg).reshape(-1)[0])
    if np.any(qual < 0) or np.any(qual > 255):
        raise IndexError("quality out of range")
    if positiondeltaq.shape[2] != 2 * L:
        # a shorter / longer cycle axis changes what negative cycles mean; only the 2L layout the
        # build produces is supported on the GPU path
        raise IndexError("positiondeltaq cycle axis must have length 2 * len(read)")
    out = _native.apply_host(seq, qual.astype(np.uint8), None, np.array([1 if secondinpair else 0], np.uint8),
                             L, 1, meanq[g:g + 1], globaldeltaq[g:g + 1], qscoredeltaq[g:g + 1],
                             positiondeltaq[g:g + 1], dinucdeltaq[g:g + 1], minscore=minscore)
    out = out.reshape(L).astype(np.int8).astype(int)  # the device keeps the low 8 bits of the sum
    return out


# ---- BAM reads (host side of SURVEY.md section 8 row f3) -----------------------------------------

def find_read_errors(read, ref, variable):
    """Errors and sites to skip of one aligned read (reference: kbbq/compare_reads.py:84-135).

    Walks the CIGAR: aligned bases (M, =, X) are errors where they differ from `ref` and skipped
    where `variable` marks the reference position; an insertion is skipped when the positions on
    both sides are; a deletion or N skips the base in front of it if it spans a variable position;
    soft clips are skipped; hard clips and pads consume nothing.  `ref` / `variable` map contig ->
    array over the contig.  Returns (errors, skips), boolean arrays over the read.
    """
    seq = np.frombuffer(read.query_sequence.encode(), dtype=np.uint8)
    errors = np.zeros(seq.shape, dtype=bool)
    skips = np.zeros(seq.shape, dtype=bool)
    lo, hi = read.reference_start, read.reference_end
    var = np.asarray(variable[read.reference_name][lo:hi], dtype=bool)
    refseq = ref[read.reference_name][lo:hi]
    if not (isinstance(refseq, np.ndarray) and refseq.dtype == np.uint8):
        refseq = np.frombuffer(''.join(refseq).encode(), dtype=np.uint8)
    q = r = 0
    for op, n in read.cigartuples:
        if op in (0, 7, 8):
            errors[q:q + n] = refseq[r:r + n] != seq[q:q + n]
            skips[q:q + n] = var[r:r + n]
            q += n
            r += n
        elif op == 1:
            skips[q:q + n] = var[r - 1] and var[r]
            q += n
        elif op in (2, 3):
            skips[q - 1] = skips[q - 1] or bool(np.any(var[r:r + n]))
            r += n
        elif op == 4:
            skips[q:q + n] = True
            q += n
        elif op in (5, 6):
            pass
        else:
            raise ValueError("Unrecognized Cigar Operation " + str(op) + " In Read\n" + str(read))
    return errors, skips


def bamread_get_oq(read):
    """Original qualities from the OQ tag (reference: kbbq/compare_reads.py:332-336)."""
    return np.frombuffer(read.get_tag('OQ').encode(), dtype=np.uint8).astype(int) - 33


def get_rg_to_pu(bamfileobj):
    """{read group ID: platform unit} from the header (reference: kbbq/compare_reads.py:338-340)."""
    return {rg['ID']: rg['PU'] for rg in bamfileobj.header.as_dict()['RG']}

#!/usr/bin/env python3
"""Golden vectors for the BAM side, made by the UNMODIFIED reference.

Run in the build container only (needs /root/reference, imported through oracle/ref_shim):

    python tests/golden/make_golden_bam.py

pysam is absent here, so the reads are in-memory stand-ins (oracle/ref_shim/stubs/pysam:
AlignedSegment / AlignmentFile / FastaFile carry exactly the attributes the reference reads).  On
them the reference's own functions run unchanged:
  kbbq.gatk.bqsr.bam_to_bqsr_covariates      (kbbq/gatk/bqsr.py:52-123)  -> the nine tables
  kbbq.compare_reads.find_read_errors         (kbbq/compare_reads.py:84-135) and
  kbbq.gatk.bqsr.trim_bamread                 (kbbq/gatk/bqsr.py:158-212) -> per-base error / skip masks
  kbbq.gatk.applybqsr.get_delta_qs / recalibrate_bamread (kbbq/gatk/applybqsr.py:65-103) -> new quals
The .npz keeps the packed inputs the C ABI takes (seq, qual = OQ, err, skip u8[N, L]; rg, flags,
aln_start, aln_end per read) and the reference's outputs.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_shim  # noqa: E402

rc, cr, ab = ref_shim.load()
import pysam  # noqa: E402  (the stand-in)
from kbbq.gatk import bqsr  # noqa: E402  (the reference's)

BASES = "ATGC"


def make_case(name, n_reads, L, R, seed, p_n=0.01, p_err=0.03, p_var=0.02):
    rng = np.random.default_rng(seed)
    reflen = 4000
    ref = "".join(rng.choice(list(BASES), size=reflen))
    variable = sorted(rng.choice(reflen, size=int(p_var * reflen), replace=False).tolist())
    pysam.register_fasta(name + ".fa", {"chr1": ref})
    rg_ids = ["rg%d" % i for i in range(R)]
    header = {"RG": [{"ID": rg, "PU": "unit." + rg} for rg in rg_ids]}
    reads = []
    for i in range(n_reads):
        reverse, read2 = bool(rng.integers(2)), bool(rng.integers(2))
        # CIGAR: optional soft clips, matches with an optional insertion / deletion in the middle
        sl = int(rng.integers(0, 6)) if rng.random() < 0.4 else 0
        sr = int(rng.integers(0, 6)) if rng.random() < 0.4 else 0
        body = L - sl - sr
        kind = rng.random()
        cig = [(4, sl)] if sl else []
        if kind < 0.2:      # insertion
            il = int(rng.integers(1, 4))
            m1 = int(rng.integers(5, body - il - 5))
            cig += [(0, m1), (1, il), (0, body - il - m1)]
        elif kind < 0.4:    # deletion
            dl = int(rng.integers(1, 4))
            m1 = int(rng.integers(5, body - 5))
            cig += [(0, m1), (2, dl), (0, body - m1)]
        else:
            cig += [(0, body)]
        if sr:
            cig += [(4, sr)]
        reflen_used = sum(l for op, l in cig if op in (0, 2, 3))
        start = int(rng.integers(1, reflen - reflen_used - 1))
        # sequence: reference with errors; clipped / inserted bases random
        seq, r = [], start
        for op, l in cig:
            if op == 0:
                seq += list(ref[r:r + l])
                r += l
            elif op in (1, 4):
                seq += list(rng.choice(list(BASES), size=l))
            elif op == 2:
                r += l
        seq = np.array(seq)
        flip = rng.random(L) < p_err
        seq[flip] = rng.choice(list(BASES), size=int(flip.sum()))
        seq[rng.random(L) < p_n] = "N"
        mu = rng.normal(33, 4)
        q = np.clip(np.rint(mu - 10 * (np.arange(L) / L) ** 2 + rng.normal(0, 3, size=L)), 2, 42).astype(int)
        if rng.random() < 0.15:
            q[L - int(rng.integers(1, L // 3)):] = 2
        q[seq == "N"] = 2
        oq = "".join(chr(int(x) + 33) for x in q)
        read = pysam.AlignedSegment("r%d" % i, "".join(seq), (q // 2).tolist(), cig, "chr1", start, reverse, read2,
                                    {"OQ": oq, "RG": rg_ids[int(rng.integers(R))]})
        if rng.random() < 0.3:  # an insert shorter than the read: adaptor bases to trim (kbbq/gatk/bqsr.py:131-212)
            if reverse:
                read.next_reference_start = int(rng.integers(start + 1, read.reference_end - 1))
                read.tlen = -int(read.reference_end - read.next_reference_start)
            else:
                read.tlen = int(rng.integers(5, reflen_used - 1))
                read.next_reference_start = start
            read.template_length = read.tlen
        reads.append(read)
    var_pos = {"chr1": variable}

    tables = bqsr.bam_to_bqsr_covariates(pysam.AlignmentFile(reads=reads, header=header), name + ".fa", var_pos)
    deltas = ab.get_delta_qs(*tables)

    # what the device entry points take
    refd = {"chr1": np.array(list(ref), dtype=np.str_)}
    fullskips = {"chr1": np.zeros(reflen, dtype=bool)}
    fullskips["chr1"][np.array(variable, dtype=int)] = True
    rg_to_int = {rg: i for i, rg in enumerate(rg_ids)}
    N = n_reads
    seq_a, qual_a = np.zeros((N, L), np.uint8), np.zeros((N, L), np.uint8)
    err_a, skip_a = np.zeros((N, L), np.uint8), np.zeros((N, L), np.uint8)
    rg_a, flags = np.zeros(N, np.uint16), np.zeros(N, np.uint8)
    a0, a1 = np.zeros(N, np.uint16), np.zeros(N, np.uint16)
    outq = np.zeros((N, L), np.int64)
    for i, read in enumerate(reads):
        e, s = cr.find_read_errors(read, refd, fullskips)
        s = np.logical_or(s, bqsr.trim_bamread(read))
        seq_a[i] = np.frombuffer(read.query_sequence.encode(), np.uint8)
        qual_a[i] = cr.bamread_get_oq(read)
        err_a[i], skip_a[i] = e, s
        rg_a[i] = rg_to_int[read.get_tag("RG")]
        flags[i] = (1 if read.is_read2 else 0) | (2 if read.is_reverse else 0)
        a0[i], a1[i] = read.query_alignment_start, read.query_alignment_end
        outq[i] = ab.recalibrate_bamread(read, tables[0], *deltas, rg_to_int)
    keys = ("meanq", "rg_errs", "rg_total", "q_errs", "q_total", "pos_errs", "pos_total", "dinuc_errs", "dinuc_total")
    # enough to rebuild the stand-in reads in the tests (host-side CIGAR walk and trimming)
    maxops = max(len(r.cigartuples) for r in reads)
    cigar = np.full((N, maxops, 2), -1, np.int32)
    for i, r in enumerate(reads):
        cigar[i, :len(r.cigartuples)] = r.cigartuples
    np.savez_compressed(os.path.join(HERE, name + ".npz"), seq=seq_a, qual=qual_a, err=err_a, skip=skip_a, rg=rg_a,
                        flags=flags, aln_start=a0, aln_end=a1, L=L, R=R, outq=outq,
                        ref=np.frombuffer(ref.encode(), np.uint8), variable=np.array(variable, np.int64), cigar=cigar,
                        ref_start=np.array([r.reference_start for r in reads], np.int64),
                        tlen=np.array([r.tlen for r in reads], np.int64),
                        next_start=np.array([r.next_reference_start for r in reads], np.int64),
                        bamq=np.array([r.query_qualities for r in reads], np.uint8),
                        rgdq=deltas[0], qdq=deltas[1], posdq=deltas[2], dindq=deltas[3],
                        **dict(zip(keys, tables)))
    print(name, "reads", N, "valid bases", int(tables[2].sum()), "errors", int(tables[1].sum()))


if __name__ == "__main__":
    make_case("bam_mixed_r2", 400, 60, 2, 501)
    make_case("bam_long_r3", 250, 151, 3, 502, p_n=0.02)

// bam.cuh -- the BAM side of the path: the same tally and the same
// gather as the FASTQ kernels, for aligned reads.
//
// build_bam_kernel replaces the loop body of bqsr.bam_to_bqsr_covariates (kbbq/gatk/bqsr.py:86-118)
// with bamread_bqsr_cycle (:23-31) and bamread_bqsr_dinuc (:33-50) fused in: errors and sites to skip
// come from the host (CIGAR walk against the reference, known variant sites, adaptor trimming:
// compare_reads.find_read_errors, bqsr.trim_bamread) as one byte per base; the kernel adds the
// q < minscore and N skips (:94-96), numbers cycles inside the aligned part of the read, backwards
// on the reverse strand, and takes the dinucleotide from the reverse complement there.
// apply_bam_kernel replaces applybqsr.recalibrate_bamread (kbbq/gatk/applybqsr.py:65-78) with
// bamread_cycle_covariates / bamread_dinuc_covariates (:46-63): no clipping, whole read.
//
// One thread per base, 64-bit global reductions / gathers from the int64 delta tables: these reads
// arrive one BAM record at a time from a host parser, so the kernels are sized for correctness and
// streaming batches, not for the shared-memory machinery of build.cuh / apply.cuh.
#pragma once
#include "common.cuh"

namespace kbbq {

constexpr int BAM_FLAG_READ2 = 1;
constexpr int BAM_FLAG_REVERSE = 2;

// reference order A T G C -> 0 1 2 3 (kbbq/compare_reads.py:199), N -> 4, anything else -> 5
__device__ __forceinline__ int bam_code(uint8_t b) {
    return b == 'A' ? 0 : b == 'T' ? 1 : b == 'G' ? 2 : b == 'C' ? 3 : b == 'N' ? 4 : 5;
}

// Dinucleotide index of base i of a read whose covariate window is [lo, hi): -1 when there is no
// previous base in read direction, the quality is below minscore or either base is N.  On the
// reverse strand the window is reverse-complemented first (complement.get(x, 'N'): anything that is
// not ACGT becomes N, kbbq/gatk/bqsr.py:40, kbbq/gatk/applybqsr.py:56), so the previous base is the
// complement of the NEXT one in stored order.  *bad is set for a non-ACGTN base at a usable site of a
// forward read (TypeError in the reference).
__device__ __forceinline__ int bam_dinuc(const uint8_t *row, int i, int lo, int hi, bool reverse, unsigned int q,
                                         int minscore, bool *bad) {
    if ((int)q < minscore) return -1;
    int cur = bam_code(row[i]), prev;
    if (reverse) {
        if (i + 1 >= hi) return -1;
        prev = bam_code(row[i + 1]);
        if (cur >= 4 || prev >= 4) return -1;
        return 4 * (prev ^ 1) + (cur ^ 1);   // A <-> T, G <-> C
    }
    if (i - 1 < lo) return -1;
    prev = bam_code(row[i - 1]);
    if (cur == 4 || prev == 4) return -1;
    if (cur == 5 || prev == 5) { *bad = true; return -1; }
    return 4 * prev + cur;
}

struct BuildBamArgs {
    const uint8_t *seq, *qual, *err, *skip;
    const uint16_t *rg;
    const uint8_t *flags;
    const uint16_t *aln_start, *aln_end;
    long long N;
    int L, R, minscore;
    unsigned long long *pos_errs, *pos_total, *din_errs, *din_total;
    int *status;
};

__global__ void build_bam_kernel(BuildBamArgs a) {
    const long long total = a.N * a.L;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long r = idx / a.L;
        const int i = (int)(idx - r * a.L);
        const unsigned int q = a.qual[idx];
        const uint8_t s = a.seq[idx];
        if (a.skip && a.skip[idx]) continue;
        if ((int)q < a.minscore || s == 'N') continue;           // kbbq/gatk/bqsr.py:94-96
        if (q > NQ - 1) { atomicOr(a.status, KBBQ_FLAG_QUAL_RANGE); continue; }
        const unsigned int g = a.rg ? a.rg[r] : 0;
        if (g >= (unsigned int)a.R) { atomicOr(a.status, KBBQ_FLAG_RG_RANGE); continue; }
        const int fl = a.flags ? a.flags[r] : 0;
        const int lo = a.aln_start ? a.aln_start[r] : 0, hi = a.aln_end ? a.aln_end[r] : a.L;
        const bool reverse = fl & BAM_FLAG_REVERSE;
        int cyc = 0, din = 0;                                     // clipped bases keep 0 (:24,:43)
        if (i >= lo && i < hi) {
            const int c = reverse ? hi - 1 - i : i - lo;
            cyc = (fl & BAM_FLAG_READ2) ? -(c + 1) : c;
            bool bad = false;
            din = bam_dinuc(a.seq + r * a.L, i, lo, hi, reverse, q, a.minscore, &bad);
            if (bad) atomicOr(a.status, KBBQ_FLAG_BAD_BASE);
        }
        const bool e = a.err[idx] != 0;
        const int c2 = cyc < 0 ? 2 * a.L + cyc : cyc;
        const size_t po = ((size_t)g * NQ + q) * (2 * a.L) + c2;
        atomicAdd(a.pos_total + po, 1ull);
        if (e) atomicAdd(a.pos_errs + po, 1ull);
        if (din >= 0) {
            const size_t d = ((size_t)g * NQ + q) * 16 + din;
            atomicAdd(a.din_total + d, 1ull);
            if (e) atomicAdd(a.din_errs + d, 1ull);
        }
    }
}

struct ApplyBamArgs {
    const uint8_t *seq, *qual;
    const uint16_t *rg;
    const uint8_t *flags;
    uint8_t *out;
    long long N;
    int L, R, minscore, nq, ndin1;
    const long long *meanq, *rgdq, *qdq, *posdq, *dindq;
    int *status;
};

__global__ void apply_bam_kernel(ApplyBamArgs a) {
    const long long total = a.N * a.L;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long r = idx / a.L;
        const int i = (int)(idx - r * a.L);
        const unsigned int q = a.qual[idx];
        if ((int)q < a.minscore) { a.out[idx] = (uint8_t)q; continue; }
        if ((int)q >= a.nq) { atomicOr(a.status, KBBQ_FLAG_QUAL_RANGE); a.out[idx] = (uint8_t)q; continue; }
        const unsigned int g = a.rg ? a.rg[r] : 0;
        if (g >= (unsigned int)a.R) { atomicOr(a.status, KBBQ_FLAG_RG_RANGE); a.out[idx] = (uint8_t)q; continue; }
        const int fl = a.flags ? a.flags[r] : 0;
        const bool reverse = fl & BAM_FLAG_REVERSE;
        const int c = reverse ? a.L - 1 - i : i;                  // kbbq/gatk/applybqsr.py:46-50
        const int c2 = (fl & BAM_FLAG_READ2) ? 2 * a.L - 1 - c : c;
        bool bad = false;
        int din = bam_dinuc(a.seq + r * a.L, i, 0, a.L, reverse, q, a.minscore, &bad);
        if (bad) atomicOr(a.status, KBBQ_FLAG_BAD_BASE);
        if (din < 0) din = a.ndin1 - 1;                           // index -1 gathers the last column
        const size_t gq = (size_t)g * a.nq + q;
        const long long v = a.meanq[g] + a.rgdq[g] + a.qdq[gq] + a.dindq[gq * a.ndin1 + din] +
                            a.posdq[gq * (2 * a.L) + c2];
        a.out[idx] = (uint8_t)v;
    }
}

}  // namespace kbbq

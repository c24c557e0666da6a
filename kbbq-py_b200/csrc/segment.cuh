// segment.cuh -- the segmented batch layout: rows sorted by (read group, read 1 / read 2).
//
// The shared-memory tables of a CTA hold ONE read group (build.cuh), so with several read groups the
// kernels either gather the pairs of a read group from all over the batch (the work list of
// prepare.cuh: one bulk copy per pair and array, 37-53 % of the HBM roofline) or find them next to each
// other.  This file provides the second: a batch whose rows are ordered by key = 2 * rg + second,
//   rows [seg[k], seg[k + 1]) hold the reads of key k, seg[k] a multiple of SEG_ALIGN rows,
// the rows between the last read of a key and seg[k + 1] being padding (quality 0: never tallied, never
// recalibrated; bases 'A').  Every span is then what a one-read-group batch of single-end reads is: contiguous,
// 16-byte aligned, every row with the same read-1 / read-2 flag -- the header-free streaming path of
// the kernels applies to any number of read groups, to single-end / paired mixtures and to mates in
// different read groups alike, and neither rg[] nor second[] is read by the hot kernels.
//
// Who makes it: kbbq_segment_plan + kbbq_segment_rows on the device (a permutation of rows: 2 B moved per
// byte, hidden under the PCIe copies of the host-buffer entry points), or any packer that knows the read
// group of a read before it writes the row.  dest[i] = row of read i in the segmented batch; the recalibrated
// qualities come back through the same index (kbbq_unsegment_rows).  Tables and output bytes do not depend on
// the order of the rows inside a span, so the plan may hand out rows in any order.
#pragma once
#include "common.cuh"

namespace kbbq {

constexpr int SEG_ALIGN = 16;        // rows; 16 rows of any length start 16-byte aligned, and G | 16
constexpr int SEG_THREADS = 256;
constexpr int SEG_SMEM_KEYS = 8192;  // keys countable in shared memory

struct SegPlanArgs {
    const uint16_t *rg;      // may be NULL (all zero)
    const uint8_t *second;   // may be NULL (all zero)
    long long N;
    int R;
    unsigned int *seg;       // [2R + 1] row offsets (out), then [2R] counts
    unsigned int *cursor;    // [2R] scratch
    unsigned int *dest;      // [N] (out)
    int *status;
};

__device__ __forceinline__ unsigned int seg_key(const SegPlanArgs &a, long long i, bool &ok) {
    const unsigned int r = a.rg ? a.rg[i] : 0u;
    ok = r < (unsigned int)a.R;
    return 2u * r + ((a.second && a.second[i]) ? 1u : 0u);
}

// MODE 0: count the rows of every key (into cursor); MODE 1: hand out rows (cursor = next free row per key).
// Per-block counts in shared memory, one global atomic per (block, key), when the keys fit.
template <int MODE>
__global__ void __launch_bounds__(SEG_THREADS) seg_bucket_kernel(SegPlanArgs a, int per_thread) {
    extern __shared__ unsigned int s_cnt[];
    const int nkeys = 2 * a.R;
    const bool use_smem = nkeys <= SEG_SMEM_KEYS;
    if (use_smem) {
        for (int i = threadIdx.x; i < nkeys; i += blockDim.x) s_cnt[i] = 0;
        __syncthreads();
    }
    // a block owns a contiguous run of reads, so that neighbours in the batch stay neighbours in their span
    const long long base = (long long)blockIdx.x * blockDim.x * per_thread;
    constexpr int MAXPT = 8;
    unsigned int key[MAXPT], slot[MAXPT];
#pragma unroll
    for (int j = 0; j < MAXPT; ++j) {
        key[j] = 0xFFFFFFFFu;
        slot[j] = 0;
        const long long i = base + (long long)j * blockDim.x + threadIdx.x;
        if (j < per_thread && i < a.N) {
            bool ok;
            const unsigned int k = seg_key(a, i, ok);
            if (!ok) {
                if (MODE == 0) atomicOr(a.status, KBBQ_FLAG_RG_RANGE);
                continue;
            }
            key[j] = k;
            slot[j] = use_smem ? atomicAdd(&s_cnt[k], 1u) : atomicAdd(&a.cursor[k], 1u);
        }
    }
    if (!use_smem) {
        if (MODE == 1) {
#pragma unroll
            for (int j = 0; j < MAXPT; ++j) {
                const long long i = base + (long long)j * blockDim.x + threadIdx.x;
                if (j < per_thread && i < a.N) a.dest[i] = key[j] != 0xFFFFFFFFu ? slot[j] : 0xFFFFFFFFu;
            }
        }
        return;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nkeys; i += blockDim.x) {
        const unsigned int c = s_cnt[i];
        if (c) s_cnt[i] = atomicAdd(&a.cursor[i], c);  // reserve the block's share
    }
    if (MODE == 0) return;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < MAXPT; ++j) {
        const long long i = base + (long long)j * blockDim.x + threadIdx.x;
        if (j < per_thread && i < a.N) a.dest[i] = key[j] != 0xFFFFFFFFu ? s_cnt[key[j]] + slot[j] : 0xFFFFFFFFu;
    }
}

// seg[k] = first row of key k (each span padded to SEG_ALIGN rows), counts behind them, cursor[k] = seg[k].
__global__ void seg_scan_kernel(SegPlanArgs a) {
    __shared__ unsigned long long s_part[1024];
    const int nkeys = 2 * a.R;
    const int T = blockDim.x;
    const int per = (nkeys + T - 1) / T;
    const int lo = min(nkeys, (int)threadIdx.x * per), hi = min(nkeys, lo + per);
    unsigned long long sum = 0;
    for (int i = lo; i < hi; ++i) sum += ((unsigned long long)a.cursor[i] + SEG_ALIGN - 1) / SEG_ALIGN * SEG_ALIGN;
    s_part[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long run = 0;
        for (int t = 0; t < T; ++t) {
            const unsigned long long v = s_part[t];
            s_part[t] = run;
            run += v;
        }
        a.seg[nkeys] = (unsigned int)run;  // callers keep N + 2R * SEG_ALIGN below 2^32
    }
    __syncthreads();
    unsigned long long run = s_part[threadIdx.x];
    for (int i = lo; i < hi; ++i) {
        const unsigned int c = a.cursor[i];
        a.seg[i] = (unsigned int)run;
        a.seg[nkeys + 1 + i] = c;
        a.cursor[i] = (unsigned int)run;
        run += ((unsigned long long)c + SEG_ALIGN - 1) / SEG_ALIGN * SEG_ALIGN;
    }
}

// Padding rows of every span: quality 0 (below any minscore >= 1: the trash row of the build, left alone by the
// apply), base 'A' in seq and corr.  One block per key.
__global__ void seg_pad_kernel(const unsigned int *seg, int nkeys, int L, uint8_t *seq, uint8_t *qual, uint8_t *corr) {
    const int k = blockIdx.x;
    const unsigned long long lo = ((unsigned long long)seg[k] + seg[nkeys + 1 + k]) * L, hi = (unsigned long long)seg[k + 1] * L;
    for (unsigned long long i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        if (seq) seq[i] = 'A';
        if (qual) qual[i] = 0;
        if (corr) corr[i] = 'A';
    }
}

// Row permutation: SCATTER  dst row dest[i] = src row i   (into the segmented batch)
//                  GATHER   dst row i = src row dest[i]   (recalibrated qualities back in read order)
// One warp per row; lanes cover the aligned 32-bit words of the destination row and assemble each from the
// two aligned source words that hold its bytes (rows of either side start at any byte).  Nothing outside
// [base, base + rows * L) rounded out to whole words is read, nothing outside the destination row written.
template <bool SCATTER>
__global__ void __launch_bounds__(SEG_THREADS) seg_move_rows_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst,
                                                                    const unsigned int *__restrict__ dest, long long N, int L,
                                                                    long long src_rows) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const uintptr_t s_lo = (uintptr_t)src & ~(uintptr_t)3, s_hi = ((uintptr_t)src + (uintptr_t)src_rows * L + 3) & ~(uintptr_t)3;
    for (long long row = warp; row < N; row += nwarps) {
        const unsigned int d = dest[row];
        if (d == 0xFFFFFFFFu) continue;  // a read whose read group was out of range (status flag set by the plan)
        const uint8_t *s = src + (SCATTER ? row : (long long)d) * L;
        uint8_t *t = dst + (SCATTER ? (long long)d : row) * L;
        const int tmis = (int)((uintptr_t)t & 3);
        uint32_t *t0 = reinterpret_cast<uint32_t *>(t - tmis);
        const int nw = (tmis + L + 3) >> 2;
        for (int w = lane; w < nw; w += 32) {
            const int b0 = 4 * w - tmis;  // row byte held by byte 0 of this destination word
            const uintptr_t sp = (uintptr_t)(s + b0);
            const unsigned int smis = (unsigned int)(sp & 3);
            const uintptr_t sa = sp - smis;
            uint32_t x = 0, y = 0;
            if (sa >= s_lo && sa < s_hi) x = *reinterpret_cast<const uint32_t *>(sa);
            if (smis && sa + 4 >= s_lo && sa + 4 < s_hi) y = *reinterpret_cast<const uint32_t *>(sa + 4);
            const uint32_t v = __funnelshift_r(x, y, 8 * smis);
            if (b0 >= 0 && b0 + 4 <= L) {
                t0[w] = v;
            } else {
#pragma unroll
                for (int b = 0; b < 4; ++b)
                    if (b0 + b >= 0 && b0 + b < L) reinterpret_cast<uint8_t *>(t0 + w)[b] = (uint8_t)(v >> (8 * b));
            }
        }
    }
}

// Prepare pass of a segmented batch for the hot kernels: check the span table and express it in groups of G rows.
// A malformed table (not starting at 0, a span not aligned, decreasing, past `rows_bound`) sets KBBQ_FLAG_SEGMENTS
// and leaves an empty walk.
__global__ void seg_prepare_kernel(const unsigned int *seg_rows, int nkeys, int G, unsigned long long rows_bound,
                                   unsigned int *seg_groups, unsigned int *uni, int *status) {
    __shared__ int s_bad;
    if (threadIdx.x == 0) s_bad = 0;
    __syncthreads();
    int bad = 0;
    for (int k = threadIdx.x; k <= nkeys; k += blockDim.x) {
        const unsigned int v = seg_rows[k];
        if (v % SEG_ALIGN) bad = 1;
        if (k == 0 && v != 0) bad = 1;
        if (k > 0 && v < seg_rows[k - 1]) bad = 1;
        if (v > rows_bound) bad = 1;
    }
    if (bad) s_bad = 1;
    __syncthreads();
    for (int k = threadIdx.x; k <= nkeys; k += blockDim.x) seg_groups[k] = s_bad ? 0u : seg_rows[k] / (unsigned int)G;
    if (threadIdx.x == 0) {
        uni[0] = 0u;  // every span is uniform by construction
        uni[1] = uni[2] = 0x01010101u;
        if (s_bad) atomicOr(status, KBBQ_FLAG_SEGMENTS);
    }
}

}  // namespace kbbq

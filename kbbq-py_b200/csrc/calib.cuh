// calib.cuh -- calibration benchmark counts.
//
// Replaces the two np.bincount calls of benchmark.calculate_q (kbbq/benchmark.py:76-91) and the
// errors[~skips] / quals[~skips] selection in front of them (kbbq/benchmark.py:102-104,131-133): per
// base total[q] += 1, errs[q] += error, unless the base is skipped.  A degenerate table build: one
// covariate (the quality), no cycle / dinucleotide index.
//
// Roofline: HBM, 2 B/base (qual + error mask) or 3 B/base (qual + seq + corrected), + 1 with a skip
// mask.  128-bit loads, two vectors in flight per thread; histograms privatised per LANE in shared
// memory ([warp][quality < 64][lane]: bank == lane, so the reduction is conflict free however
// skewed the qualities are), total and errors packed in one u32 counter (1 + 65025 * error), folded
// into the global int64 counts before a total field could overflow.  Qualities >= 64 (legal for
// bincount, absent from real data) go straight to global atomics.
#pragma once
#include "common.cuh"

namespace kbbq {

constexpr int CAL_BINS = 64;
constexpr int CAL_THREADS = 256;
constexpr int CAL_SMEM = (CAL_THREADS / 32) * CAL_BINS * 32 * 4;  // 64 KB
constexpr long long CAL_MAX_ITERS = 2000;  // x 32 bases per thread and iteration: a lane counter's total stays < 65025

struct CalibArgs {
    const uint8_t *qual, *err, *seq, *corr, *skip;
    long long n;
    unsigned long long *total, *errs;
};

__device__ __forceinline__ uint4 ldg_stream16(const uint8_t *p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
// bit 7 of every non-zero byte
__device__ __forceinline__ uint32_t nonzero_bytes(uint32_t x) {
    return (((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & H4;
}

constexpr uint32_t CAL_ERR_UNIT = 65025u;  // 255 * 255: what an error adds on top of the 1 (as in build.cuh)

// Four bases.  Branch free on the common path (every quality < 64): the increment of base b is
// keep_b + 65025 * (error_b & keep_b), two IDP.4A with one-hot byte constants on byte masks, and its
// address q_b * 128 + lane column is a third; a skipped base adds 0.
template <bool HAS_SKIP>
__device__ __forceinline__ void cal_word(uint32_t qw, uint32_t ew, uint32_t kw, uint32_t base, const CalibArgs &a) {
    const uint32_t e = nonzero_bytes(ew);                                    // bit 7 per byte
    const uint32_t keep1 = HAS_SKIP ? ((~nonzero_bytes(kw)) >> 7) & ONE4 : ONE4;   // 1 per kept byte
    uint32_t errff;  // 0xFF per kept error: PTX prmt replicates the sign of byte b for selector nibble 8 | b
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(errff) : "r"(e), "r"(0u), "r"(0xBA98u));
    if (HAS_SKIP) errff &= keep1 * 255u;
    if ((qw & 0xC0C0C0C0u) == 0u) {
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const uint32_t inc = __dp4a(errff, 255u << (8 * b), __dp4a(keep1, 1u << (8 * b), 0u));
            const uint32_t addr = __dp4a(qw, 128u << (8 * b), base);
            asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(inc) : "memory");
        }
        return;
    }
#pragma unroll
    for (int b = 0; b < 4; ++b) {  // a quality >= 64 somewhere in the word: legal for bincount, never seen in real data
        const uint32_t q = (qw >> (8 * b)) & 0xFFu;
        if (!((keep1 >> (8 * b)) & 1u)) continue;
        const uint32_t er = (errff >> (8 * b)) & 1u;
        if (q < CAL_BINS) {
            asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(base + q * 128u), "r"(1u + er * CAL_ERR_UNIT) : "memory");
        } else {
            atomicAdd(a.total + q, 1ull);
            if (er) atomicAdd(a.errs + q, 1ull);
        }
    }
}

__device__ __forceinline__ void cal_fold(unsigned int *hist, const CalibArgs &a) {
    __syncthreads();
    // one warp per (quality) row of every warp's table: sum the 32 lane columns
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = CAL_THREADS / 32;
    for (int q = warp; q < CAL_BINS; q += nwarps) {
        unsigned int tot = 0, er = 0;
        for (int w = 0; w < nwarps; ++w) {
            unsigned int *p = hist + (w * CAL_BINS + q) * 32 + lane;
            const unsigned int v = *p;
            *p = 0;
            tot += v % CAL_ERR_UNIT;
            er += v / CAL_ERR_UNIT;
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            tot += __shfl_xor_sync(0xFFFFFFFFu, tot, o);
            er += __shfl_xor_sync(0xFFFFFFFFu, er, o);
        }
        if (lane == 0) {
            if (tot) atomicAdd(a.total + q, (unsigned long long)tot);
            if (er) atomicAdd(a.errs + q, (unsigned long long)er);
        }
    }
    __syncthreads();
}

template <bool HAS_ERR, bool HAS_SKIP>
__global__ void __launch_bounds__(CAL_THREADS) calibration_kernel(const __grid_constant__ CalibArgs a) {
    extern __shared__ __align__(16) unsigned int cal_hist[];
    for (int i = threadIdx.x; i < CAL_SMEM / 4; i += CAL_THREADS) cal_hist[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(cal_hist) + (uint32_t)(warp * CAL_BINS * 32 + lane) * 4u;

    // the host launches this kernel on 16-byte aligned arrays only and in chunks small enough that no
    // thread runs more than CAL_MAX_ITERS iterations (kbbq_calibration_counts)
    const long long nvec = a.n / 16;
    const long long stride = (long long)gridDim.x * CAL_THREADS;
    long long v = (long long)blockIdx.x * CAL_THREADS + threadIdx.x;
    for (; v < nvec; v += 2 * stride) {
        const long long v2 = v + stride;
        const bool two = v2 < nvec;
        uint4 q0 = ldg_stream16(a.qual + 16 * v), q1 = two ? ldg_stream16(a.qual + 16 * v2) : make_uint4(0, 0, 0, 0);
        uint4 e0, e1, k0 = make_uint4(0, 0, 0, 0), k1 = k0;
        if (HAS_ERR) {
            e0 = ldg_stream16(a.err + 16 * v);
            e1 = two ? ldg_stream16(a.err + 16 * v2) : k0;
        } else {
            const uint4 s0 = ldg_stream16(a.seq + 16 * v), c0 = ldg_stream16(a.corr + 16 * v);
            e0 = make_uint4(s0.x ^ c0.x, s0.y ^ c0.y, s0.z ^ c0.z, s0.w ^ c0.w);
            if (two) {
                const uint4 s1 = ldg_stream16(a.seq + 16 * v2), c1 = ldg_stream16(a.corr + 16 * v2);
                e1 = make_uint4(s1.x ^ c1.x, s1.y ^ c1.y, s1.z ^ c1.z, s1.w ^ c1.w);
            } else e1 = k0;
        }
        if (HAS_SKIP) {
            k0 = ldg_stream16(a.skip + 16 * v);
            if (two) k1 = ldg_stream16(a.skip + 16 * v2);
        }
        cal_word<HAS_SKIP>(q0.x, e0.x, k0.x, base, a);
        cal_word<HAS_SKIP>(q0.y, e0.y, k0.y, base, a);
        cal_word<HAS_SKIP>(q0.z, e0.z, k0.z, base, a);
        cal_word<HAS_SKIP>(q0.w, e0.w, k0.w, base, a);
        if (two) {
            cal_word<HAS_SKIP>(q1.x, e1.x, k1.x, base, a);
            cal_word<HAS_SKIP>(q1.y, e1.y, k1.y, base, a);
            cal_word<HAS_SKIP>(q1.z, e1.z, k1.z, base, a);
            cal_word<HAS_SKIP>(q1.w, e1.w, k1.w, base, a);
        }
    }
    // tail: the last n % 16 bases, one thread each
    if (blockIdx.x == 0 && threadIdx.x < (int)(a.n - nvec * 16)) {
        const long long i = nvec * 16 + threadIdx.x;
        const uint32_t q = a.qual[i];
        const uint32_t er = HAS_ERR ? (a.err[i] != 0) : (a.seq[i] != a.corr[i]);
        if (!(HAS_SKIP && a.skip[i])) {
            atomicAdd(a.total + q, 1ull);
            if (er) atomicAdd(a.errs + q, 1ull);
        }
    }
    cal_fold(cal_hist, a);
}

// Arrays that are not 16-byte aligned: one base per thread, CTA-wide shared-memory histogram.
__global__ void __launch_bounds__(CAL_THREADS) calibration_scalar_kernel(const __grid_constant__ CalibArgs a) {
    __shared__ unsigned int tot[256], er[256];
    tot[threadIdx.x] = 0;
    er[threadIdx.x] = 0;
    __syncthreads();
    for (long long i = (long long)blockIdx.x * CAL_THREADS + threadIdx.x; i < a.n; i += (long long)gridDim.x * CAL_THREADS) {
        if (a.skip && a.skip[i]) continue;
        const uint32_t q = a.qual[i];
        atomicAdd(&tot[q], 1u);
        if (a.err ? (a.err[i] != 0) : (a.seq[i] != a.corr[i])) atomicAdd(&er[q], 1u);
    }
    __syncthreads();
    if (tot[threadIdx.x]) atomicAdd(a.total + threadIdx.x, (unsigned long long)tot[threadIdx.x]);
    if (er[threadIdx.x]) atomicAdd(a.errs + threadIdx.x, (unsigned long long)er[threadIdx.x]);
}

}  // namespace kbbq

// apply.cuh -- K2: recalibration apply kernel.
//
// Replaces compare_reads.recalibrate_fastq (kbbq/compare_reads.py:320-328) over a packed batch:
//   out = q < minscore ? q : meanq[rg] + rgdq[rg] + qdq[rg,q] + dindq[rg,q,dinuc] + posdq[rg,q,cycle]
// Every term is an integer, so the five-term sum is folded (model.cuh: fold_kernel) into two small
// tables per read group -- fold_cyc[q][cycle] (meanq + rgdq + qdq + posdq) and fold_din[q][dinuc]
// -- and the kernel does two shared-memory gathers per base instead of five global ones.
//
// Roofline: HBM, 3 B/base (seq + qual in, new qual out).  Same super-row thread mapping as the
// build kernel (common.cuh): the cycle table is gathered at consecutive banks by consecutive lanes
// and the dinuc table is replicated per lane, so both gathers are bank-conflict free.
#pragma once
#include "build.cuh"
#include "common.cuh"
#include "stage.cuh"

namespace kbbq {

struct ApplyArgs {
    const uint8_t *seq, *qual;
    uint8_t *out;
    Geom g;
    StageLayout sl;
    int R, nq;
    const entry_t *entries;
    const unsigned int *seg;
    const short *fold_cyc;  // [R][43][2L]
    const short *fold_din;  // [R][43][32]
    int *status;
};

template <int DREP>
__global__ void __launch_bounds__(MAX_THREADS, 1) apply_smem_kernel(ApplyArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const Geom &g = a.g;
    const StageLayout &sl = a.sl;
    const int nqv = g.nqv, row = g.row;
    int *t_cyc = reinterpret_cast<int *>(smem_raw);  // [nqv + 1][row]        (+ zero trash row)
    int *t_din = t_cyc + (nqv + 1) * row;            // [nqv + 1][17][DREP]   (+ zero trash row)
    const int nconsumers = g.threads;

    const unsigned long long E = a.seg[a.R];
    const uint32_t lo = (uint32_t)(E * blockIdx.x / gridDim.x), hi = (uint32_t)(E * (blockIdx.x + 1) / gridDim.x);

    const uint32_t bar0 = smem_u32(smem_raw + sl.bar_off);
    if (threadIdx.x == 0) {
        for (int s = 0; s < sl.stages; ++s) {
            mbar_init(bar0 + s * 8, 1);
            mbar_init(bar0 + (sl.stages + s) * 8, nconsumers / 32);
        }
        mbar_fence_init();
    }
    __syncthreads();

    if ((int)threadIdx.x >= nconsumers) {  // ---- producer warp ----
        ProducerArgs p;
        p.arr[0] = a.seq; p.arr[1] = a.qual; p.arr[2] = nullptr;
        p.entries = a.entries; p.seg = a.seg; p.R = a.R; p.lo = lo; p.hi = hi;
        p.gbytes = g.gbytes; p.ng = g.ng;
        producer_loop(p, sl, smem_raw);
        return;
    }

    const ThreadMap m = make_thread_map(g);
    const int lane = threadIdx.x & 31;
    const uint32_t minq4 = (uint32_t)g.minscore * ONE4;
    const uint32_t trash4 = (uint32_t)NQ * ONE4;
    const uint32_t cyc_base = smem_addr(t_cyc) - (uint32_t)g.minscore * row * 4;
    const uint32_t din_base = pin(smem_addr(t_din) + (lane & (DREP - 1)) * 4 - (uint32_t)g.minscore * DIN_SLOTS * DREP * 4);
    const uint32_t row_bytes = row * 4, dq_bytes = DIN_SLOTS * DREP * 4;
    const uint32_t rowbit = m.row >= 0 ? (1u << m.row) : 0u, secbit = rowbit << 8;
    uint32_t afwd[4], arev[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) { afwd[b] = pin(cyc_base + m.fwd[b]); arev[b] = pin(cyc_base + m.rev[b]); }
    const uint32_t data0 = smem_u32(smem_raw + sl.data_off) + m.toff;
    const uint32_t hdr0 = smem_u32(smem_raw + sl.hdr_off) + m.grp * 16;
    const uint32_t stage_bytes = sl.narr * sl.abytes, abytes = sl.abytes, hdr_stride = g.ng * 16;
    const uint32_t prev_keep = m.need_prev ? 0u : 7u;
    const uint32_t nqlim4 = (uint32_t)(127 - (a.nq - 1)) * ONE4;  // q + this has bit 7 set iff q >= nq
    const bool full_word = m.rowmask == 0xFFFFFFFFu;
    uint32_t stage = 0, phase = 0;
    uint32_t qbad = 0;

    for (int rg = 0; rg < a.R; ++rg) {
        uint32_t s_lo = a.seg[rg], s_hi = a.seg[rg + 1];
        if (s_hi <= lo) continue;
        if (s_lo >= hi) break;
        if (s_lo < lo) s_lo = lo;
        if (s_hi > hi) s_hi = hi;

        // stage this read group's folded tables (trash row = 0)
        consumer_sync(nconsumers);
        const int L2 = 2 * g.L;
        const short *fc = a.fold_cyc + ((size_t)rg * NQ + g.minscore) * L2;
        for (int i = threadIdx.x; i < nqv * L2; i += nconsumers) {
            const int q = i / L2, c2 = i - q * L2;
            t_cyc[q * row + plane_pos(c2, g.sj)] = fc[i];
        }
        for (int i = threadIdx.x; i < row; i += nconsumers) t_cyc[nqv * row + i] = 0;
        const short *fd = a.fold_din + ((size_t)rg * NQ + g.minscore) * 32;
        for (int i = threadIdx.x; i < (nqv + 1) * DIN_SLOTS * DREP; i += nconsumers) {
            const int cell = i / DREP, q = cell / DIN_SLOTS, sl_ = cell - q * DIN_SLOTS;
            t_din[i] = q < nqv ? fd[q * 32 + sl_] : 0;
        }
        consumer_sync(nconsumers);

        for (uint32_t first = s_lo; first < s_hi; first += g.ng) {
            mbar_wait(bar0 + stage * 8, phase);
            uint32_t bits, soff, grp, hpad;
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(bits), "=r"(soff), "=r"(grp), "=r"(hpad)
                         : "r"(hdr0 + stage * hdr_stride));
            uint32_t sw = 0, qw = 0, pb = 0;
            const bool mine = (bits & rowbit) != 0;
            if (mine) {
                const uint32_t wa = data0 + stage * stage_bytes + soff;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(sw) : "r"(wa));
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(qw) : "r"(wa + abytes));
                asm volatile("ld.shared.u8 %0, [%1];" : "=r"(pb) : "r"(wa - 1));
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar0 + (sl.stages + stage) * 8);
            if (++stage == (uint32_t)sl.stages) { stage = 0; phase ^= 1; }
            if (!mine) continue;
            const uint32_t am = m.rowmask;
            // global byte offset of this word: group index * group bytes + offset inside the group
            const long long off = (long long)((unsigned long long)grp * g.gbytes) + m.toff;

            const uint32_t code3 = (sw >> 1) & 0x07070707u;
            const uint32_t pv3 = ((pb >> 1) & 7u) | prev_keep;
            const uint32_t pc3 = __byte_perm(pv3, code3, 0x6540);

            const uint32_t bad = ((qw + nqlim4) | qw) & H4 & am;  // q >= nq: IndexError in the reference
            qbad |= bad;
            const uint32_t vm = ((qw | H4) - minq4) & H4 & am & ~bad;
            const uint32_t anyn = ((code3 | pc3) << 5) & H4;
            const uint32_t dm = vm & ~anyn & m.notfirst;
            const uint32_t vm8 = (vm >> 7) * 0xFFu, dm8 = (dm >> 7) * 0xFFu;
            const uint32_t q4 = (qw & vm8) | (trash4 & ~vm8);
            const uint32_t din4 = ((pc3 << 2) & 0x0C0C0C0Cu) | (code3 & 0x03030303u);
            const uint32_t d4 = (din4 & dm8) | (0x10101010u & ~dm8);  // invalid dinuc -> the pad slot
            const bool sec = (bits & secbit) != 0;

            uint32_t v[4];
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const uint32_t qb = __byte_perm(q4, 0, 0x4440 + b);
                const uint32_t db = __byte_perm(d4, 0, 0x4440 + b);
                uint32_t x, y;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(x) : "r"(qb * row_bytes + (sec ? arev[b] : afwd[b])));
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(y) : "r"(qb * dq_bytes + (db * (DREP * 4) + din_base)));
                v[b] = x + y;
            }
            const uint32_t sum4 = __byte_perm(__byte_perm(v[0], v[1], 0x0040), __byte_perm(v[2], v[3], 0x0040), 0x5410);
            const uint32_t res = (sum4 & vm8) | (qw & ~vm8);  // untouched below minscore
            if (full_word) {
                *reinterpret_cast<unsigned int *>(a.out + off) = res;
            } else {
#pragma unroll
                for (int b = 0; b < 4; ++b)
                    if ((am >> (8 * b)) & 1u) a.out[off + b] = (uint8_t)(res >> (8 * b));
            }
        }
    }
    if (qbad) atomicOr(a.status, KBBQ_FLAG_QUAL_RANGE);
}

// Generic path (any L): one thread per base, folded tables gathered from global memory (L1/L2).
struct ApplyGenericArgs {
    const uint8_t *seq, *qual;
    const uint16_t *rg;
    const uint8_t *second;
    uint8_t *out;
    long long N;
    int L, R, minscore, nq;
    const short *fold_cyc, *fold_din;
    int *status;
};

__global__ void apply_generic_kernel(ApplyGenericArgs a) {
    const long long total = a.N * a.L;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / a.L;
        const int c = (int)(i - r * a.L);
        const unsigned int q = a.qual[i];
        if ((int)q < a.minscore) { a.out[i] = (uint8_t)q; continue; }
        if ((int)q >= a.nq) { atomicOr(a.status, KBBQ_FLAG_QUAL_RANGE); a.out[i] = (uint8_t)q; continue; }
        const unsigned int g = a.rg ? a.rg[r] : 0;
        if (g >= (unsigned int)a.R) { atomicOr(a.status, KBBQ_FLAG_RG_RANGE); a.out[i] = (uint8_t)q; continue; }
        const int c2 = (a.second && a.second[r]) ? 2 * a.L - 1 - c : c;
        int slot = 16;
        if (c > 0) {
            const uint8_t s = a.seq[i], ps = a.seq[i - 1];
            if (s != 'N' && ps != 'N') slot = (((ps >> 1) & 3) << 2) | ((s >> 1) & 3);
        }
        const size_t gq = (size_t)g * NQ + q;
        const int v = a.fold_cyc[gq * (2 * a.L) + c2] + a.fold_din[gq * 32 + slot];
        a.out[i] = (uint8_t)v;
    }
}

inline size_t apply_smem_bytes(const Geom &g, int drep) {
    return sizeof(int) * ((size_t)(g.nqv + 1) * g.row + (size_t)(g.nqv + 1) * DIN_SLOTS * drep);
}

}  // namespace kbbq

// build.cuh -- K1: fused mismatch + covariate-index + histogram kernel (table build).
//
// Replaces, per base, find_corrected_sites (kbbq/recalibrate.py:13-20), the cycle and dinucleotide
// covariates (kbbq/compare_reads.py:275-302), the mask block (kbbq/recalibrate.py:96-101) and the
// four pos_* / dinuc_* np.add.at calls (:116-119).
//
// Roofline: HBM.  Algorithmic traffic 3 B/base (seq + qual + corrected), nothing written but the
// tables.  In practice the kernel is bound by instruction issue and shared-memory atomics (two
// per base), so the design minimises instructions per base and keeps every ATOMS conflict free:
//   * one thread owns one (row, 32-bit word) position of a group of G reads (common.cuh), i.e.
//     four fixed cycles of one read, and walks down the groups; lanes of a warp own consecutive
//     words, so at every byte position they hit consecutive banks of the cycle table whatever the
//     qualities are;
//   * the dinucleotide table (37 x 16 cells, every lane could hit the same cell) is replicated
//     per lane ([cell][lane]), so bank == lane;
//   * bases that must not be tallied (q < minscore, bytes of a neighbouring read, invalid dinuc)
//     are steered to a TRASH row / slot by byte-parallel selects instead of per-byte predicates,
//     so the eight reductions of a word are unconditional `red.shared.add.u32`;
//   * reads reach the SM through a TMA bulk-copy ring filled by a producer warp (stage.cuh): HBM
//     is read exactly once and the consumers' hot loop holds no global loads;
//   * mismatches (about 1 % of bases) take a divergent slow path into unreplicated tables;
//   * tables are per-CTA u32 in shared memory, flushed once per read-group segment to the global
//     int64 tables with 64-bit reductions (zero cells skipped).
#pragma once
#include "common.cuh"
#include "prepare.cuh"
#include "stage.cuh"

namespace kbbq {

constexpr int DIN_SLOTS = 17;  // 16 dinucleotides + 1 trash slot

struct BuildArgs {
    const uint8_t *seq, *qual, *corr;
    Geom g;
    StageLayout sl;
    int R;
    const entry_t *entries;
    const unsigned int *seg;  // [R + 1]
    unsigned long long *pos_errs, *pos_total, *din_errs, *din_total;
    int *status;
};

__device__ __forceinline__ void red_shared_inc(uint32_t saddr) {
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(saddr));
}

__device__ __forceinline__ uint32_t smem_addr(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

// Per-thread constants of the (row, word) mapping.
struct ThreadMap {
    int grp;             // thread-group index inside the CTA
    int row;             // row of the group this thread works on (-1: padding lane)
    int toff;            // byte offset of this thread's word inside a group
    uint32_t rowmask;    // 0xFF for the bytes of the word that belong to the row
    uint32_t notfirst;   // H4 bit for bytes whose cycle is not 0
    uint32_t fwd[4];     // forward table position (bytes inside a quality row) of each byte's cycle
    uint32_t rev[4];     // reverse (read-2) position
    bool need_prev;      // byte 0 has a predecessor in the same row (word index > 0)
};

__device__ __forceinline__ ThreadMap make_thread_map(const Geom &g) {
    ThreadMap m;
    const int tid = threadIdx.x;
    m.grp = tid / g.lps;
    const int t = tid - m.grp * g.lps;
    m.row = -1;
    int w = 0;
#pragma unroll
    for (int k = 0; k < MAX_G; ++k)
        if (k < g.G && t >= g.wstart[k] && t < g.wstart[k + 1]) { m.row = k; w = t - g.wstart[k]; }
    m.rowmask = 0; m.notfirst = 0; m.toff = 0; m.need_prev = false;
#pragma unroll
    for (int b = 0; b < 4; ++b) { m.fwd[b] = 0; m.rev[b] = 0; }
    if (m.row >= 0) {
        const int a = (m.row * g.L) & 3;
        m.toff = m.row * g.L - a + 4 * w;
        m.need_prev = w > 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int c = 4 * w + b - a;
            if (c >= 0 && c < g.L) {
                m.rowmask |= 0xFFu << (8 * b);
                if (c != 0) m.notfirst |= 0x80u << (8 * b);
                m.fwd[b] = 4u * plane_pos(c, g.sj);
                m.rev[b] = 4u * plane_pos(2 * g.L - 1 - c, g.sj);
            }
        }
    }
    return m;
}

// keep a loop-invariant value in its register instead of letting ptxas rematerialise it
__device__ __forceinline__ uint32_t pin(uint32_t v) {
    asm volatile("" : "+r"(v));
    return v;
}

template <int DREP, bool VALIDATE>
__global__ void __launch_bounds__(MAX_THREADS, 1) build_smem_kernel(BuildArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned int *smem = reinterpret_cast<unsigned int *>(smem_raw);
    const Geom &g = a.g;
    const StageLayout &sl = a.sl;
    const int nqv = g.nqv, row = g.row;
    unsigned int *pos_t = smem;                              // [nqv + 1][row]           (+ trash row)
    unsigned int *din_t = pos_t + (nqv + 1) * row;           // [nqv + 1][17][DREP]      (+ trash row / slot)
    unsigned int *pos_e = din_t + (nqv + 1) * DIN_SLOTS * DREP;  // [nqv][row]
    unsigned int *din_e = pos_e + nqv * row;                 // [nqv][16]
    const int table_words = (nqv + 1) * row + (nqv + 1) * DIN_SLOTS * DREP + nqv * row + nqv * 16;
    const int nconsumers = g.threads;                        // + one producer warp

    // this CTA's slice of the concatenated work list
    const unsigned long long E = a.seg[a.R];
    const uint32_t lo = (uint32_t)(E * blockIdx.x / gridDim.x), hi = (uint32_t)(E * (blockIdx.x + 1) / gridDim.x);

    const uint32_t bar0 = smem_u32(smem_raw + sl.bar_off);
    if (threadIdx.x == 0) {
        for (int s = 0; s < sl.stages; ++s) {
            mbar_init(bar0 + s * 8, 1);                               // full: the producer's arrive + tx bytes
            mbar_init(bar0 + (sl.stages + s) * 8, nconsumers / 32);   // empty: one arrive per consumer warp
        }
        mbar_fence_init();
    }
    __syncthreads();

    if ((int)threadIdx.x >= nconsumers) {  // ---- producer warp ----
        ProducerArgs p;
        p.arr[0] = a.seq; p.arr[1] = a.qual; p.arr[2] = a.corr;
        p.entries = a.entries; p.seg = a.seg; p.R = a.R; p.lo = lo; p.hi = hi;
        p.gbytes = g.gbytes; p.ng = g.ng;
        producer_loop(p, sl, smem_raw);
        return;
    }

    // ---- consumer warps ----
    const ThreadMap m = make_thread_map(g);
    const int lane = threadIdx.x & 31;
    const uint32_t minq4 = (uint32_t)g.minscore * ONE4;
    const uint32_t trash4 = (uint32_t)NQ * ONE4;  // quality 43 -> row nqv of the shared tables
    // byte addresses in the shared window, with the "- minscore rows" folded in
    const uint32_t pos_base = smem_addr(pos_t) - (uint32_t)g.minscore * row * 4;
    const uint32_t din_base = pin(smem_addr(din_t) + (lane & (DREP - 1)) * 4 - (uint32_t)g.minscore * DIN_SLOTS * DREP * 4);
    const uint32_t row_bytes = row * 4, dq_bytes = DIN_SLOTS * DREP * 4;
    const uint32_t rowbit = m.row >= 0 ? (1u << m.row) : 0u, secbit = rowbit << 8;
    uint32_t afwd[4], arev[4];  // absolute shared addresses of quality row 0 at this thread's cycles
#pragma unroll
    for (int b = 0; b < 4; ++b) { afwd[b] = pin(pos_base + m.fwd[b]); arev[b] = pin(pos_base + m.rev[b]); }
    const uint32_t data0 = smem_u32(smem_raw + sl.data_off) + m.toff;
    const uint32_t hdr0 = smem_u32(smem_raw + sl.hdr_off) + m.grp * 16;
    const uint32_t stage_bytes = sl.narr * sl.abytes, abytes = sl.abytes, hdr_stride = g.ng * 16;
    const uint32_t prev_keep = m.need_prev ? 0u : 7u;  // no predecessor in this row -> treat as N
    uint32_t stage = 0, phase = 0;
    uint32_t qbad = 0, bbad = 0;

    for (int rg = 0; rg < a.R; ++rg) {
        uint32_t s_lo = a.seg[rg], s_hi = a.seg[rg + 1];
        if (s_hi <= lo) continue;
        if (s_lo >= hi) break;
        if (s_lo < lo) s_lo = lo;
        if (s_hi > hi) s_hi = hi;

        for (int i = threadIdx.x; i < table_words; i += nconsumers) smem[i] = 0;
        consumer_sync(nconsumers);

        for (uint32_t first = s_lo; first < s_hi; first += g.ng) {
            mbar_wait(bar0 + stage * 8, phase);
            uint32_t bits, soff;
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(bits), "=r"(soff) : "r"(hdr0 + stage * hdr_stride));
            uint32_t sw = 0, qw = 0, cw = 0, pb = 0;
            const bool mine = (bits & rowbit) != 0;
            if (mine) {
                const uint32_t wa = data0 + stage * stage_bytes + soff;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(sw) : "r"(wa));
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(qw) : "r"(wa + abytes));
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(cw) : "r"(wa + 2 * abytes));
                asm volatile("ld.shared.u8 %0, [%1];" : "=r"(pb) : "r"(wa - 1));
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar0 + (sl.stages + stage) * 8);  // the stage may be refilled
            if (++stage == (uint32_t)sl.stages) { stage = 0; phase ^= 1; }
            if (!mine) continue;  // padding lane, or the row belongs to another read group
            const uint32_t am = m.rowmask;

            // 3-bit base code (b >> 1) & 7: A=0 C=1 T=2 G=3 N=7, injective on ACGTN
            const uint32_t code3 = (sw >> 1) & 0x07070707u;
            const uint32_t pv3 = ((pb >> 1) & 7u) | prev_keep;
            const uint32_t pc3 = __byte_perm(pv3, code3, 0x6540);  // previous base of every byte

            const uint32_t bad = ((qw + 0x55555555u) | qw) & H4 & am;             // q > 42
            qbad |= bad;
            const uint32_t vm = ((qw | H4) - minq4) & H4 & am & ~bad;             // minscore <= q <= 42
            const uint32_t anyn = ((code3 | pc3) << 5) & H4;                      // cur or prev is N
            const uint32_t dm = vm & ~anyn & m.notfirst;                          // dinuc valid
            const uint32_t vm8 = (vm >> 7) * 0xFFu, dm8 = (dm >> 7) * 0xFFu;
            const uint32_t q4 = (qw & vm8) | (trash4 & ~vm8);                     // untallied bytes -> trash row
            const uint32_t din4 = ((pc3 << 2) & 0x0C0C0C0Cu) | (code3 & 0x03030303u);
            const uint32_t d4 = (din4 & dm8) | (0x10101010u & ~dm8);              // invalid dinuc -> trash slot
            if (VALIDATE) {
                // rebuild each byte from its code with an 8-entry byte LUT; any difference = bad base
                const uint32_t y = code3 | (code3 >> 4);
                const uint32_t sel = __byte_perm(y, 0, 0x4420);
                const uint32_t recon = __byte_perm(0x47544341u /* A C T G */, 0x4E000000u /* . . . N */, sel);
                bbad |= (recon ^ sw) & am;
            }
            // read-2 rows count cycles from the end of the axis
            const bool sec = (bits & secbit) != 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const uint32_t qb = __byte_perm(q4, 0, 0x4440 + b);
                const uint32_t db = __byte_perm(d4, 0, 0x4440 + b);
                red_shared_inc(qb * row_bytes + (sec ? arev[b] : afwd[b]));
                red_shared_inc(qb * dq_bytes + (db * (DREP * 4) + din_base));
            }
            // mismatches: rare, divergent
            const uint32_t x = sw ^ cw;
            const uint32_t xm = (((x | H4) - ONE4) | x) & vm;  // byte differs and is tallied
            if (xm) {
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    if (xm & (0x80u << (8 * b))) {
                        const uint32_t qb = (qw >> (8 * b)) & 0xFFu;
                        const uint32_t pos = (sec ? m.rev[b] : m.fwd[b]) >> 2;
                        atomicAdd(&pos_e[(qb - g.minscore) * row + pos], 1u);
                        if (dm & (0x80u << (8 * b)))
                            atomicAdd(&din_e[(qb - g.minscore) * 16 + ((din4 >> (8 * b)) & 0xFu)], 1u);
                    }
                }
            }
        }
        consumer_sync(nconsumers);

        // flush this read group's partial tables: u32 shared -> int64 global
        const int L2 = 2 * g.L;
        unsigned long long *gpe = a.pos_errs + (size_t)rg * NQ * L2, *gpt = a.pos_total + (size_t)rg * NQ * L2;
        for (int i = threadIdx.x; i < nqv * L2; i += nconsumers) {
            const int q = i / L2, c2 = i - q * L2;
            const int s = q * row + plane_pos(c2, g.sj);
            const unsigned int t = pos_t[s], er = pos_e[s];
            const size_t o = (size_t)(q + g.minscore) * L2 + c2;
            if (t) atomicAdd(gpt + o, (unsigned long long)t);
            if (er) atomicAdd(gpe + o, (unsigned long long)er);
        }
        unsigned long long *gde = a.din_errs + (size_t)rg * NQ * 16, *gdt = a.din_total + (size_t)rg * NQ * 16;
        for (int i = threadIdx.x; i < nqv * 16; i += nconsumers) {
            const int q = i >> 4, dn = i & 15;
            unsigned int t = 0;
#pragma unroll 8
            for (int k = 0; k < DREP; ++k)
                t += din_t[(q * DIN_SLOTS + dn) * DREP + ((k + threadIdx.x) & (DREP - 1))];
            const unsigned int er = din_e[i];
            const int dref = nat_to_ref(dn >> 2) * 4 + nat_to_ref(dn & 3);
            const size_t o = (size_t)(q + g.minscore) * 16 + dref;
            if (t) atomicAdd(gdt + o, (unsigned long long)t);
            if (er) atomicAdd(gde + o, (unsigned long long)er);
        }
        consumer_sync(nconsumers);
    }
    if (qbad) atomicOr(a.status, KBBQ_FLAG_QUAL_RANGE);
    if (VALIDATE && bbad) atomicOr(a.status, KBBQ_FLAG_BAD_BASE);
}

// Generic path: one thread per base, global 64-bit reductions.  Used when the shared-memory tables
// do not fit (very long reads) and as an independent on-device cross-check in the tests.
struct BuildGenericArgs {
    const uint8_t *seq, *qual, *corr;
    const uint16_t *rg;
    const uint8_t *second;
    long long N;
    int L, R, minscore;
    unsigned long long *pos_errs, *pos_total, *din_errs, *din_total;
    int *status;
};

__device__ __forceinline__ int base_code_ref(uint8_t b) {  // A T G C -> 0 1 2 3, N -> 4, else 5
    return b == 'A' ? 0 : b == 'T' ? 1 : b == 'G' ? 2 : b == 'C' ? 3 : b == 'N' ? 4 : 5;
}

__global__ void build_generic_kernel(BuildGenericArgs a) {
    const long long total = a.N * a.L;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / a.L;
        const int c = (int)(i - r * a.L);
        const unsigned int q = a.qual[i];
        const uint8_t s = a.seq[i];
        const int code = base_code_ref(s);
        if (code == 5) atomicOr(a.status, KBBQ_FLAG_BAD_BASE);
        if (q > NQ - 1) { atomicOr(a.status, KBBQ_FLAG_QUAL_RANGE); continue; }
        const unsigned int g = a.rg ? a.rg[r] : 0;
        if (g >= (unsigned int)a.R) { atomicOr(a.status, KBBQ_FLAG_RG_RANGE); continue; }
        if ((int)q < a.minscore) continue;
        const bool err = s != a.corr[i];
        const int c2 = (a.second && a.second[r]) ? 2 * a.L - 1 - c : c;
        const size_t po = ((size_t)g * NQ + q) * (2 * a.L) + c2;
        atomicAdd(a.pos_total + po, 1ull);
        if (err) atomicAdd(a.pos_errs + po, 1ull);
        if (c > 0) {
            const int pcode = base_code_ref(a.seq[i - 1]);
            if (code < 4 && pcode < 4) {
                const size_t d = ((size_t)g * NQ + q) * 16 + pcode * 4 + code;
                atomicAdd(a.din_total + d, 1ull);
                if (err) atomicAdd(a.din_errs + d, 1ull);
            }
        }
    }
}

inline size_t build_smem_bytes(const Geom &g, int drep) {
    return sizeof(unsigned int) * ((size_t)(g.nqv + 1) * g.row + (size_t)(g.nqv + 1) * DIN_SLOTS * drep +
                                   (size_t)g.nqv * g.row + (size_t)g.nqv * 16);
}

}  // namespace kbbq

// build.cuh -- K1: fused mismatch + covariate-index + histogram kernel (table build).
//
// Replaces, per base, find_corrected_sites (kbbq/recalibrate.py:13-20), the cycle and dinucleotide
// covariates (kbbq/compare_reads.py:275-302), the mask block (kbbq/recalibrate.py:96-101) and the
// four pos_* / dinuc_* np.add.at calls (:116-119).
//
// Roofline: HBM.  Algorithmic traffic 3 B/base (seq + qual + corrected), nothing written but the
// tables.  The limiter in practice is shared-memory atomic throughput (2 per base), so the design
// goal is conflict-free ATOMS:
//   * one thread owns one 32-bit word position of a super-row (common.cuh), i.e. four fixed
//     cycles, and walks down the reads; lanes of a warp own consecutive words, so at every byte
//     position they hit consecutive banks of the cycle table whatever the qualities are;
//   * the dinucleotide table (37 x 16 cells, every lane could hit the same cell) is replicated
//     per lane ([cell][lane]), so bank == lane;
//   * errors (about 1 % of bases) take a divergent slow path into unreplicated tables;
//   * tables are per-CTA u32 in shared memory, flushed once per read-group segment to the global
//     int64 tables with 64-bit reductions (zero cells skipped).
#pragma once
#include "common.cuh"
#include "prepare.cuh"

namespace kbbq {

struct BuildArgs {
    const uint8_t *seq, *qual, *corr;
    long long total_bytes;  // N * L
    Geom g;
    int R;
    const entry_t *entries;
    const unsigned int *seg;  // [R + 1]
    unsigned long long *pos_errs, *pos_total, *din_errs, *din_total;
    int *status;
};

__device__ __forceinline__ uint32_t ld_word_guarded(const uint8_t *p, long long off, long long total) {
    if (off + 4 <= total) return __ldg(reinterpret_cast<const unsigned int *>(p + off));
    uint32_t v = 0;
    for (int b = 0; b < 4; ++b)
        if (off + b < total) v |= (uint32_t)__ldg(p + off + b) << (8 * b);
    return v;
}

// Per-thread constants of the super-row mapping.
struct ThreadMap {
    int grp, j;
    bool act;
    int rho0;            // row (within the super-row) of byte 0
    uint32_t lo_mask;    // 0xFF for the bytes that lie in row rho0
    uint32_t hi_mask;    // 0xFF for the bytes that lie in row rho0 + 1
    uint32_t notfirst;   // H4 bit for bytes whose cycle is not 0
    uint32_t off[4];     // (forward position) | (reverse position << 16) inside a quality row
    bool first0;
};

__device__ __forceinline__ ThreadMap make_thread_map(const Geom &g) {
    ThreadMap m;
    const int tid = threadIdx.x;
    m.grp = tid / g.lps;
    m.j = tid - m.grp * g.lps;
    m.act = m.j < g.wps;
    const int flat0 = 4 * m.j;
    m.rho0 = flat0 / g.L;
    const int c0 = flat0 - m.rho0 * g.L;
    const int nb0 = min(4, g.L - c0);
    m.lo_mask = 0; m.hi_mask = 0; m.notfirst = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        int c = (b < nb0) ? c0 + b : b - nb0;
        if (b < nb0) m.lo_mask |= 0xFFu << (8 * b); else m.hi_mask |= 0xFFu << (8 * b);
        if (c != 0) m.notfirst |= 0x80u << (8 * b);
        m.off[b] = (uint32_t)plane_pos(c, g.sj) | ((uint32_t)plane_pos(2 * g.L - 1 - c, g.sj) << 16);
    }
    m.first0 = (c0 == 0);
    if (!m.act) { m.lo_mask = 0; m.hi_mask = 0; }
    return m;
}

template <int DREP, bool VALIDATE>
__global__ void __launch_bounds__(MAX_THREADS, 1) build_smem_kernel(BuildArgs a) {
    extern __shared__ unsigned int smem[];
    const Geom &g = a.g;
    const int nqv = g.nqv, row = g.row;
    unsigned int *pos_t = smem;
    unsigned int *pos_e = pos_t + nqv * row;
    unsigned int *din_t = pos_e + nqv * row;
    unsigned int *din_e = din_t + nqv * 16 * DREP;
    const int smem_words = 2 * nqv * row + nqv * 16 * DREP + nqv * 16;

    const ThreadMap m = make_thread_map(g);
    const int lane = threadIdx.x & 31;
    const uint32_t minq4 = (uint32_t)g.minscore * ONE4;
    const int qbase_pos = g.minscore * row;                       // subtract from q*row
    const int din_lane = (lane & (DREP - 1)) - g.minscore * 16 * DREP;

    // this CTA's slice of the concatenated work list
    const unsigned long long E = a.seg[a.R];
    const unsigned long long lo = E * blockIdx.x / gridDim.x, hi = E * (blockIdx.x + 1) / gridDim.x;
    uint32_t qbad = 0, bbad = 0;

    for (int rg = 0; rg < a.R; ++rg) {
        unsigned long long s_lo = a.seg[rg], s_hi = a.seg[rg + 1];
        if (s_hi <= lo) continue;
        if (s_lo >= hi) break;
        if (s_lo < lo) s_lo = lo;
        if (s_hi > hi) s_hi = hi;

        for (int i = threadIdx.x; i < smem_words; i += blockDim.x) smem[i] = 0;
        __syncthreads();

        for (unsigned long long it = s_lo + m.grp; it < s_hi; it += g.ng) {
            const entry_t e = __ldg(a.entries + it);
            const uint32_t sr = (uint32_t)e;
            const uint32_t rowbits = (uint32_t)(e >> 32);
            const uint32_t mA = (rowbits >> m.rho0) & 1u, mB = (rowbits >> (m.rho0 + 1)) & 1u;
            const uint32_t sA = (rowbits >> (4 + m.rho0)) & 1u, sB = (rowbits >> (5 + m.rho0)) & 1u;
            const uint32_t am = (mA ? m.lo_mask : 0u) | (mB ? m.hi_mask : 0u);
            const long long off = (long long)sr * g.srb + 4 * m.j;

            uint32_t sw = 0, qw = 0, cw = 0;
            if (am) {
                sw = ld_word_guarded(a.seq, off, a.total_bytes);
                qw = ld_word_guarded(a.qual, off, a.total_bytes);
                cw = ld_word_guarded(a.corr, off, a.total_bytes);
            }
            // 3-bit base code (b >> 1) & 7: A=0 C=1 T=2 G=3 N=7, injective on ACGTN
            const uint32_t code3 = (sw >> 1) & 0x07070707u;
            uint32_t pv3 = __shfl_up_sync(0xFFFFFFFFu, code3 >> 24, 1);
            if (lane == 0) pv3 = ((am & 0xFFu) && !m.first0) ? ((uint32_t)__ldg(a.seq + off - 1) >> 1) & 7u : 7u;
            const uint32_t pc3 = __byte_perm(pv3, code3, 0x6540);  // previous base of every byte

            const uint32_t bad = ((qw + 0x55555555u) | qw) & H4 & am;             // q > 42
            qbad |= bad;
            const uint32_t vm = ((qw | H4) - minq4) & H4 & am & ~bad;             // minscore <= q <= 42
            const uint32_t anyn = ((code3 | pc3) << 5) & H4;                      // cur or prev is N
            const uint32_t dm = vm & ~anyn & m.notfirst;                          // dinuc valid
            const uint32_t din4 = ((pc3 << 2) & 0x0C0C0C0Cu) | (code3 & 0x03030303u);
            if (VALIDATE) {
                // rebuild each byte from its code with an 8-entry byte LUT; any difference = bad base
                const uint32_t y = code3 | (code3 >> 4);
                const uint32_t sel = __byte_perm(y, 0, 0x4420);
                const uint32_t recon = __byte_perm(0x47544341u /* A C T G */, 0x4E000000u /* . . . N */, sel);
                bbad |= (recon ^ sw) & am;
            }
            // forward / reverse cycle position of every byte (read-2 rows count from the end)
            const uint32_t shA = sA ? 16 : 0, shB = sB ? 16 : 0;

#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const uint32_t qb = (qw >> (8 * b)) & 0xFFu;
                const bool inA = (m.lo_mask >> (8 * b)) & 1u;
                const uint32_t pos = (m.off[b] >> (inA ? shA : shB)) & 0xFFFFu;
                if (vm & (0x80u << (8 * b)))
                    atomicAdd(&pos_t[qb * row + pos - qbase_pos], 1u);
                if (dm & (0x80u << (8 * b))) {
                    const uint32_t db = (din4 >> (8 * b)) & 0xFu;
                    atomicAdd(&din_t[(qb * 16 + db) * DREP + din_lane], 1u);
                }
            }
            // mismatches: rare, divergent
            const uint32_t x = sw ^ cw;
            const uint32_t xm = (((x | H4) - ONE4) | x) & vm;  // byte differs and is tallied
            if (xm) {
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    if (xm & (0x80u << (8 * b))) {
                        const uint32_t qb = (qw >> (8 * b)) & 0xFFu;
                        const bool inA = (m.lo_mask >> (8 * b)) & 1u;
                        const uint32_t pos = (m.off[b] >> (inA ? shA : shB)) & 0xFFFFu;
                        atomicAdd(&pos_e[qb * row + pos - qbase_pos], 1u);
                        if (dm & (0x80u << (8 * b))) {
                            const uint32_t db = (din4 >> (8 * b)) & 0xFu;
                            atomicAdd(&din_e[(qb - g.minscore) * 16 + db], 1u);
                        }
                    }
                }
            }
        }
        __syncthreads();

        // flush this read group's partial tables: u32 shared -> int64 global
        const int L2 = 2 * g.L;
        unsigned long long *gpe = a.pos_errs + (size_t)rg * NQ * L2, *gpt = a.pos_total + (size_t)rg * NQ * L2;
        for (int i = threadIdx.x; i < nqv * L2; i += blockDim.x) {
            const int q = i / L2, c2 = i - q * L2;
            const int s = q * row + plane_pos(c2, g.sj);
            const unsigned int t = pos_t[s], er = pos_e[s];
            const size_t o = (size_t)(q + g.minscore) * L2 + c2;
            if (t) atomicAdd(gpt + o, (unsigned long long)t);
            if (er) atomicAdd(gpe + o, (unsigned long long)er);
        }
        unsigned long long *gde = a.din_errs + (size_t)rg * NQ * 16, *gdt = a.din_total + (size_t)rg * NQ * 16;
        for (int i = threadIdx.x; i < nqv * 16; i += blockDim.x) {
            unsigned int t = 0;
#pragma unroll 8
            for (int k = 0; k < DREP; ++k) t += din_t[i * DREP + ((k + threadIdx.x) & (DREP - 1))];
            const unsigned int er = din_e[i];
            const int q = i >> 4, dn = i & 15;
            const int dref = nat_to_ref(dn >> 2) * 4 + nat_to_ref(dn & 3);
            const size_t o = (size_t)(q + g.minscore) * 16 + dref;
            if (t) atomicAdd(gdt + o, (unsigned long long)t);
            if (er) atomicAdd(gde + o, (unsigned long long)er);
        }
        __syncthreads();
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
    return sizeof(unsigned int) * ((size_t)2 * g.nqv * g.row + (size_t)g.nqv * 16 * drep + (size_t)g.nqv * 16);
}

}  // namespace kbbq

// build.cuh -- K1: fused mismatch + covariate-index + histogram kernel (table build).
//
// Replaces, per base, find_corrected_sites (kbbq/recalibrate.py:13-20), the cycle and dinucleotide
// covariates (kbbq/compare_reads.py:275-302), the mask block (kbbq/recalibrate.py:96-101) and the
// four pos_* / dinuc_* np.add.at calls (:116-119).
//
// Roofline: HBM.  Algorithmic traffic 3 B/base (seq + qual + corrected), nothing written but the
// tables.  What actually limits the kernel is the SM: two shared-memory reductions per base (LSU,
// about 1.6 cycles per warp instruction) and the ALU pipe, on which a LOP3 costs 0.75 - 1 cycle per
// warp against 0.5 for IADD3 / PRMT / SHF and 0.5 on the separate FMA pipe for IMAD / IDP.4A
// (tools/microbench.cu, measured on B200).  The design therefore does all per-base address
// arithmetic on the FMA pipe and keeps the ALU pipe for a handful of byte-parallel ops per word:
//   * one thread owns one (row, 32-bit word) position of a group of G reads (common.cuh), i.e.
//     four fixed cycles of one read; lanes of a warp own consecutive words, so at every byte
//     position they hit consecutive banks of the cycle table whatever the qualities are;
//   * table addresses come from dot products: IDP.2A (16-bit x 8-bit, same rate as IDP.4A) of a
//     16-bit stride constant with the byte-parallel row-index word extracts byte b AND scales it
//     by the row stride in one FMA-pipe instruction; for the dinucleotide table the (slot, row)
//     bytes of a base are interleaved by one PRMT per two bases, so slot * slot stride + row * row
//     stride is ONE instruction per base -- no per-byte extract, select or multiply on the ALU pipe;
//   * bases that must not be tallied (q < minscore, bytes of a neighbouring read, invalid dinuc)
//     get row index 0 (a trash row) through byte masks made by sign-replicating PRMTs whose
//     selectors are per-thread constants (they also encode which bytes of the word the thread owns
//     and which one is cycle 0);
//   * every counter packs total and errors: a base adds 1 + 65025 * mismatch with ONE reduction
//     per table (65025 = 255 * 255, again one IDP.4A), so mismatches cost no branch and no
//     second table; counters are decoded as (v % 65025, v / 65025) and folded before the total
//     field could reach 65025;
//   * the dinucleotide table (every lane could hit the same cell) is replicated per lane
//     ([row][slot][lane]), so bank == lane;
//   * reads reach the SM through a TMA bulk-copy ring filled by a producer warp (stage.cuh): HBM
//     is read exactly once and the consumers' hot loop holds no global loads;
//   * tables are per-CTA u32 in shared memory, flushed per read-group segment to the global int64
//     tables with 64-bit reductions (zero cells skipped).
#pragma once
#include <cstddef>
#include "common.cuh"
#include "prepare.cuh"
#include "stage.cuh"

namespace kbbq {

constexpr uint32_t ERR_UNIT = 65025u;  // 255 * 255: what a mismatching base adds on top of the 1
constexpr int DIN_SLOTS = 16;
constexpr int DIN_REP = 32;            // dinuc table replicas when they fit: bank == lane
constexpr int MAX_WARPS = MAX_THREADS / 32;

// Shared-memory table geometry, chosen on the host (make_table_cfg).
// Row index of a base: qrow = q - (minscore - 1) for minscore <= q <= 42, 0 (trash) otherwise.
struct TableCfg {
    int nrows;        // 44 - minscore (row 0 = trash)
    int rs;           // cycle table: row stride in bytes, multiple of 128
    int sj;           // plane stride in words: cell of cycle c at word (c & 3) * sj + (c >> 2)
    int revoff;       // bytes from the first cycle table (slot 0) to the second (slot 1) = nrows * rs + the bank displacement
    int dq;           // dinuc table: row stride in bytes = 16 slots x drep replicas x 4
    int drep;         // replicas of the dinuc table (32: bank == lane; 16: lanes l and l + 16 share one)
    int pos_off, din_off;            // byte offsets from the start of dynamic shared memory
    int table_bytes;  // zeroed at the start of every segment
    int flush_pos;    // iterations between flushes of the cycle table
    int fold_din;     // iterations between folds of the dinuc replicas
    uint32_t addq;    // 0x81 - minscore in every byte: (q + addq) has bit 7 set iff q >= minscore - 1
    // 16-bit constants of the IDP.2A address arithmetic (dp2a: a.h0 * b.b0 + a.h1 * b.b1 for .lo, bytes 2, 3 for
    // .hi).  Kept in the kernel parameters so that the instruction reads them straight from the constant bank.
    uint32_t cyc16[2];  // {rs, rs << 16}: with the row-index word as b this is (byte 0 or 2) * rs, (byte 1 or 3) * rs
    uint32_t din16;     // (2 * drep) | dq << 16: b = (2 * slot, row) byte pairs -> slot * drep * 4 + row * dq
    uint32_t ohe[4];    // 255 << 8b: IDP.4A with the mismatch byte mask 0xFF -> 65025
    uint32_t val16[2];  // {1 | 16 << 16, 256 | 4096 << 16}: packs the 3-bit codes of four bases into four selector nibbles
};

// Shared-memory wavefronts of one pass of the CTA's consumer warps over the cycle tables in the uniform walk (every
// lane one access per byte position, rows -- i.e. addresses -- all different in the worst case): for every warp and
// byte position the largest number of lanes that meet in one bank.  `shift` = words by which the second table is
// displaced from a whole number of bank rows.  Thread-groups (76 lanes for a pair of 150 bp reads) are packed back to
// back, so a warp holds the tail of one row or group and the head of the next, whose bank ranges overlap; the
// displacement moves the overlap of the row-0 / row-1 transition.  Host side, a few 10 000 operations.
inline int cycle_table_wavefronts(const Geom &g, int sj, int rs_words, int shift, bool word_owner) {
    const int nwords = g.gbytes >> 2;
    int total = 0;
    for (int warp = 0; warp < g.threads / 32; ++warp)
        for (int b = 0; b < 4; ++b) {
            int cnt[32] = {0}, worst = 0;
            for (int lane = 0; lane < 32; ++lane) {
                const int tid = warp * 32 + lane, grp = tid / g.lps, t = tid - grp * g.lps;
                int cell = (lane + 32 * b) % rs_words;   // idle lanes: make_thread_map
                if (grp < g.ng && word_owner) {          // lane t owns word t of the group
                    if (t < nwords) {
                        const int j = 4 * t + b, k = j / g.L, c = j - k * g.L;
                        cell = (c & 3) * sj + (c >> 2) + (k & 1) * shift;
                    }
                } else if (grp < g.ng) {                 // lane t owns one (row, word); foreign bytes go to cell 0
                    for (int k = 0; k < g.G; ++k)
                        if (t >= g.wstart[k] && t < g.wstart[k + 1]) {
                            const int c = 4 * (t - g.wstart[k]) + b - ((k * g.L) & 3);
                            cell = c >= 0 && c < g.L ? (c & 3) * sj + (c >> 2) + (k & 1) * shift : 0;
                        }
                }
                worst = std::max(worst, ++cnt[cell & 31]);
            }
            total += worst;
        }
    return total;
}

// word_owner: which ownership the uniform walk of this kernel uses (make_thread_map): the apply kernel lets a lane own
// a whole word of its group, the build kernel one (row, word)
inline bool make_table_cfg(const Geom &g, int kps, int drep, bool word_owner, TableCfg *t) {
    if (g.minscore < 1) return false;  // row 0 is the trash row
    if (drep != 32 && drep != 16 && drep != 8) return false;
    t->nrows = NQ + 1 - g.minscore;
    t->sj = (g.L + 3) / 4;
    t->rs = (16 * t->sj + 127) / 128 * 128;            // 4 planes x sj words, rounded to whole bank rows
    if (t->rs > 1152) return false;                      // L <= 288: longer reads take the generic kernels
    // the second cycle table starts `shift` words past a whole number of bank rows: the displacement with the fewest
    // modelled bank conflicts (150 bp: 17 words for the apply kernel, 209 wavefronts per pass instead of 240; 15 words
    // for the build kernel, 202 instead of 236; 124 would be conflict free)
    int shift = 0;
    if (const char *e = getenv("KBBQ_SLOT_SHIFT")) shift = std::max(0, std::min(31, atoi(e)));   // tuning hook
    else {
        int best = cycle_table_wavefronts(g, t->sj, t->rs / 4, 0, word_owner);
        for (int d = 1; d < 32; ++d) {
            const int w = cycle_table_wavefronts(g, t->sj, t->rs / 4, d, word_owner);
            if (w < best) { best = w; shift = d; }
        }
    }
    t->revoff = t->nrows * t->rs + 4 * shift;
    t->drep = drep;
    t->dq = DIN_SLOTS * drep * 4;
    t->pos_off = 0;
    t->din_off = (t->revoff + t->nrows * t->rs + 127) / 128 * 128;
    t->table_bytes = t->din_off + t->nrows * t->dq;
    // a cycle cell is hit at most once per (thread-group, row) and iteration; a dinuc replica cell
    // at most 4 times per thread of that lane id and iteration
    t->flush_pos = std::max(1, (int)((ERR_UNIT - 1) / (uint32_t)(kps * g.ng * g.G)));
    t->fold_din = std::max(1, (int)((ERR_UNIT - 1) / (uint32_t)(kps * 4 * ((g.threads + drep - 1) / drep))));
    t->addq = (uint32_t)(0x81 - g.minscore) * ONE4;
    t->cyc16[0] = (uint32_t)t->rs;
    t->cyc16[1] = (uint32_t)t->rs << 16;
    t->din16 = (uint32_t)(drep * 2) | ((uint32_t)t->dq << 16);
    for (int b = 0; b < 4; ++b) t->ohe[b] = 255u << (8 * b);
    t->val16[0] = 1u | (16u << 16);
    t->val16[1] = 256u | (4096u << 16);
    return true;
}

struct BuildArgs {
    const uint8_t *seq, *qual, *corr;
    long long total_bytes;  // N * L
    Geom g;
    TableCfg t;
    StageLayout sl;
    int R;
    int nsub;     // sub-segments of the walk: R list segments, or 2R (read group, read 1 / read 2) spans of a segmented batch
    int segmode;  // 1: segmented batch (segment.cuh): rows sorted by 2 * rg + second, every span contiguous and uniform
    const entry_t *entries;
    const unsigned int *seg;  // [nsub + 1], in groups
    const unsigned int *uni;  // prepare.cuh: {0 = every group has the row flags of group 0, flags 0-3, flags 4-7}
    unsigned long long *pos_errs, *pos_total, *din_errs, *din_total;
    int *status;
};

// ohe[0..3] must start a 16-byte unit of the parameter bank (one LDCU.128 per word instead of four LDC:
// see Geom::pad_ in common.cuh) -- a field added in front of them must keep that
static_assert((offsetof(BuildArgs, t) + offsetof(TableCfg, ohe)) % 16 == 0,
              "BuildArgs: TableCfg::ohe has to be 16-byte aligned in the kernel parameters; pad Geom");

__device__ __forceinline__ void red_shared_add(uint32_t saddr, uint32_t v) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(saddr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t smem_addr(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
// byte permute with the sign-replicate selector bit honoured (PTX prmt.b32, default mode)
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}
// keep a loop-invariant value in its register instead of letting ptxas rematerialise it
__device__ __forceinline__ uint32_t pin(uint32_t v) {
    asm volatile("" : "+r"(v));
    return v;
}

// a + c on the FMA pipe: a * one + c with `one` a register ptxas cannot see through (pin(1u)); ptxas
// would emit IADD3, which runs on the ALU pipe, the busiest unit of the hot kernels (66 % against 27 %
// for the FMA pipe).  Measured: the apply gains 2.5 % from its two byte-parallel adds, the build nothing
// from its three, so only the apply uses it.
__device__ __forceinline__ uint32_t add_fma(uint32_t a, uint32_t one, uint32_t c) {
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(one), "r"(c));
    return d;
}

// Per-thread constants of the (row, word) mapping.
struct ThreadMap {
    int grp;             // thread-group index inside the CTA
    int row;             // row of the group this thread works on (-1: padding lane)
    int toff;            // byte offset of this thread's word inside a group
    uint32_t rowmask;    // 0xFF for the bytes of the word that belong to the row
    uint32_t selv;       // prmt selector: owned byte b -> sign of byte b of the first source, else byte 0 of the second
    uint32_t seln;       // same, but cycle 0 (no previous base) also takes the second source
    uint32_t cell[4];    // byte offset of each byte's cycle inside a row of the cycle table
    int cyc[4];          // cycle of each byte (-1: not owned)
};

// by_parity: the uniform walk -- the table slot of a row is its parity inside the group (interleaved pairs: read 1 /
// read 2; rows that are all the same mate: two replicas of one table, which halves the collisions of a warp that
// holds the end of one row and the start of the next); the cell offsets then include the slot's table and never
// move.  Otherwise the caller moves the addresses between the two tables as the row flags of the work list say.
// (A layout of the second table with its planes shifted so that every warp touches 32 different banks -- 168
// instead of 220 wavefronts per CTA and byte position for 150 bp -- was built and measured on B200: no
// faster, and its larger rows cost the apply kernel a group per stage.  The cycle-table reductions are not what
// the kernel waits for.)
__device__ __forceinline__ ThreadMap make_thread_map(const Geom &g, const TableCfg &tc, bool by_parity, bool word_owner) {
    ThreadMap m;
    const int tid = threadIdx.x;
    m.grp = tid / g.lps;
    const int t = tid - m.grp * g.lps;
    m.row = -1;
    int w = 0;
#pragma unroll
    for (int k = 0; k < MAX_G; ++k)
        if (k < g.G && t >= g.wstart[k] && t < g.wstart[k + 1]) { m.row = k; w = t - g.wstart[k]; }
    if (m.grp >= g.ng) m.row = -1;  // idle lanes behind the last thread-group
    m.rowmask = 0; m.toff = 0; m.selv = 0x4444u; m.seln = 0x4444u;
#pragma unroll
    for (int b = 0; b < 4; ++b) { m.cell[b] = 0; m.cyc[b] = -1; }
    if (by_parity && word_owner) {
        // Uniform walk: every row of a group is tallied and the table of a byte follows from the parity of its row, so
        // nothing ties a thread to ONE row: lane t of the thread-group owns all four bytes of word t of the group,
        // whichever rows they belong to (a word that straddles two reads has one owner instead of two half-owners; the
        // lanes behind the last word idle).  Every live thread then writes whole aligned words in the apply kernel.
        const int nwords = g.gbytes >> 2;
        m.row = (t < nwords && m.grp < g.ng) ? (4 * t) / g.L : -1;
        if (m.row >= 0) {
            m.toff = 4 * t;
            m.rowmask = 0xFFFFFFFFu;
            m.selv = 0xBA98u;
            m.seln = 0xBA98u;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int j = 4 * t + b, k = j / g.L, c = j - k * g.L;
                m.cyc[b] = c;
                if (c == 0) m.seln = (m.seln & ~(0xFu << (4 * b))) | (0x4u << (4 * b));
                m.cell[b] = 4u * (uint32_t)((c & 3) * tc.sj + (c >> 2)) + (uint32_t)((k & 1) * tc.revoff);
            }
        }
    } else if (m.row >= 0) {
        const int a = (m.row * g.L) & 3;
        m.toff = m.row * g.L - a + 4 * w;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int c = 4 * w + b - a;
            if (c >= 0 && c < g.L) {
                m.cyc[b] = c;
                m.rowmask |= 0xFFu << (8 * b);
                m.selv = (m.selv & ~(0xFu << (4 * b))) | ((8u | b) << (4 * b));
                if (c != 0) m.seln = (m.seln & ~(0xFu << (4 * b))) | ((8u | b) << (4 * b));
                m.cell[b] = 4u * (uint32_t)((c & 3) * tc.sj + (c >> 2)) + (by_parity ? (uint32_t)((m.row & 1) * tc.revoff) : 0u);
            }
        }
    }
    if (m.row < 0) {
        // Idle lanes (the tail of the last warp) run along in full stages with nothing selected and tally into the
        // trash row: spread them over its banks instead of letting them collide on one cell.
#pragma unroll
        for (int b = 0; b < 4; ++b) m.cell[b] = 4u * (uint32_t)(((tid & 31) + 32 * b) % (tc.rs / 4));
    }
    if (m.grp >= g.ng) { m.grp = 0; }
    return m;
}

// Fold the per-lane dinuc replicas of this CTA into the global int64 tables of read group rg (one
// warp per cell) and clear them.
__device__ __forceinline__ void fold_din_replicas(const BuildArgs &a, unsigned char *smem_raw, int rg, int nconsumers) {
    const TableCfg &t = a.t;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = nconsumers >> 5;
    unsigned long long *gde = a.din_errs + (size_t)rg * NQ * 16, *gdt = a.din_total + (size_t)rg * NQ * 16;
    for (int cell = DIN_SLOTS + warp; cell < t.nrows * DIN_SLOTS; cell += nwarps) {  // row 0 is the trash row
        const int r = cell / DIN_SLOTS, dn = cell - r * DIN_SLOTS;
        unsigned int *p = reinterpret_cast<unsigned int *>(smem_raw + t.din_off + r * t.dq + dn * (t.drep * 4)) + lane;
        const unsigned int v = lane < t.drep ? *p : 0u;
        if (__any_sync(0xFFFFFFFFu, v != 0)) {
            if (lane < t.drep) *p = 0;
            unsigned int tot = v % ERR_UNIT, er = v / ERR_UNIT;
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                tot += __shfl_xor_sync(0xFFFFFFFFu, tot, o);
                er += __shfl_xor_sync(0xFFFFFFFFu, er, o);
            }
            if (lane == 0) {
                // slot = 4 * code(prev) + code(cur) in natural order; the reference orders A T G C
                const int dref = nat_to_ref(dn >> 2) * 4 + nat_to_ref(dn & 3);
                const size_t o = (size_t)(r + a.g.minscore - 1) * 16 + dref;
                atomicAdd(gdt + o, (unsigned long long)tot);
                if (er) atomicAdd(gde + o, (unsigned long long)er);
            }
        }
    }
}

// Flush both cycle tables of this CTA into the global int64 tables of read group rg and clear them.  hs[s] says which
// half of the cycle axis slot s holds: read 1 (0) or read 2 (1); both slots may hold the same half (single-end reads,
// a span of a segmented batch), the global reductions simply add up.
__device__ __forceinline__ void flush_pos_table(const BuildArgs &a, unsigned char *smem_raw, int rg, int nconsumers, int hs0, int hs1) {
    const Geom &g = a.g;
    const TableCfg &t = a.t;
    const int L = g.L, L2 = 2 * g.L;
    unsigned long long *gpe = a.pos_errs + (size_t)rg * NQ * L2, *gpt = a.pos_total + (size_t)rg * NQ * L2;
    const int per_slot = (t.nrows - 1) * L;
    for (int i = threadIdx.x; i < 2 * per_slot; i += nconsumers) {
        const int slot = i >= per_slot, j = i - slot * per_slot;
        const int r = j / L + 1, c = j - (r - 1) * L;
        unsigned int *p = reinterpret_cast<unsigned int *>(smem_raw + t.pos_off + slot * t.revoff + r * t.rs) +
                          ((c & 3) * t.sj + (c >> 2));
        const unsigned int v = *p;
        if (v) {
            *p = 0;
            const unsigned int tot = v % ERR_UNIT, er = v / ERR_UNIT;
            // row r holds quality r + minscore - 1; read-2 cycles count from the end of the axis
            const int half = slot ? hs1 : hs0;
            const size_t o = (size_t)(r + g.minscore - 1) * L2 + (half ? L2 - 1 - c : c);
            atomicAdd(gpt + o, (unsigned long long)tot);
            if (er) atomicAdd(gpe + o, (unsigned long long)er);
        }
    }
}

// Consumer side of the build kernel.  UNI: one read group and every group of the batch has the row
// flags of group 0 (interleaved pairs, single-end reads: the usual case, detected by the prepare
// pass), so nothing has to be read from the stage headers: the place of a group in the stage and its
// index follow from the iteration, and the row flag is a per-thread constant.
template <int KPS, bool VALIDATE, bool UNI>
__device__ __forceinline__ void build_consume(const BuildArgs &a, unsigned char *smem_raw, uint32_t lo, uint32_t hi,
                                              uint32_t uni_flo, uint32_t uni_fhi) {
    const Geom &g = a.g;
    const TableCfg &t = a.t;
    const StageLayout &sl = a.sl;
    const int nconsumers = g.threads;
    const uint32_t bar0 = pin(smem_u32(smem_raw + sl.bar_off));
    // one (row, word) per lane in either walk: the build has no stores to simplify, and its reductions collide less
    // this way (ncu, 10 M x 150 bp: 51.7 M conflict wavefronts against 72 M with whole-word owners)
    const ThreadMap m = make_thread_map(g, t, UNI, false);
    const int lane = threadIdx.x & 31;
    const uint32_t pos_base = smem_addr(smem_raw + t.pos_off);
    const uint32_t din_base = pin(smem_addr(smem_raw + t.din_off) + (lane & (t.drep - 1)) * 4);
    // shared address of row 0 of this thread's cycles: fixed (its slot's table) in the uniform walk, moved between the
    // read-1 (cur_flag 1) and read-2 (3) table by the work-list walk
    uint32_t aeff[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) aeff[b] = pos_base + m.cell[b];
    uint32_t cur_flag = 1;
    const uint32_t selv = pin(m.selv), seln = pin(m.seln);
    const uint32_t rowsel = m.row >= 0 ? (uint32_t)m.row : 0u;  // prmt selector: flag byte of this thread's row
    const uint32_t lanemask = pin(m.row >= 0 ? 0xFFu : 0u);     // padding lanes never see a live row
    const uint32_t rowmask = pin(m.rowmask);
    const bool live = m.row >= 0;   // idle lanes skip partial stages (in full ones they run along, see make_thread_map)
    // UNI: group j of a stage starts at (misalignment of the stage's first group) + j * gbytes
    const uint32_t data0 = pin(smem_u32(smem_raw + sl.data_off) + m.toff + (UNI ? m.grp * g.gbytes : 0));
    const uint32_t hdr0 = pin(smem_u32(smem_raw + sl.hdr_off) + m.grp * 16);
    const uint32_t kgrp = g.ng * g.gbytes;
    const uint32_t stage_bytes = pin(sl.narr * sl.abytes), abytes = pin(sl.abytes), abytes2 = pin(2 * sl.abytes), hdr_stride = sl.ngs * 16, krec = g.ng * 16;
    const uint32_t nstages = pin(sl.stages), ngs = pin(sl.ngs);
    const uint32_t revoff = t.revoff, addq = t.addq;
    // base validation: the low three bits tell A (1), C (3), T (4), N (6), G (7) apart; 0, 2, 5 map to a byte that no
    // byte with those low bits can equal
    const uint32_t one = pin(1u), minus1 = pin(0xFFFFFFFFu), lut_lo = pin(0x43014101u), lut_hi = pin(0x474E0154u);
    uint32_t stage = 0, phase = 0;
    uint32_t qgood = 0xFFFFFFFFu, bbad = 0;

    for (int i = threadIdx.x; i < t.table_bytes / 4; i += nconsumers) reinterpret_cast<unsigned int *>(smem_raw)[i] = 0;
    consumer_sync(nconsumers);
    // Counters are packed u32 (total + ERR_UNIT * errors): the cycle table is flushed at the latest every flush_pos
    // iterations (stages) and the dinuc replicas folded every fold_din, so that no total field can reach ERR_UNIT.
    // Both flush and fold leave the tables zero (the trash row is never read), so they double as the reset between
    // read groups.
    int cur_rg = -1, cur_sub = -1;
    int hs0 = 0, hs1 = 1;   // half of the cycle axis each table slot holds
    uint32_t since_pos = 0, since_din = 0;
    auto flush_all = [&](int rg) {
        consumer_sync(nconsumers);
        flush_pos_table(a, smem_raw, rg, nconsumers, hs0, hs1);
        fold_din_replicas(a, smem_raw, rg, nconsumers);
        consumer_sync(nconsumers);
        since_pos = since_din = 0;
    };
    if (UNI && !a.segmode) {   // one read group: slot s holds what the flag of row s says (rows of equal parity agree)
        hs0 = (int)((uni_flo & 0xFFu) >> 1);
        hs1 = (int)(((uni_flo >> 8) & 0xFFu) >> 1);
    }

    for (int sub = 0; sub < a.nsub; ++sub) {
        uint32_t s_lo = a.seg[sub], s_hi = a.seg[sub + 1];
        if (s_hi <= lo) continue;
        if (s_lo >= hi) break;
        if (s_lo < lo) s_lo = lo;
        if (s_hi > hi) s_hi = hi;
        if (s_lo >= s_hi) continue;
        // a segmented batch lists (read group, read 1) and (read group, read 2) spans in turn; inside a span both
        // table slots count the same half of the cycle axis, so the tables are flushed whenever the span changes
        const int rg = a.segmode ? sub >> 1 : sub;
        if (cur_rg >= 0 && (rg != cur_rg || (a.segmode && sub != cur_sub))) flush_all(cur_rg);
        cur_rg = rg;
        cur_sub = sub;
        if (a.segmode) hs0 = hs1 = sub & 1;

        for (uint32_t first = s_lo; first < s_hi;) {
            if (since_pos >= (uint32_t)t.flush_pos || since_din >= (uint32_t)t.fold_din) {
                consumer_sync(nconsumers);
                if (since_pos >= (uint32_t)t.flush_pos) { flush_pos_table(a, smem_raw, rg, nconsumers, hs0, hs1); since_pos = 0; }
                if (since_din >= (uint32_t)t.fold_din) { fold_din_replicas(a, smem_raw, rg, nconsumers); since_din = 0; }
                consumer_sync(nconsumers);
            }
            const uint32_t room = min((uint32_t)t.flush_pos - since_pos, (uint32_t)t.fold_din - since_din);
            const uint32_t chunk_end = (uint32_t)min((unsigned long long)s_hi, (unsigned long long)first + (unsigned long long)room * sl.ngs);
            const uint32_t iters = (chunk_end - first + ngs - 1) / ngs;
            since_pos += iters;
            since_din += iters;
            for (; first < chunk_end; first += ngs) {
                mbar_wait(bar0 + stage * 8, phase);
                const uint32_t shdr = pin(hdr0 + stage * hdr_stride);
                uint32_t sdata = data0 + stage * stage_bytes;
                uint32_t nlive = sl.ngs;
                if (UNI) {
                    sdata += (uint32_t)((unsigned long long)first * g.gbytes) & 15u;
                    nlive = s_hi - first;  // groups in this stage (>= ngs except in the last one)
                }
                sdata = pin(sdata);
                // one (row, word) of one group: four bases of one read
                auto tally_word = [&](const uint32_t wa) {
                    uint32_t sw, qw, cw, pb;
                    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(sw) : "r"(wa));
                    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(qw) : "r"(wa + abytes));
                    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(cw) : "r"(wa + abytes2));
                    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(pb) : "r"(wa - 1));

                    // ---- quality -> row index, byte-parallel ----
                    // bit 7 clear <=> q >= 43 (for q < 128): ~(qw + 0x55555555) = qw * -1 - 0x55555556, one IMAD on the FMA
                    // pipe (the ALU pipe is the one the kernel stalls on)
                    const uint32_t nu = add_fma(qw, minus1, 0xAAAAAAAAu);
                    const uint32_t w5 = qw + addq;              // bit 7 <=> q >= minscore - 1, low bits q - (minscore - 1)
                    qgood &= nu & ~qw;                          // bit 7 stays set while every quality is <= 42
                    const uint32_t vraw = w5 & nu & ~qw;        // bit 7 <=> minscore - 1 <= q <= 42
                    const uint32_t vm8 = prmt(vraw, 0u, selv);  // 0xFF for owned bytes in range
                    const uint32_t qrow4 = w5 & vm8 & 0x3F3F3F3Fu;  // 0 = trash row

                    // ---- dinucleotide slot, byte-parallel: 2 * (4 * code(prev) + code(cur)), code = (b >> 1) & 3 ----
                    const uint32_t pw = prmt(pb, sw, 0x6540u);  // previous base of every byte
                    const uint32_t d4 = (pw & 0x06060606u) * 4u + (sw & 0x06060606u);
                    // 'N' is the only accepted base with bit 3 set; cycle 0 and foreign bytes take 0xFF from the selector
                    const uint32_t nm8 = prmt((sw | pw) * 16u, 0xFFFFFFFFu, seln);
                    // (2 * slot, dinuc row) byte pairs of bases 0, 1 and 2, 3: one IDP.2A per base gives the address
                    const uint32_t qd = qrow4 & ~nm8;
                    const uint32_t x01 = prmt(d4, qd, 0x5140u), x23 = prmt(d4, qd, 0x7362u);

                    // ---- mismatch: 0xFF where the corrected base differs (bases are 7-bit) ----
                    const uint32_t e8 = prmt(add_fma(sw ^ cw, one, 0x7F7F7F7Fu), 0u, 0xBA98u);

                    if (VALIDATE) {
                        // rebuild each byte from its 3-bit code (bits 1-3) with an 8-entry byte LUT; any difference = bad
                        // base.  The four codes are gathered into the selector nibbles by two 16-bit x 8-bit dot products
                        // (weights 1, 16 and 256, 4096) on the FMA pipe instead of shift / or / permute on the ALU pipe.
                        const uint32_t code3 = sw & 0x07070707u;
                        const uint32_t sel = __dp2a_hi(t.val16[1], code3, __dp2a_lo(t.val16[0], code3, 0u));
                        const uint32_t recon = prmt(lut_lo, lut_hi, sel);  // raw PRMT: __byte_perm would mask the selector
                        bbad |= recon ^ sw;  // foreign bytes are masked off at the end (ownership is per byte position)
                    }

#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const uint32_t inc = __dp4a(e8, t.ohe[b], one);
                        const uint32_t pa = (b & 2) ? __dp2a_hi(t.cyc16[b & 1], qrow4, aeff[b]) : __dp2a_lo(t.cyc16[b & 1], qrow4, aeff[b]);
                        const uint32_t xb = (b & 2) ? x23 : x01;
                        const uint32_t da = (b & 1) ? __dp2a_hi(t.din16, xb, din_base) : __dp2a_lo(t.din16, xb, din_base);
                        red_shared_add(pa, inc);
                        red_shared_add(da, inc);
                    }
                };
                if (UNI && nlive >= sl.ngs) {
                    // a full stage of a uniform batch, the usual case: no per-word predicates, no divergence.  The
                    // padding lanes of the last warp run along on group 0 with nothing selected: trash rows only.
#pragma unroll
                    for (int k = 0; k < KPS; ++k) tally_word(sdata + k * kgrp);
                } else {
#pragma unroll
                    for (int k = 0; k < KPS; ++k) {
                        uint32_t soff;
                        if (UNI) {
                            if (!live || (uint32_t)(m.grp + k * g.ng) >= nlive) continue;
                            soff = k * kgrp;
                        } else {
                            // this thread-group's k-th record of the stage
                            uint32_t hgrp, flo, fhi;
                            if (g.contig) {   // the identity list, read where it lies; past the span: no rows
                                const uint32_t e = first + (uint32_t)m.grp + (uint32_t)(k * g.ng);
                                const uint4 rec = e < s_hi ? __ldg(a.entries + e) : make_uint4(0u, 0u, 0u, 0u);
                                soff = rec.x; hgrp = rec.y; flo = rec.z; fhi = rec.w;
                            } else {
                                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                                             : "=r"(soff), "=r"(hgrp), "=r"(flo), "=r"(fhi)
                                             : "r"(shdr + k * krec));
                            }
                            // flag byte of this thread's row: 0 = not in this segment (or a padding lane), 1 = read 1,
                            // 3 = read 2.  A row usually keeps its flag from group to group, so the cycle-table
                            // addresses are kept ready for the last flag seen and only re-based when it changes.
                            const uint32_t flag = prmt(flo, fhi, rowsel) & lanemask;
                            if (flag != cur_flag) {
                                if (!flag) continue;
                                const uint32_t delta = ((flag >> 1) - (cur_flag >> 1)) * revoff;
#pragma unroll
                                for (int b = 0; b < 4; ++b) aeff[b] += delta;
                                cur_flag = flag;
                            }
                        }
                        tally_word(sdata + soff);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(bar0 + (nstages + stage) * 8);  // the stage may be refilled
                if (++stage == nstages) { stage = 0; phase ^= 1; }
            }
        }
    }
    if (cur_rg >= 0) flush_all(cur_rg);
    if (~qgood & rowmask & H4) atomicOr(a.status, KBBQ_FLAG_QUAL_RANGE);
    if (VALIDATE && (bbad & rowmask)) atomicOr(a.status, KBBQ_FLAG_BAD_BASE);
}


template <int KPS, bool VALIDATE>
__global__ void __launch_bounds__(MAX_THREADS, 1) build_smem_kernel(const __grid_constant__ BuildArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const Geom &g = a.g;
    const StageLayout &sl = a.sl;
    const int nconsumers = g.threads;  // + one producer warp

    // this CTA's slice of the concatenated work list
    const unsigned long long E = a.seg[a.nsub];
    const uint32_t lo = (uint32_t)(E * blockIdx.x / gridDim.x), hi = (uint32_t)(E * (blockIdx.x + 1) / gridDim.x);
    // one read group (or a segmented batch) whose groups all look alike: contiguous spans, no work list
    const bool contig = a.R == 1 || a.segmode;
    const bool uniform = contig && a.uni[0] == 0u;

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
        p.entries = a.entries; p.seg = a.seg; p.nsub = a.nsub; p.contig = contig; p.lo = lo; p.hi = hi;
        p.gbytes = g.gbytes; p.ng = sl.ngs; p.total_bytes = a.total_bytes;
        p.pw = ((int)threadIdx.x - nconsumers) >> 5; p.nprod = g.nprod;
        p.uniform = uniform;
        producer_loop(p, sl, smem_raw);
        return;
    }

    // ---- consumer warps ----
    if (uniform) build_consume<KPS, VALIDATE, true>(a, smem_raw, lo, hi, a.uni[1], a.uni[2]);
    else build_consume<KPS, VALIDATE, false>(a, smem_raw, lo, hi, 0u, 0u);
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

}  // namespace kbbq

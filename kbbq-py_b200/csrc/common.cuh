// common.cuh -- shared definitions of the B200 kbbq hot-path kernels (sm_100a only).
#pragma once
#include <algorithm>
#include <atomic>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <stdio.h>

#include "../../include/kbbq_b200.h"
#include "../../include/kbbq_model_constants.h"

#ifndef KBBQ_SM_COUNT_FALLBACK
#define KBBQ_SM_COUNT_FALLBACK 148  // B200: 2 dies x 74 SMs
#endif

namespace kbbq {

constexpr int NQ = KBBQ_NQ;            // 43 quality rows (0..42)
#ifndef KBBQ_CTA_THREADS
#define KBBQ_CTA_THREADS 1024
#endif
constexpr int MAX_THREADS = KBBQ_CTA_THREADS;  // consumer + producer threads of a hot-kernel CTA (64 registers per thread at 1024)
constexpr uint32_t H4 = 0x80808080u;   // high bit of every byte
constexpr uint32_t ONE4 = 0x01010101u;

// Launch counter (bench.py reports it as gpu_launches).
extern std::atomic<long long> g_launches;
extern char g_last_cuda_error[256];

#define KBBQ_CUDA(call)                                                                  \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess) {                                                         \
            snprintf(kbbq::g_last_cuda_error, sizeof(kbbq::g_last_cuda_error), "%s:%d %s: %s", \
                     __FILE__, __LINE__, #call, cudaGetErrorString(e_));                 \
            return KBBQ_E_CUDA;                                                          \
        }                                                                                \
    } while (0)

#define KBBQ_LAUNCHED()                  \
    do {                                 \
        ++kbbq::g_launches;              \
        KBBQ_CUDA(cudaGetLastError());   \
    } while (0)

// ---------------------------------------------------------------------------------------------
// Geometry of the thread mapping shared by the build and apply kernels.
//
// Reads are u8 rows of length L packed back to back, so row r starts at byte r*L, which is in
// general only 1- or 2-byte aligned.  A GROUP is G consecutive reads whose total byte length is a
// multiple of 4 (G is a multiple of 4 / gcd(L, 4)), so every group starts 4-byte aligned.  Inside a
// group, row k starts at misalignment a_k = (k*L) & 3 and is covered by W_k = ceil((a_k + L)/4)
// aligned 32-bit words; ONE THREAD OWNS ONE (row, word) of every group its thread-group processes,
// so the cycles of its 4 bytes -- and therefore its shared-memory table offsets -- are fixed for
// the whole kernel, and all its bytes belong to one read (one read group, one `second` flag).
// Words that straddle two rows are loaded by both neighbours; each uses only its own bytes.  (The apply kernel's
// uniform walk lets a lane own the whole word instead -- the table of a byte is then the parity of its row, a
// per-byte constant as well -- so that every lane stores whole words: make_thread_map, build.cuh.)
// A thread-group is the sum W_k lanes of one group; thread-groups are packed back to back over the
// CTA's consumer threads (a warp may hold lanes of two thread-groups -- everything a lane needs is a
// per-thread constant), so only the last warp has padding lanes.  G (<= 8) is the smallest group
// with the fewest straddling words (L = 150: G = 4, 152 lanes for 150 words; 6 thread-groups = 912
// lanes in 29 warps).  A segmented batch (segment.cuh) takes the same geometry with G a power of two.
//
// Shared-memory cycle tables are laid out [q - minscore][plane = c2 & 3][c2 >> 2] with a plane
// stride SJ that makes the row stride a multiple of 32 words: lanes that own consecutive words of
// a row touch consecutive banks at every byte position, whatever their qualities.
// ---------------------------------------------------------------------------------------------
constexpr int MAX_G = 8;

struct Geom {
    int L;            // read length
    int G;            // reads per group
    int gbytes;       // bytes per group = G * L (multiple of 4)
    int lanes;        // busy lanes per thread-group = sum of words per row
    int lps;          // lanes per thread-group (= lanes)
    int ng;           // thread-groups (groups in flight) per CTA
    int threads;      // consumer threads: ng * lps rounded up to whole warps
    int nprod;        // producer warps behind the consumers (1 with one read group or a segmented batch, up to 8 otherwise)
    int contig;       // one read group or a segmented batch: a stage is one contiguous span of groups and the work list,
                      // if one is needed at all, is the identity (its records are read where they lie, not staged)
    int pad_[3];      // keeps what follows this struct in the kernel parameters (TableCfg: the factors of the address
                      // dot products) where it was modulo 16 bytes: ptxas fetches four of them with one LDCU.128 per
                      // word when they are 16-byte aligned, and with four LDC (+ 5 % instructions) when they are not
    int sj;           // plane stride (words)
    int row;          // words per quality row = 4 * sj
    int minscore;     // first tallied quality
    int nqv;          // tallied quality rows = 43 - minscore (row nqv is the trash row)
    int wstart[MAX_G + 1];  // first lane of every row inside the thread-group
};

__host__ __device__ inline int gcd_int(int a, int b) {
    while (b) { int t = a % b; a = b; b = t; }
    return a;
}

// `single_rg`: with one read group every row of a group is tallied, so the group only has to be
// large enough to leave at most 16 thread-groups (several groups per thread-group and stage then fit
// the 32-groups-per-stage limit).  With several read groups the smallest group is taken: mates of a
// pair normally share their read group, so a group of one pair is staged once, while a larger group
// would be staged once per read group present in it.
inline bool make_geom(int L, int minscore, bool single_rg, int nprod, Geom *g, bool pow2_groups = false) {
    if (L < 4 || minscore < 0 || minscore >= NQ) return false;
    const int rps = 4 / gcd_int(L, 4);
    // With several read groups every group of a stage is a copy of its own per array and one warp
    // issues a bulk copy every ~90 cycles, so `nprod` producer warps share the stages round-robin
    // (measured, R = 8 x 150 bp build: 2 warps 2.93 ms, 4 warps 2.13 ms, 8 warps 1.74 ms).
    if (single_rg) nprod = 1;
    const int budget = MAX_THREADS - 32 * nprod;
    int best_g = 0, best_lanes = 0;
    for (int G = rps; G <= MAX_G; G += rps) {
        if (pow2_groups && (G & (G - 1))) continue;  // the spans of a segmented batch start at multiples of 16 rows
        int lanes = 0;
        for (int k = 0; k < G; ++k) lanes += ((k * L) % 4 + L + 3) / 4;
        if (lanes > budget) break;
        if (best_g == 0) { best_g = G; best_lanes = lanes; continue; }
        if (!single_rg) break;
        // fewer lanes per read wins; on a tie a larger group only while the smaller one leaves > 16 thread-groups
        const long long a = (long long)lanes * best_g, b = (long long)best_lanes * G;
        if (a < b || (a == b && budget / best_lanes > 16)) { best_g = G; best_lanes = lanes; }
    }
    if (best_g == 0) return false;
    g->L = L;
    g->G = best_g;
    g->gbytes = best_g * L;
    g->lanes = best_lanes;
    g->lps = best_lanes;
    g->nprod = nprod;
    g->contig = single_rg ? 1 : 0;
    // several read groups: a stage holds at most 32 groups (one per producer lane)
    g->ng = single_rg ? budget / g->lps : std::min(32, budget / g->lps);
    g->threads = (g->ng * g->lps + 31) / 32 * 32;
    int w = 0;
    for (int k = 0; k <= MAX_G; ++k) {
        g->wstart[k] = w;
        if (k < best_g) w += ((k * L) % 4 + L + 3) / 4;
    }
    const int planes = (2 * L + 3) / 4;
    g->sj = (planes + 7) / 8 * 8;
    g->row = 4 * g->sj;
    g->minscore = minscore;
    g->nqv = NQ - minscore;
    return true;
}

// position of cycle-axis index c2 (0 .. 2L-1) inside a shared-memory quality row
__host__ __device__ inline int plane_pos(int c2, int sj) { return (c2 & 3) * sj + (c2 >> 2); }

// Work-list record: one group of G reads seen from one read group, ready to be dropped into a
// stage header by the copy engine (stage.cuh):
//   x  byte offset of the group inside the stage's data area (the prepare pass knows where every
//      record lands: which CTA, which iteration, which slot)
//   y  group index
//   z  flag bytes of rows 0-3, w of rows 4-7: 0 = row not tallied in this segment (other read
//      group / past the end), 1 = read 1, 3 = read 2 of a pair
typedef uint4 entry_t;

// natural 2-bit code (b >> 1) & 3: A=0 C=1 T=2 G=3  ->  reference order A=0 T=1 G=2 C=3
// (Dinucleotide.nucleotides, kbbq/compare_reads.py:199)
__host__ __device__ inline int nat_to_ref(int nat) { return (0x9C >> (2 * nat)) & 3; }

}  // namespace kbbq

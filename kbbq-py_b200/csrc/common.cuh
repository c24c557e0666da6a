// common.cuh -- shared definitions of the B200 kbbq hot-path kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/kbbq_b200.h"
#include "../../include/kbbq_model_constants.h"

#ifndef KBBQ_SM_COUNT_FALLBACK
#define KBBQ_SM_COUNT_FALLBACK 148  // B200: 2 dies x 74 SMs
#endif

namespace kbbq {

constexpr int NQ = KBBQ_NQ;            // 43 quality rows (0..42)
constexpr int MAX_THREADS = 1024;
constexpr uint32_t H4 = 0x80808080u;   // high bit of every byte
constexpr uint32_t ONE4 = 0x01010101u;

// Launch counter (bench.py reports it as gpu_launches).
extern long long g_launches;
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
// Geometry of the "super-row" mapping shared by the build and apply kernels.
//
// Reads are u8 rows of length L packed back to back, so a row starts at byte r*L, which is in
// general only 1- or 2-byte aligned.  A super-row is the smallest run of RPS consecutive reads
// whose byte length is a multiple of 4 (RPS = 4 / gcd(L, 4)); it starts 4-byte aligned and is
// WPS = RPS*L/4 words long.  One thread owns word j of every super-row its group processes, so
// the (row-in-super-row, cycle) of each of its 4 bytes is fixed for the whole kernel and the
// shared-memory table offsets are computed once.  A group is LPS = roundup(WPS, 32) lanes, i.e.
// whole warps, so a warp never mixes two super-rows.
//
// Shared-memory cycle tables are laid out [q - minscore][plane = c2 & 3][c2 >> 2] with a plane
// stride SJ that makes the row stride a multiple of 32 words: lanes with consecutive j then
// touch consecutive banks at every byte position, whatever their qualities.
// ---------------------------------------------------------------------------------------------
struct Geom {
    int L;         // read length
    int rps;       // reads per super-row
    int srb;       // bytes per super-row
    int wps;       // words per super-row
    int lps;       // lanes per super-row (multiple of 32)
    int ng;        // groups (super-rows in flight) per CTA
    int threads;   // ng * lps
    int sj;        // plane stride (words)
    int row;       // words per quality row = 4 * sj
    int minscore;  // first tallied quality
    int nqv;       // tallied quality rows = 43 - minscore
};

__host__ __device__ inline int gcd_int(int a, int b) {
    while (b) { int t = a % b; a = b; b = t; }
    return a;
}

inline bool make_geom(int L, int minscore, Geom *g) {
    if (L < 4 || minscore < 0 || minscore >= NQ) return false;
    g->L = L;
    g->rps = 4 / gcd_int(L, 4);
    g->srb = g->rps * L;
    g->wps = g->srb / 4;
    g->lps = (g->wps + 31) / 32 * 32;
    if (g->lps > MAX_THREADS) return false;
    g->ng = MAX_THREADS / g->lps;
    g->threads = g->ng * g->lps;
    int planes = (2 * L + 3) / 4;
    g->sj = (planes + 7) / 8 * 8;
    g->row = 4 * g->sj;
    g->minscore = minscore;
    g->nqv = NQ - minscore;
    return true;
}

// position of cycle-axis index c2 (0 .. 2L-1) inside a shared-memory quality row
__host__ __device__ inline int plane_pos(int c2, int sj) { return (c2 & 3) * sj + (c2 >> 2); }

// Work-list entry: one super-row seen from one read group.
//   bits  0..31  super-row index
//   bits 32..35  rows of the super-row that belong to this read group (and exist)
//   bits 36..39  `second` flag of each row
typedef unsigned long long entry_t;

// natural 2-bit code (b >> 1) & 3: A=0 C=1 T=2 G=3  ->  reference order A=0 T=1 G=2 C=3
// (Dinucleotide.nucleotides, kbbq/compare_reads.py:199)
__host__ __device__ inline int nat_to_ref(int nat) { return (0x9C >> (2 * nat)) & 3; }

}  // namespace kbbq

// prepare.cuh -- builds the per-read-group work list the build and apply kernels iterate over.
//
// The shared-memory tables of one CTA hold ONE read group (a cycle table for 32 read groups of
// 250 bp reads would need 2.4 MB).  Reads of all read groups are interleaved in the batch, so a tiny
// pre-pass buckets GROUP indices (not data) by read group: entries[seg[g] .. seg[g+1]) lists the
// groups of G reads (common.cuh) that contain at least one read of read group g, with one flag byte
// per row (is it of this read group, is it read 2) and the place the group will take in the
// staging ring.  The main kernels walk contiguous slices of that list; the read data itself is
// never moved.  Cost: 3 B/read in, 16 B/group out and back in (< 1 % of the 3 B/base the build reads).
#pragma once
#include "common.cuh"

namespace kbbq {

constexpr int PREP_THREADS = 256;
constexpr int PREP_SMEM_RG = 8192;  // read groups countable in shared memory

struct PrepArgs {
    const uint16_t *rg;     // may be NULL (all zero)
    const uint8_t *second;  // may be NULL (all zero)
    long long N;            // reads
    long long ngroups;      // ceil(N / G)
    int G;
    int R;
    unsigned int *seg;      // [R + 1] segment offsets (out)
    unsigned int *cursor;   // [R] scratch
    entry_t *entries;       // (out)
    unsigned int *uni;      // (out) [0] != 0: some group's row flags differ from group 0's; [1], [2]: those flags
    int *status;
    // staging geometry of the kernel that will walk the list (stage.cuh)
    int grid;               // its number of CTAs
    int ng;                 // groups per stage
    unsigned int gbytes;    // bytes per group
    unsigned int slot;      // bytes reserved per group when a stage is not one contiguous span
};

// flag bytes from the rows of the group that are tallied (`match`) and their `second` bits
__device__ __forceinline__ entry_t make_entry(unsigned int soff, long long grp, unsigned int match, unsigned int sec) {
    unsigned int f[2] = {0u, 0u};
#pragma unroll
    for (int k = 0; k < MAX_G; ++k) {
        const unsigned int live = (match >> k) & 1u, s2 = (sec >> k) & 1u;
        f[k >> 2] |= (live | ((live & s2) << 1)) << (8 * (k & 3));
    }
    return make_uint4(soff, (unsigned int)grp, f[0], f[1]);
}

// First list position of the slice CTA b walks (the kernels use the same expression).
__device__ __forceinline__ unsigned long long slice_lo(unsigned long long E, unsigned long long b, int grid) {
    return E * b / (unsigned long long)grid;
}
// The CTA whose slice holds list position i: lo(b) <= i < lo(b + 1).
__device__ __forceinline__ unsigned long long slice_of(unsigned long long E, unsigned long long i, int grid) {
    return ((i + 1) * (unsigned long long)grid + E - 1) / E - 1;
}

// rows of group `grp`: rg value (0xFFFFFFFF if the row does not exist) and second bits
__device__ __forceinline__ void load_rows(const PrepArgs &a, long long grp, unsigned int rgv[MAX_G],
                                          unsigned int &sec, unsigned int &exist) {
    sec = 0;
    exist = 0;
#pragma unroll
    for (int k = 0; k < MAX_G; ++k) {
        rgv[k] = 0xFFFFFFFFu;
        if (k < a.G) {
            const long long r = grp * a.G + k;
            if (r < a.N) {
                rgv[k] = a.rg ? a.rg[r] : 0u;
                exist |= 1u << k;
                if (a.second && a.second[r]) sec |= 1u << k;
            }
        }
    }
}

// flags of a group with one read group: rows that exist and have rg == 0 (anything else is an error)
__device__ __forceinline__ void identity_rows(const PrepArgs &a, long long grp, unsigned int &exist, unsigned int &sec,
                                              bool report) {
    unsigned int rgv[MAX_G];
    load_rows(a, grp, rgv, sec, exist);
    if (a.rg) {
        unsigned int ok = 0;
#pragma unroll
        for (int k = 0; k < MAX_G; ++k)
            if ((exist >> k) & 1) {
                if (rgv[k] == 0) ok |= 1u << k;
                else if (report) atomicOr(a.status, KBBQ_FLAG_RG_RANGE);
            }
        exist = ok;
    }
}

// Single read group, pass 1: do all groups look like group 0, with every row tallied?  Then the
// kernels need no work list at all (uni[0] stays 0) and pass 2 returns at once.  Reads 1-3 B/read.
__global__ void prep_uniform_kernel(PrepArgs a) {
    const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    if (t0 == 0) {
        a.seg[0] = 0;
        a.seg[1] = (unsigned int)a.ngroups;
    }
    unsigned int exist0, sec0;
    identity_rows(a, 0, exist0, sec0, false);
    bool differs = exist0 != (1u << a.G) - 1u;
    // the uniform walk keeps one cycle table per row parity (build.cuh: make_thread_map): rows of equal parity must
    // be the same mate
    for (int k = 2; k < a.G; ++k) differs |= ((sec0 >> k) & 1u) != ((sec0 >> (k & 1)) & 1u);
    // Whole groups of a power-of-two G repeat every 16 rows: compare 16 mate flags (and 8 read-group numbers) per
    // load against the pattern of group 0 instead of walking the rows one byte at a time.
    const bool wide = (a.G & (a.G - 1)) == 0 && a.G <= 16 && (!a.second || ((uintptr_t)a.second & 15u) == 0) &&
                      (!a.rg || ((uintptr_t)a.rg & 15u) == 0);
    long long done_groups = 0;
    if (wide) {
        const long long chunks = a.N / 16;   // rows [0, 16 * chunks)
        uint4 pat;
        {
            unsigned int v[4];
            for (int j = 0; j < 4; ++j) {
                v[j] = 0;
                for (int k = 0; k < 4; ++k) v[j] |= ((sec0 >> ((4 * j + k) % a.G)) & 1u) << (8 * k);
            }
            pat = make_uint4(v[0], v[1], v[2], v[3]);
        }
        for (long long c = t0; c < chunks; c += stride) {
            if (a.second) {
                const uint4 s = reinterpret_cast<const uint4 *>(a.second)[c];
                // second[] holds any non-zero value for "read 2": normalise to 0 / 1 per byte
                auto norm = [](unsigned int x) { return ((x | ((x | 0x80808080u) - 0x01010101u)) >> 7) & 0x01010101u; };
                differs |= norm(s.x) != pat.x || norm(s.y) != pat.y || norm(s.z) != pat.z || norm(s.w) != pat.w;
            } else {
                differs |= sec0 != 0;
            }
            if (a.rg) {
                const uint4 r0 = reinterpret_cast<const uint4 *>(a.rg)[2 * c], r1 = reinterpret_cast<const uint4 *>(a.rg)[2 * c + 1];
                if (r0.x | r0.y | r0.z | r0.w | r1.x | r1.y | r1.z | r1.w) { differs = true; atomicOr(a.status, KBBQ_FLAG_RG_RANGE); }
            }
        }
        done_groups = chunks * 16 / a.G;
    }
    for (long long grp = done_groups + t0; grp < a.ngroups; grp += stride) {  // the rest (or every group), row by row
        unsigned int exist, sec;
        identity_rows(a, grp, exist, sec, true);
        differs |= exist != exist0 || sec != sec0;
    }
    if (differs) a.uni[0] = 1u;
    if (t0 == 0) {
        const entry_t p0 = make_entry(0u, 0, exist0, sec0);
        a.uni[1] = p0.z;
        a.uni[2] = p0.w;
    }
}

// Single read group, pass 2 (only when the groups differ): the list is the identity.
__global__ void prep_identity_kernel(PrepArgs a) {
    if (a.uni[0] == 0u) return;
    const unsigned long long E = (unsigned long long)a.ngroups;
    for (long long grp = (long long)blockIdx.x * blockDim.x + threadIdx.x; grp < a.ngroups;
         grp += (long long)gridDim.x * blockDim.x) {
        unsigned int exist, sec;
        identity_rows(a, grp, exist, sec, false);
        // a stage is one contiguous span: ng consecutive groups from the start of the CTA's slice, the
        // first one 16-byte aligned down
        const unsigned long long lo = slice_lo(E, slice_of(E, (unsigned long long)grp, a.grid), a.grid);
        const unsigned int j = (unsigned int)(((unsigned long long)grp - lo) % (unsigned int)a.ng);
        const unsigned int mis0 = (unsigned int)((((unsigned long long)grp - j) * a.gbytes) & 15ull);
        a.entries[grp] = make_entry(mis0 + j * a.gbytes, grp, exist, sec);
    }
}

// mode 0: count list entries per read group into a.cursor; mode 1: scatter entries.
// One group per thread; per-block counts in shared memory, one global atomic per (block, read group).
template <int MODE>
__global__ void __launch_bounds__(PREP_THREADS) prep_bucket_kernel(PrepArgs a) {
    extern __shared__ unsigned int s_cnt[];  // [R] when R <= PREP_SMEM_RG
    const bool use_smem = a.R <= PREP_SMEM_RG;
    if (use_smem) {
        for (int i = threadIdx.x; i < a.R; i += blockDim.x) s_cnt[i] = 0;
        __syncthreads();
    }
    const long long grp = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int rgv[MAX_G], exist = 0, sec = 0;
    unsigned int slot[MAX_G], mask[MAX_G];
    if (grp < a.ngroups) load_rows(a, grp, rgv, sec, exist);
#pragma unroll
    for (int k = 0; k < MAX_G; ++k) {
        mask[k] = 0;
        slot[k] = 0;
        if (!((exist >> k) & 1)) continue;
        const unsigned int v = rgv[k];
        if (v >= (unsigned int)a.R) {
            atomicOr(a.status, KBBQ_FLAG_RG_RANGE);
            continue;
        }
        bool first = true;  // is row k the first row of this group with value v?
        unsigned int m = 0;
#pragma unroll
        for (int k2 = 0; k2 < MAX_G; ++k2) {
            if (((exist >> k2) & 1) && rgv[k2] == v) {
                if (k2 < k) first = false;
                m |= 1u << k2;
            }
        }
        if (!first) continue;
        mask[k] = m;
        if (use_smem) slot[k] = atomicAdd(&s_cnt[v], 1u);
        else slot[k] = atomicAdd(&a.cursor[v], 1u);
    }
    if (use_smem) {
        __syncthreads();
        for (int i = threadIdx.x; i < a.R; i += blockDim.x) {
            const unsigned int c = s_cnt[i];
            if (c) s_cnt[i] = atomicAdd(&a.cursor[i], c);  // reserve the block's share
        }
        if (MODE == 0) return;
        __syncthreads();
    }
    if (MODE == 1) {
        const unsigned long long E = a.seg[a.R];
#pragma unroll
        for (int k = 0; k < MAX_G; ++k) {
            if (!mask[k]) continue;
            const unsigned int pos = slot[k] + (use_smem ? s_cnt[rgv[k]] : 0u);
            // iterations restart at the start of every (CTA slice, segment) intersection; every
            // group has its own slot in the stage
            const unsigned long long lo = slice_lo(E, slice_of(E, pos, a.grid), a.grid);
            const unsigned long long start = max((unsigned long long)a.seg[rgv[k]], lo);
            const unsigned int j = (unsigned int)((pos - start) % (unsigned int)a.ng);
            const unsigned int mis = (unsigned int)(((unsigned long long)grp * a.gbytes) & 15ull);
            a.entries[pos] = make_entry(j * a.slot + mis, grp, mask[k], sec);
        }
    }
}

// Exclusive scan of the per-read-group counts (one block): seg[0..R], cursor[g] = seg[g].
__global__ void prep_scan_kernel(PrepArgs a) {
    __shared__ unsigned int s_part[1024];
    const int T = blockDim.x;
    const int per = (a.R + T - 1) / T;
    const int lo = threadIdx.x * per, hi = min(a.R, lo + per);
    unsigned int sum = 0;
    for (int i = lo; i < hi; ++i) sum += a.cursor[i];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int run = 0;
        for (int t = 0; t < T; ++t) {
            const unsigned int v = s_part[t];
            s_part[t] = run;
            run += v;
        }
        a.seg[a.R] = run;
    }
    __syncthreads();
    unsigned int run = s_part[threadIdx.x];
    for (int i = lo; i < hi; ++i) {
        const unsigned int v = a.cursor[i];
        a.seg[i] = run;
        a.cursor[i] = run;
        run += v;
    }
}

// Workspace carve-up shared by build and apply.
struct Workspace {
    unsigned int *seg;
    unsigned int *cursor;
    unsigned int *uni;     // [4]
    entry_t *entries;
    short *fold_cyc;   // [R][43][2L]   apply only
    short *fold_din;   // [R][43][16]   apply only (natural dinuc order, relative to the pad column)
    size_t bytes;
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

inline Workspace carve_workspace(void *base, long long N, int L, int R) {
    Workspace w = {};
    size_t off = 0;
    char *p = (char *)base;
    // a group never lists more entries than it has rows, so N (+ padding) entries always suffice;
    // with one read group the list is one entry per group of at least 1 read
    const size_t max_entries = (size_t)N + MAX_G;
    w.seg = (unsigned int *)(p + off);      off = align_up(off + sizeof(unsigned int) * (2 * (size_t)R + 2), 256);  // [R + 1], or [2R + 1] spans of a segmented batch
    w.cursor = (unsigned int *)(p + off);   off = align_up(off + sizeof(unsigned int) * (size_t)R, 256);
    w.uni = (unsigned int *)(p + off);      off = align_up(off + sizeof(unsigned int) * 4, 256);
    w.entries = (entry_t *)(p + off);       off = align_up(off + sizeof(entry_t) * max_entries, 256);
    w.fold_cyc = (short *)(p + off);        off = align_up(off + sizeof(short) * (size_t)R * NQ * 2 * L, 256);
    w.fold_din = (short *)(p + off);        off = align_up(off + sizeof(short) * (size_t)R * NQ * 16, 256);
    w.bytes = off;
    return w;
}

// Enqueue the pre-pass.  After it, seg[R] (device) holds the number of entries.
inline int run_prepare(const uint16_t *rg, const uint8_t *second, long long N, int G, int R,
                       const Workspace &w, int *status, int grid, int ng, unsigned int gbytes,
                       unsigned int slot, cudaStream_t st) {
    PrepArgs a;
    a.rg = rg; a.second = second; a.N = N; a.G = G;
    a.ngroups = (N + G - 1) / G; a.R = R;
    a.seg = w.seg; a.cursor = w.cursor; a.entries = w.entries; a.uni = w.uni; a.status = status;
    a.grid = grid; a.ng = ng; a.gbytes = gbytes; a.slot = slot;
    // several read groups (or nothing to do): not uniform; one read group: the identity pass decides
    KBBQ_CUDA(cudaMemsetAsync(w.uni, (R == 1 && a.ngroups > 0) ? 0 : 0xFF, sizeof(unsigned int) * 4, st));
    if (a.ngroups == 0) {
        KBBQ_CUDA(cudaMemsetAsync(w.seg, 0, sizeof(unsigned int) * ((size_t)R + 1), st));
        return KBBQ_OK;
    }
    const unsigned int blocks = (unsigned int)((a.ngroups + PREP_THREADS - 1) / PREP_THREADS);
    if (R == 1) {
        const unsigned int few = std::min<unsigned int>(blocks, (unsigned int)std::max(grid, 1) * 8u);
        prep_uniform_kernel<<<few, PREP_THREADS, 0, st>>>(a);
        KBBQ_LAUNCHED();
        prep_identity_kernel<<<few, PREP_THREADS, 0, st>>>(a);
        KBBQ_LAUNCHED();
        return KBBQ_OK;
    }
    const size_t smem = R <= PREP_SMEM_RG ? sizeof(unsigned int) * (size_t)R : 0;
    KBBQ_CUDA(cudaMemsetAsync(w.cursor, 0, sizeof(unsigned int) * (size_t)R, st));
    prep_bucket_kernel<0><<<blocks, PREP_THREADS, smem, st>>>(a);
    KBBQ_LAUNCHED();
    prep_scan_kernel<<<1, 1024, 0, st>>>(a);
    KBBQ_LAUNCHED();
    prep_bucket_kernel<1><<<blocks, PREP_THREADS, smem, st>>>(a);
    KBBQ_LAUNCHED();
    return KBBQ_OK;
}

}  // namespace kbbq

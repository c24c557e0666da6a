// stage.cuh -- TMA bulk-copy staging ring shared by the build and apply kernels.
//
// A dedicated producer warp streams the reads of the CTA's work-list slice from HBM into a ring of
// shared-memory stages with 1-D bulk copies (cp.async.bulk.shared::cluster.global, SASS UBLKCP),
// completion signalled on an mbarrier per stage; the consumer warps wait on that barrier, pull
// their 32-bit words out of the stage into registers and release it through a second mbarrier.
// HBM is therefore read exactly once, in whole 16-byte units, with no load instructions or
// address arithmetic in the consumer's hot loop, and the copy engine runs ahead of the tally.
//
// One stage holds the `ng` groups (common.cuh) the thread-groups of the CTA process in one
// iteration, plus their `ng` work-list records (prepare.cuh), which the copy engine drops into
// the stage header as they are: the record tells a thread-group where its group starts inside the
// stage and, one flag byte per row, which rows are tallied (0 = not in this segment, 1 = read 1,
// 3 = read 2 of a pair).  With one read group the list is the identity, so the producer issues ONE
// copy per array of the aligned span that covers the ng consecutive groups and never reads the
// list itself -- per stage it is one elected lane issuing four copies.  With several read groups
// every group has its own slot and the lanes of the producer warp issue one copy per group and array.
// Copies never reach past the 16-byte unit that holds the last byte of the arrays: callers must
// make the arrays readable up to there (any cudaMalloc'ed or torch buffer is).
#pragma once
#include "common.cuh"
#include "prepare.cuh"

namespace kbbq {

constexpr int MAX_STAGES = 8;

struct StageLayout {
    int stages;       // ring depth
    int kps;          // groups every thread-group takes out of one stage
    int ngs;          // groups per stage = ng * kps
    int narr;         // arrays staged (3 for build, 2 for apply)
    int abytes;       // bytes of one array inside a stage (multiple of 16: bulk-copy destinations)
    int slot;         // bytes reserved per group when the groups of a stage are not contiguous
    int data_off;     // byte offsets from the start of dynamic shared memory
    int hdr_off;      // stages x ng work-list records (entry_t); none with contiguous spans (Geom::contig)
    int bar_off;      // full[stages], empty[stages] (8 bytes each)
    int total;        // dynamic shared memory bytes including the tables in front
};

// With contiguous spans (one read group, a segmented batch) a stage is one span per array -- no per-group slots with
// their alignment slack -- and no work-list records are staged: the uniform walk needs none and the rare batch whose
// groups differ (unpaired mates in no order) reads the identity list from global memory.  Every byte counts here: at
// 150 bp this is what lets a third stage of three groups per thread-group fit next to the tables.
inline StageLayout make_stage_layout(const Geom &g, int narr, int stages, int kps, size_t table_bytes) {
    StageLayout s;
    s.stages = stages;
    s.kps = kps;
    s.ngs = g.ng * kps;
    s.narr = narr;
    s.slot = (g.gbytes + 15 + 15) / 16 * 16;           // group + worst-case misalignment, 16-byte units
    s.abytes = g.contig ? (s.ngs * g.gbytes + 32 + 15) / 16 * 16 : (s.ngs * s.slot + 127) / 128 * 128;
    s.data_off = (int)((table_bytes + 127) / 128 * 128);
    s.hdr_off = s.data_off + stages * narr * s.abytes;
    s.bar_off = (s.hdr_off + (g.contig ? 0 : stages * s.ngs * 16) + 15) / 16 * 16;
    s.total = s.bar_off + 2 * stages * 8;
    return s;
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
// 1-D bulk copy global -> shared (TMA), 16-byte aligned addresses and size
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void consumer_sync(int nthreads) {
    asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
}

// What the producer needs to know about the CTA's job.
struct ProducerArgs {
    const uint8_t *arr[3];
    const entry_t *entries;
    const unsigned int *seg;
    int nsub;             // sub-segments (BuildArgs::nsub)
    bool contig;          // one read group or a segmented batch: every stage is one contiguous span of groups
    uint32_t lo, hi;      // this CTA's slice of the work list
    uint32_t gbytes;
    int ng;                 // groups per stage (StageLayout::ngs)
    long long total_bytes;  // bytes of each array (N * L)
    int pw, nprod;          // this producer warp / number of producer warps
    bool uniform;           // one read group, no work list: the consumers do not read the stage headers
};

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

// Producer warp: walks the same (segment, iteration) sequence as the consumers.
__device__ __forceinline__ void producer_loop(const ProducerArgs &p, const StageLayout &sl, unsigned char *smem_base) {
    const int lane = threadIdx.x & 31;
    const uint32_t data0 = smem_u32(smem_base + sl.data_off);
    const uint32_t hdr0 = smem_u32(smem_base + sl.hdr_off);
    const uint32_t bar0 = smem_u32(smem_base + sl.bar_off);
    uint32_t stage = 0, phase = 0;
    const unsigned long long end16 = ((unsigned long long)p.total_bytes + 15ull) & ~15ull;

    if (p.contig) {  // identity list: one contiguous span per stage, one producer warp
        if (lane != 0 || p.pw != 0) return;
        for (int sub = 0; sub < p.nsub; ++sub) {
            uint32_t s_lo = p.seg[sub], s_hi = p.seg[sub + 1];
            if (s_hi <= p.lo) continue;
            if (s_lo >= p.hi) break;
            if (s_lo < p.lo) s_lo = p.lo;
            if (s_hi > p.hi) s_hi = p.hi;
            for (uint32_t first = s_lo; first < s_hi; first += p.ng) {
                const uint32_t n = min((uint32_t)p.ng, s_hi - first);
                const uint32_t full = bar0 + stage * 8, empty = bar0 + (sl.stages + stage) * 8;
                const unsigned long long start = (unsigned long long)first * p.gbytes;
                const uint32_t mis = (uint32_t)start & 15u;
                const unsigned long long src = start - mis;
                uint32_t bytes = (mis + n * p.gbytes + 15u) & ~15u;
                // the last group of a batch may be partial: stop at the end of the arrays
                bytes = src >= end16 ? 0u : (uint32_t)min((unsigned long long)bytes, end16 - src);
                mbar_wait(empty, phase ^ 1);
                // (no records are staged: a batch whose groups differ reads the identity list from global memory)
                mbar_arrive_expect_tx(full, bytes * sl.narr);
                const uint32_t dst = data0 + stage * sl.narr * sl.abytes;
                if (bytes) {
#pragma unroll
                    for (int k = 0; k < 3; ++k)
                        if (k < sl.narr) bulk_g2s(dst + k * sl.abytes, p.arr[k] + src, bytes, full);
                }
                if (++stage == (uint32_t)sl.stages) { stage = 0; phase ^= 1; }
            }
        }
        return;
    }

    // Several read groups: lane j copies group j of the stage into slot j; the producer warps take
    // the iterations round-robin (one warp issues a bulk copy every ~90 cycles).  The group indices of
    // a warp's next iteration are fetched while the current one is issued.
    uint32_t it = 0;  // iteration counter over all segments: stage = it % stages
    for (int rg = 0; rg < p.nsub; ++rg) {
        uint32_t s_lo = p.seg[rg], s_hi = p.seg[rg + 1];
        if (s_hi <= p.lo) continue;
        if (s_lo >= p.hi) break;
        if (s_lo < p.lo) s_lo = p.lo;
        if (s_hi > p.hi) s_hi = p.hi;
        const uint32_t iters = (s_hi - s_lo + p.ng - 1) / p.ng;
        // first iteration of this segment that is this warp's: it0 + i with (it0 + i) % nprod == pw
        uint32_t i = (uint32_t)((p.pw + p.nprod - (int)(it % p.nprod)) % p.nprod);
        auto fetch = [&](uint32_t iter) -> uint32_t {
            const unsigned long long e = (unsigned long long)s_lo + (unsigned long long)iter * p.ng + lane;
            return (iter < iters && lane < p.ng && e < s_hi) ? __ldg(&p.entries[e].y) : 0u;
        };
        uint32_t nxt = fetch(i);
        for (; i < iters; i += p.nprod) {
            const uint32_t first = s_lo + i * p.ng;
            const uint32_t grp = nxt;
            nxt = fetch(i + p.nprod);
            const uint32_t g_it = it + i;
            stage = g_it % (uint32_t)sl.stages;
            phase = (g_it / (uint32_t)sl.stages) & 1u;
            const uint32_t n = min((uint32_t)p.ng, s_hi - first);
            const uint32_t full = bar0 + stage * 8, empty = bar0 + (sl.stages + stage) * 8;
            const unsigned long long start = (unsigned long long)grp * p.gbytes;
            const uint32_t mis = (uint32_t)start & 15u;
            const unsigned long long src = start - mis;
            uint32_t bytes = 0;
            if ((uint32_t)lane < n) {
                bytes = (mis + p.gbytes + 15u) & ~15u;
                bytes = src >= end16 ? 0u : (uint32_t)min((unsigned long long)bytes, end16 - src);
            }
            if (lane == 0) mbar_wait(empty, phase ^ 1);
            __syncwarp();
            if (bytes) {
                mbar_expect_tx(full, bytes * sl.narr);
                const uint32_t dst = data0 + stage * sl.narr * sl.abytes + lane * sl.slot;
#pragma unroll
                for (int k = 0; k < 3; ++k)
                    if (k < sl.narr) bulk_g2s(dst + k * sl.abytes, p.arr[k] + src, bytes, full);
            }
            if ((uint32_t)lane >= n && lane < p.ng)  // a short last stage: "no rows" past the list
                asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(hdr0 + (stage * p.ng + lane) * 16u), "r"(0u) : "memory");
            if (lane == 0) {
                mbar_expect_tx(full, n * 16u);
                bulk_g2s(hdr0 + stage * p.ng * 16u, p.entries + first, n * 16u, full);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(full);  // after every expect_tx and header store of the stage
        }
        it += iters;
    }
}

}  // namespace kbbq

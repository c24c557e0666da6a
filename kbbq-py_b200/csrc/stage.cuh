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
// iteration.  When the ng work-list entries are consecutive groups (always, with one read group)
// the producer issues ONE copy per array of the aligned span that covers them; otherwise one copy
// per group and array into fixed slots.  A header per group slot tells the consumers where the
// group starts inside the stage (`soff`) and which of its rows belong to the read group (`bits`).
#pragma once
#include "common.cuh"
#include "prepare.cuh"

namespace kbbq {

constexpr int MAX_STAGES = 4;

struct StageLayout {
    int stages;       // ring depth
    int narr;         // arrays staged (3 for build, 2 for apply)
    int abytes;       // bytes of one array inside a stage (multiple of 128)
    int slot;         // bytes reserved per group when the groups of a stage are not contiguous
    int data_off;     // byte offsets from the start of dynamic shared memory
    int hdr_off;      // stages x ng x uint4 {bits, soff, group index, 0}
    int bar_off;      // full[stages], empty[stages] (8 bytes each)
    int total;        // dynamic shared memory bytes including the tables in front
};

inline StageLayout make_stage_layout(const Geom &g, int narr, int stages, size_t table_bytes) {
    StageLayout s;
    s.stages = stages;
    s.narr = narr;
    s.slot = (g.gbytes + 15 + 15) / 16 * 16;           // group + worst-case misalignment, 16-byte units
    s.abytes = (g.ng * s.slot + 127) / 128 * 128;      // also covers the contiguous span (ng*gbytes + 30)
    s.data_off = (int)((table_bytes + 127) / 128 * 128);
    s.hdr_off = s.data_off + stages * narr * s.abytes;
    s.bar_off = s.hdr_off + stages * g.ng * 16;
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
    int R;
    uint32_t lo, hi;      // this CTA's slice of the work list
    uint32_t gbytes;
    int ng;
};

// Producer warp: walks the same (segment, iteration) sequence as the consumers.
__device__ __forceinline__ void producer_loop(const ProducerArgs &p, const StageLayout &sl, unsigned char *smem_base) {
    const int lane = threadIdx.x & 31;
    const uint32_t data0 = smem_u32(smem_base + sl.data_off);
    uint4 *hdr = reinterpret_cast<uint4 *>(smem_base + sl.hdr_off);
    const uint32_t bar0 = smem_u32(smem_base + sl.bar_off);
    uint32_t stage = 0, phase = 0;
    for (int rg = 0; rg < p.R; ++rg) {
        uint32_t s_lo = p.seg[rg], s_hi = p.seg[rg + 1];
        if (s_hi <= p.lo) continue;
        if (s_lo >= p.hi) break;
        if (s_lo < p.lo) s_lo = p.lo;
        if (s_hi > p.hi) s_hi = p.hi;
        for (uint32_t first = s_lo; first < s_hi; first += p.ng) {
            const uint32_t n = min((uint32_t)p.ng, s_hi - first);
            const uint32_t full = bar0 + stage * 8, empty = bar0 + (sl.stages + stage) * 8;
            if (lane == 0) mbar_wait(empty, phase ^ 1);
            __syncwarp();
            uint32_t grp = 0, bits = 0;
            if ((uint32_t)lane < n) {
                const entry_t e = __ldg(p.entries + first + lane);
                grp = (uint32_t)e;
                bits = (uint32_t)(e >> 32);
            }
            const uint32_t grp0 = __shfl_sync(0xFFFFFFFFu, grp, 0);
            const bool contig = __all_sync(0xFFFFFFFFu, (uint32_t)lane >= n || grp == grp0 + lane);
            const unsigned long long start = (unsigned long long)(contig ? grp0 : grp) * p.gbytes;
            const uint32_t mis = (uint32_t)start & 15u;
            uint32_t soff, bytes;
            if (contig) {
                soff = mis + lane * p.gbytes;
                bytes = (lane == 0) ? (mis + n * p.gbytes + 15u) & ~15u : 0u;
            } else {
                soff = lane * sl.slot + mis;
                bytes = ((uint32_t)lane < n) ? (mis + p.gbytes + 15u) & ~15u : 0u;
            }
            if (lane < p.ng) hdr[stage * p.ng + lane] = make_uint4(bits, soff, grp, 0u);
            uint32_t tx = bytes;
#pragma unroll
            for (int o = 16; o; o >>= 1) tx += __shfl_xor_sync(0xFFFFFFFFu, tx, o);
            __syncwarp();
            if (lane == 0) mbar_arrive_expect_tx(full, tx * sl.narr);
            __syncwarp();
            if (bytes) {
                const uint32_t dst = data0 + stage * sl.narr * sl.abytes + (contig ? 0u : lane * sl.slot);
                const unsigned long long src = start - mis;
#pragma unroll
                for (int k = 0; k < 3; ++k)
                    if (k < sl.narr) bulk_g2s(dst + k * sl.abytes, p.arr[k] + src, bytes, full);
            }
            if (++stage == (uint32_t)sl.stages) { stage = 0; phase ^= 1; }
        }
    }
}

}  // namespace kbbq

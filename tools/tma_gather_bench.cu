// tma_gather_bench.cu -- how fast can one SM gather short rows (a pair of reads: ~300 B per array)
// from HBM?  Three ways, persistent grid of one CTA per SM, rows at pseudo-random 16-byte aligned
// offsets of a buffer much larger than L2:
//   mode 0  cp.async.bulk (UBLKCP) issued by lane 0 of P producer warps, one copy per row
//   mode 1  cp.async.bulk issued by all 32 lanes of P producer warps (the hardware serialises them)
//   mode 2  ld.global.u32 by W warps, a warp reads one row (lanes = consecutive words, rows of <= 128 B x k),
//           U rows in flight per warp, result xor-reduced into a register
//   mode 3  cp.async (LDGSTS) 16 B per lane, 20 lanes per row, U rows in flight per warp
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tma_gather_bench tma_gather_bench.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// Every producer warp owns its own ring of `depth` slots of 32 rows; rows per SM = rows_per_sm.
template <int MODE>
__global__ void __launch_bounds__(1024, 1) k_tma(const uint8_t *buf, uint64_t nunits /* 16-byte units */, uint32_t rowbytes,
                                                 int P, int depth, uint32_t rows_per_warp, unsigned long long *sink) {
    extern __shared__ __align__(128) unsigned char sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp >= P) return;
    const uint32_t slot = (rowbytes + 127) / 128 * 128;
    // per warp: depth stages x 32 rows x slot bytes, + depth barriers
    unsigned char *mine = sm + (size_t)warp * (depth * (32 * slot) + 128);
    const uint32_t bar0 = smem_u32(mine + depth * 32 * slot);
    const uint32_t data0 = smem_u32(mine);
    if (lane == 0)
        for (int s = 0; s < depth; ++s) mbar_init(bar0 + s * 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    uint32_t seed = (blockIdx.x * 64 + warp) * 0x9E3779B9u;
    const uint32_t nst = rows_per_warp / 32;
    uint32_t acc = 0;
    for (uint32_t it = 0; it < nst + depth; ++it) {
        if (it >= (uint32_t)depth) {  // consume stage (it - depth)
            const uint32_t s = (it - depth) % depth, ph = ((it - depth) / depth) & 1;
            mbar_wait(bar0 + s * 8, ph);
            acc += *(volatile uint32_t *)(mine + s * 32 * slot + lane * slot);
            __syncwarp();
        }
        if (it < nst) {
            const uint32_t s = it % depth;
            const uint32_t full = bar0 + s * 8;
            if (MODE == 0) {
                if (lane == 0) {
                    mbar_arrive_expect_tx(full, 32 * rowbytes);
                    for (int j = 0; j < 32; ++j) {
                        const uint64_t u = (uint64_t)hash32(seed + it * 32 + j) * 2654435761ull % nunits;
                        bulk_g2s(data0 + s * 32 * slot + j * slot, buf + u * 16, rowbytes, full);
                    }
                }
            } else {
                const uint64_t u = (uint64_t)hash32(seed + it * 32 + lane) * 2654435761ull % nunits;
                mbar_expect_tx(full, rowbytes);
                bulk_g2s(data0 + s * 32 * slot + lane * slot, buf + u * 16, rowbytes, full);
                __syncwarp();
                if (lane == 0) mbar_arrive(full);
            }
            __syncwarp();
        }
    }
    if (acc == 0x12345u) sink[0] = acc;
}

// mode 2: plain loads.  A warp reads rows of `rowbytes` (<= 384: 3 x 128 B), lanes own words.
template <int U>
__global__ void __launch_bounds__(1024, 1) k_ldg(const uint8_t *buf, uint64_t nunits, uint32_t rowbytes, int W,
                                                 uint32_t rows_per_warp, unsigned long long *sink) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp >= W) return;
    uint32_t seed = (blockIdx.x * 64 + warp) * 0x9E3779B9u;
    uint32_t acc = 0;
    const uint32_t words = rowbytes / 4;
    for (uint32_t it = 0; it < rows_per_warp; it += U) {
        uint32_t v[U][3];
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const uint64_t u = (uint64_t)hash32(seed + it + j) * 2654435761ull % nunits;
            const uint32_t *p = reinterpret_cast<const uint32_t *>(buf + u * 16);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                v[j][k] = 0;
                if (lane + 32 * k < words) asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v[j][k]) : "l"(p + lane + 32 * k));
            }
        }
#pragma unroll
        for (int j = 0; j < U; ++j)
#pragma unroll
            for (int k = 0; k < 3; ++k) acc ^= v[j][k];
    }
    if (acc == 0x12345u) sink[0] = acc;
}

// mode 3: LDGSTS 16 B per lane; a row = ceil(rowbytes/16) lanes; a warp moves floor(32 / lanes_per_row) rows per instruction
__global__ void __launch_bounds__(1024, 1) k_ldgsts(const uint8_t *buf, uint64_t nunits, uint32_t rowbytes, int W, int depth,
                                                   uint32_t rows_per_warp, unsigned long long *sink) {
    extern __shared__ __align__(128) unsigned char sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp >= W) return;
    const uint32_t lpr = (rowbytes + 15) / 16;           // lanes per row
    // one instruction: every lane copies 16 B; lane l -> row l / lpr of this instruction's batch ... simply:
    // the warp handles one row per ceil(lpr/32) instructions when lpr > 32; here lpr <= 32 and a warp takes
    // rpi = 32 / lpr rows per instruction.
    const uint32_t rpi = 32 / lpr;
    const uint32_t myrow = lane / lpr, mycol = lane % lpr;
    const bool active = myrow < rpi;
    unsigned char *mine = sm + (size_t)warp * depth * 512;
    const uint32_t data0 = smem_u32(mine);
    uint32_t seed = (blockIdx.x * 64 + warp) * 0x9E3779B9u;
    uint32_t acc = 0;
    const uint32_t n = rows_per_warp / rpi;
    for (uint32_t it = 0; it < n + depth; ++it) {
        if (it >= (uint32_t)depth) {
            asm volatile("cp.async.wait_group %0;" ::"n"(7) : "memory");  // depth = 8: oldest group done
            acc += *(volatile uint32_t *)(mine + ((it - depth) % depth) * 512 + lane * 16);
        }
        if (it < n && active) {
            const uint64_t u = (uint64_t)hash32(seed + it * rpi + myrow) * 2654435761ull % nunits;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(data0 + (it % depth) * 512 + lane * 16),
                         "l"(buf + (u + mycol) * 16) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    if (acc == 0x12345u) sink[0] = acc;
}

int main(int argc, char **argv) {
    const size_t bytes = 4ull << 30;
    uint8_t *buf;
    unsigned long long *sink;
    CK(cudaMalloc(&buf, bytes + 4096));
    CK(cudaMemset(buf, 1, bytes + 4096));
    CK(cudaMalloc(&sink, 64));
    int sms = 148;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    sms = prop.multiProcessorCount;
    const uint64_t nunits = bytes / 16 - 64;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const int smem = 200 * 1024;
    CK(cudaFuncSetAttribute(k_tma<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(k_tma<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(k_ldgsts, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const uint32_t rows_per_sm = 64 * 1024;
    printf("SMs %d; rows per SM %u\n", sms, rows_per_sm);
    const uint32_t sizes[] = {304, 320, 512, 1024, 4096};
    for (uint32_t rb : sizes) {
        for (int mode = 0; mode < 2; ++mode) {
            for (int P : {1, 2, 4, 8, 16}) {
                const uint32_t slot = (rb + 127) / 128 * 128;
                int depth = (smem / P - 128) / (32 * slot);
                if (depth > 8) depth = 8;
                if (depth < 1) continue;
                const uint32_t rpw = rows_per_sm / P / 32 * 32;
                for (int rep = 0; rep < 2; ++rep) {
                    CK(cudaEventRecord(e0));
                    if (mode == 0) k_tma<0><<<sms, 32 * P, smem>>>(buf, nunits, rb, P, depth, rpw, sink);
                    else k_tma<1><<<sms, 32 * P, smem>>>(buf, nunits, rb, P, depth, rpw, sink);
                    CK(cudaEventRecord(e1));
                    CK(cudaDeviceSynchronize());
                }
                float ms;
                CK(cudaEventElapsedTime(&ms, e0, e1));
                const double tot = (double)rpw * P * sms;
                printf("tma mode %d row %5u B  P %2d depth %d : %7.3f ms  %7.1f GB/s  %6.1f cycles/copy/SM (1.9 GHz)\n", mode, rb, P,
                       depth, ms, tot * rb / ms / 1e6, ms * 1e-3 * 1.9e9 / (rpw * P));
            }
        }
    }
    for (uint32_t rb : {300u, 304u, 320u}) {
        for (int W : {8, 16, 31}) {
            const uint32_t rpw = rows_per_sm / W / 8 * 8;
            for (int U : {1, 2, 4, 8}) {
                for (int rep = 0; rep < 2; ++rep) {
                    CK(cudaEventRecord(e0));
                    if (U == 1) k_ldg<1><<<sms, 32 * W>>>(buf, nunits, rb, W, rpw, sink);
                    if (U == 2) k_ldg<2><<<sms, 32 * W>>>(buf, nunits, rb, W, rpw, sink);
                    if (U == 4) k_ldg<4><<<sms, 32 * W>>>(buf, nunits, rb, W, rpw, sink);
                    if (U == 8) k_ldg<8><<<sms, 32 * W>>>(buf, nunits, rb, W, rpw, sink);
                    CK(cudaEventRecord(e1));
                    CK(cudaDeviceSynchronize());
                }
                float ms;
                CK(cudaEventElapsedTime(&ms, e0, e1));
                const double tot = (double)rpw * W * sms;
                printf("ldg row %u B  W %2d U %d : %7.3f ms  %7.1f GB/s\n", rb, W, U, ms, tot * rb / ms / 1e6);
            }
        }
    }
    for (uint32_t rb : {304u, 320u}) {
        for (int W : {4, 8, 16, 31}) {
            const uint32_t lpr = (rb + 15) / 16, rpi = 32 / lpr;
            const uint32_t rpw = rows_per_sm / W / rpi * rpi;
            for (int rep = 0; rep < 2; ++rep) {
                CK(cudaEventRecord(e0));
                k_ldgsts<<<sms, 32 * W, smem>>>(buf, nunits, rb, W, 8, rpw, sink);
                CK(cudaEventRecord(e1));
                CK(cudaDeviceSynchronize());
            }
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            const double tot = (double)rpw * W * sms;
            printf("ldgsts row %u B  W %2d depth 8 : %7.3f ms  %7.1f GB/s\n", rb, W, ms, tot * rb / ms / 1e6);
        }
    }
    return 0;
}

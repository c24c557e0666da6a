// microbench.cu -- per-SM throughput of the instructions the build/apply kernels lean on (sm_100a).
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o microbench microbench.cu
// Each test runs one CTA of 1024 threads per SM for ITER iterations of an unrolled body and reports
// SM cycles per warp-instruction (clock64 around the loop, max over warps of CTA 0..).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

constexpr int ITER = 2048;
constexpr int UNROLL = 8;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// mode 0: red.shared.add 1 (ATOMS.POPC.INC), conflict-free (bank == lane), address varies per iteration
// mode 1: same but 2-way bank conflict
// mode 2: red.shared.add v (v != 1)
// mode 3: lds (conflict-free)
// mode 4: sts
// mode 5: atom.shared.add with return
// mode 6: red, all lanes of a warp random banks (pseudo random rows, bank = hash)
template <int MODE>
__global__ void __launch_bounds__(1024, 1) k_lsu(long long *out, int nwarps_active) {
    extern __shared__ unsigned int sm[];
    for (int i = threadIdx.x; i < 32 * 1024; i += blockDim.x) sm[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp >= nwarps_active) return;
    uint32_t base = smem_u32(sm);
    uint32_t x = threadIdx.x * 2654435761u + 12345u;
    uint32_t acc = 0;
    long long t0 = clock64();
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            x = x * 1664525u + 1013904223u;
            uint32_t row = (x >> 20) & 1023u;   // 1024 rows of 32 words
            uint32_t a;
            if (MODE == 1) a = base + row * 128u + ((lane >> 1) << 2) + ((lane & 1) << 16);  // pairs share a bank
            else if (MODE == 6) a = base + ((x >> 8) & 0x7FFFu) * 4u;
            else a = base + row * 128u + lane * 4u;
            if (MODE == 0 || MODE == 1 || MODE == 6) asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(a) : "memory");
            if (MODE == 2) asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a), "r"(x | 2u) : "memory");
            if (MODE == 3) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); acc += v; }
            if (MODE == 4) asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(x) : "memory");
            if (MODE == 5) { uint32_t v; asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(v) : "r"(a) : "memory"); acc += v; }
        }
    }
    long long t1 = clock64();
    if (lane == 0) out[blockIdx.x * 32 + warp] = t1 - t0;
    if (acc == 0x12345678u) out[0] = acc;
}

// Pure LSU rate: 8 loop-invariant, conflict-free addresses per thread; nothing else in the loop.
// mode 0: red.add 1 (POPC.INC); 1: red.add reg value; 2: ld.shared.u32; 3: ld.shared.u8;
// 4: red.add 1, 2 lanes per bank pair -> 2-way conflict; 5: red.add reg, 2-way conflict
template <int MODE>
__global__ void __launch_bounds__(1024, 1) k_lsu2(long long *out, int nwarps_active, uint32_t val) {
    extern __shared__ unsigned int sm[];
    for (int i = threadIdx.x; i < 32 * 1024; i += blockDim.x) sm[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp >= nwarps_active) return;
    uint32_t a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint32_t row = (warp * 8 + i) * 37u % 1000u;
        a[i] = smem_u32(sm) + row * 128u + ((MODE == 4 || MODE == 5) ? ((lane >> 1) << 2) + ((lane & 1) << 16) : lane * 4u);
    }
    uint32_t acc = 0, v = val + lane;
    long long t0 = clock64();
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0 || MODE == 4) asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(a[i]) : "memory");
            if (MODE == 1 || MODE == 5) asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a[i]), "r"(v) : "memory");
            if (MODE == 2) { uint32_t x; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(x) : "r"(a[i]) : "memory"); acc ^= x; }
            if (MODE == 3) { uint32_t x; asm volatile("ld.shared.u8 %0, [%1+1];" : "=r"(x) : "r"(a[i]) : "memory"); acc ^= x; }
        }
    }
    long long t1 = clock64();
    if (lane == 0) out[blockIdx.x * 32 + warp] = t1 - t0;
    if (acc == 0x12345678u) out[0] = acc;
}

// ALU / FMA pipe mixes: mode 0: LOP3 chain x4 independent; 1: IMAD x4; 2: 2 LOP3 + 2 IMAD; 3: IDP.4A x4;
// 4: 2 IDP + 2 LOP3; 5: PRMT x4; 6: 2 IDP + 2 IMAD
template <int MODE>
__global__ void __launch_bounds__(1024, 1) k_alu(long long *out, uint32_t seed) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t a = threadIdx.x + seed, b = a * 3u + 1u, c = a ^ 0x5555u, d = a + 77u;
    const uint32_t k1 = seed | 0x01010101u, k2 = seed ^ 0x0F0F0F0Fu;
    uint32_t r1 = k1 + threadIdx.x * 0x01000193u, r2 = k2 ^ (threadIdx.x * 0x9E3779B1u);
    asm volatile("" : "+r"(r1), "+r"(r2));
    long long t0 = clock64();
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            if (MODE == 0) { a = (a & k1) ^ b; b = (b | k2) ^ c; c = (c & k1) ^ d; d = (d | k2) ^ a; }
            if (MODE == 1) { a = a * k1 + b; b = b * k2 + c; c = c * k1 + d; d = d * k2 + a; }
            if (MODE == 2) { a = (a & k1) ^ b; b = b * k2 + c; c = (c & k1) ^ d; d = d * k2 + a; }
            if (MODE == 3) { a = __dp4a(a, k1, b); b = __dp4a(b, k2, c); c = __dp4a(c, k1, d); d = __dp4a(d, k2, a); }
            if (MODE == 4) { a = __dp4a(a, k1, b); b = (b | k2) ^ c; c = __dp4a(c, k1, d); d = (d | k2) ^ a; }
            if (MODE == 5) { a = __byte_perm(a, b, k1); b = __byte_perm(b, c, k2); c = __byte_perm(c, d, k1); d = __byte_perm(d, a, k2); }
            if (MODE == 6) { a = __dp4a(a, k1, b); b = b * k2 + c; c = __dp4a(c, k1, d); d = d * k2 + a; }
            if (MODE == 7) { a = (a & r1) ^ b; b = (b | r2) ^ c; c = (c & r1) ^ d; d = (d | r2) ^ a; }
            if (MODE == 8) { a = a + b + r1; b = b + c + r2; c = c + d + r1; d = d + a + r2; }
            if (MODE == 9) { a = __funnelshift_r(a, b, r1); b = __funnelshift_r(b, c, r2); c = __funnelshift_r(c, d, r1); d = __funnelshift_r(d, a, r2); }
            if (MODE == 10) { a = (b & 1) ? a : r1; b = (c & 1) ? b : r2; c = (d & 1) ? c : r1; d = (a & 1) ? d : r2; }
            if (MODE == 11) { a = (a & 0x0F0F0F0Fu) ^ b; b = (b | 0x80808080u) ^ c; c = (c & 0x07070707u) ^ d; d = (d | 0x01010101u) ^ a; }
            if (MODE == 12) { a = a + 0x55555555u; b = b + 0x01010101u; c = c + 0x33333333u; d = d + 0x0F0F0F0Fu; a ^= d; }
            if (MODE == 13) { a = (a & r1) ^ r2; b = (b | r2) ^ r1; c = (c & r1) ^ r2; d = (d | r2) ^ r1; }
            if (MODE == 14) { a = a & r1; b = b | r2; c = c ^ r1; d = d & r2; a |= 1u; }
            if (MODE == 15) { a = (a << 2) + b; b = (b << 3) + c; c = (c << 2) + d; d = (d << 3) + a; }
            if (MODE == 16) { a = __popc(a) + b; b = __popc(b) + c; c = __popc(c) + d; d = __popc(d) + a; }
            if (MODE == 17) { a = a >> (r1 & 7); b = b << (r2 & 7); c = (c >> 1) | r1; d = (d << 1) | r2; a |= c; b |= d; }
        }
    }
    long long t1 = clock64();
    if (lane == 0) out[blockIdx.x * 32 + warp] = t1 - t0;
    if ((a ^ b ^ c ^ d) == 0x12345678u) out[0] = a;
}

// Skeleton of the build inner loop: per "word": NA alu ops, NF fma ops, 8 ATOMS (conflict free), 3 LDS.
template <int NA, int NF>
__global__ void __launch_bounds__(1024, 1) k_mix(long long *out, uint32_t seed) {
    extern __shared__ unsigned int sm[];
    for (int i = threadIdx.x; i < 32 * 1024; i += blockDim.x) sm[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t base = smem_u32(sm) + lane * 4u;
    uint32_t a = threadIdx.x + seed, b = a * 3u + 1u;
    const uint32_t k1 = seed | 0x01010101u, k2 = (seed & 0xFFu) | 0x80u;
    long long t0 = clock64();
    for (int it = 0; it < ITER / 4; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            uint32_t v0, v1, v2;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v0) : "r"(base + ((a >> 3) & 0x3F80u)) : "memory");
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v1) : "r"(base + ((b >> 3) & 0x3F80u)) : "memory");
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v2) : "r"(base + ((a >> 9) & 0x3F80u)) : "memory");
            a += v0 + v1; b ^= v2;
#pragma unroll
            for (int i = 0; i < NA; ++i) { a = (a & k1) ^ b; b = __byte_perm(b, a, 0x3210 + i); }
#pragma unroll
            for (int i = 0; i < NF; ++i) { a = a * k2 + b; b = b * k2 + a; }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                uint32_t row = __byte_perm(i & 1 ? a : b, 0, 0x4440 + (i >> 1)) & 0xFFu;
                asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(base + row * 128u) : "memory");
            }
        }
    }
    long long t1 = clock64();
    if (lane == 0) out[blockIdx.x * 32 + warp] = t1 - t0;
    if ((a ^ b) == 0x12345678u) out[0] = a;
}

static long long *d_out;
static long long h_out[148 * 32];

template <class F> int run(const char *name, F launch, int nwarps, double warp_instr_per_iter) {
    CK(cudaMemset(d_out, 0, sizeof(h_out)));
    launch();
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h_out, d_out, sizeof(h_out), cudaMemcpyDeviceToHost));
    long long mx = 0;
    for (int i = 0; i < nwarps; ++i) mx = h_out[i] > mx ? h_out[i] : mx;  // CTA 0
    const double total_warp_instr = (double)nwarps * warp_instr_per_iter;
    printf("%-44s warps=%2d  cycles=%9lld  SM-cycles per warp-instr = %.3f\n", name, nwarps, mx, mx / total_warp_instr);
    return 0;
}

int main() {
    CK(cudaMalloc(&d_out, sizeof(h_out)));
    const int smem = 128 * 1024 + 64 * 1024;
#define LSU(MODE, NAME)                                                                              \
    CK(cudaFuncSetAttribute(k_lsu<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));        \
    for (int nw : {4, 8, 16, 32})                                                                    \
        run(NAME, [&] { k_lsu<MODE><<<148, 1024, smem>>>(d_out, nw); }, nw, (double)ITER * UNROLL);
    LSU(0, "red.shared.add 1 (POPC.INC) conflict-free");
    LSU(1, "red.shared.add 1 2-way conflict");
    LSU(2, "red.shared.add v conflict-free");
    LSU(6, "red.shared.add 1 random banks");
    LSU(5, "atom.shared.add (return) conflict-free");
    LSU(3, "ld.shared.u32 conflict-free");
    LSU(4, "st.shared.u32 conflict-free");
#define LSU2(MODE, NAME)                                                                            \
    CK(cudaFuncSetAttribute(k_lsu2<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));       \
    for (int nw : {8, 16, 32})                                                                       \
        run(NAME, [&] { k_lsu2<MODE><<<148, 1024, smem>>>(d_out, nw, 3u); }, nw, (double)ITER * 8);
    LSU2(0, "PURE red.add 1 (POPC.INC)");
    LSU2(1, "PURE red.add reg");
    LSU2(4, "PURE red.add 1 2-way conflict");
    LSU2(5, "PURE red.add reg 2-way conflict");
    LSU2(2, "PURE ld.shared.u32");
    LSU2(3, "PURE ld.shared.u8");
#define ALU(MODE, NAME) run(NAME, [&] { k_alu<MODE><<<148, 1024>>>(d_out, 0x9E3779B9u); }, 32, (double)ITER * UNROLL * 4);
    ALU(0, "LOP3 x4");
    ALU(1, "IMAD x4");
    ALU(2, "2 LOP3 + 2 IMAD");
    ALU(3, "IDP.4A x4");
    ALU(4, "2 IDP.4A + 2 LOP3");
    ALU(6, "2 IDP.4A + 2 IMAD");
    ALU(5, "PRMT x4");
    ALU(7, "LOP3 x4 reg operands");
    ALU(8, "IADD3 x4");
    ALU(9, "SHF x4");
    ALU(10, "SEL x4 (+4 LOP/ISETP)");
    ALU(11, "LOP3 imm x4");
    ALU(12, "IADD imm x4 + 1 LOP");
    ALU(13, "LOP3 reg x4, independent chains");
    ALU(14, "LOP 2-input x4 + 1");
    ALU(15, "LEA x4");
    ALU(16, "POPC+IADD x4");
    ALU(17, "shifts (6 ops)");
#define MIX(NA, NF)                                                                                   \
    CK(cudaFuncSetAttribute(k_mix<NA, NF>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));       \
    run("mix NA=" #NA " NF=" #NF " (2NA alu + 2NF fma + 8 red + 3 lds)/word",                        \
        [&] { k_mix<NA, NF><<<148, 1024, smem>>>(d_out, 0x9E3779B9u); }, 32, (double)ITER);
    MIX(0, 0);
    MIX(10, 0);
    MIX(10, 10);
    MIX(20, 10);
    MIX(20, 20);
    MIX(30, 15);
    printf("(mix rows: cycles per WORD-iteration per warp-slot; x32 warps -> SM cycles per warp-word)\n");
    return 0;
}

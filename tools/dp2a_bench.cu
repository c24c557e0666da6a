// dp2a_bench.cu -- is IDP.2A (16-bit x 8-bit dot product) issued at the rate of IDP.4A on sm_100a?
// One CTA of 1024 threads per SM, 4 independent chains per thread; SM cycles per warp instruction.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int ITER = 4096, UNROLL = 8;
template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(long long *out, uint32_t seed) {
    uint32_t a = threadIdx.x + seed, b = a * 3u + 1u, c = a ^ 0x5555u, d = a + 77u;
    const uint32_t k1 = seed | 0x01010101u, k2 = seed ^ 0x0F0F0F0Fu;
    long long t0 = clock64();
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            if (MODE == 0) { a = __dp4a(a, k1, b); b = __dp4a(b, k2, c); c = __dp4a(c, k1, d); d = __dp4a(d, k2, a); }
            if (MODE == 1) { a = __dp2a_lo(k1, a, b); b = __dp2a_hi(k2, b, c); c = __dp2a_lo(k1, c, d); d = __dp2a_hi(k2, d, a); }
            if (MODE == 2) { a = __dp2a_lo(k1, a, b); b = __dp4a(b, k2, c); c = __dp2a_hi(k1, c, d); d = __dp4a(d, k2, a); }
            if (MODE == 3) { a = __dp2a_lo(k1, a, b); b = (b | k2) ^ c; c = __dp2a_hi(k1, c, d); d = (d | k2) ^ a; }
        }
    }
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) out[blockIdx.x * 32 + (threadIdx.x >> 5)] = t1 - t0;
    if ((a ^ b ^ c ^ d) == 0x12345678u) out[0] = a;
}
int main() {
    long long *d, h[32];
    cudaMalloc(&d, 148 * 32 * 8);
    const char *names[] = {"IDP.4A x4", "IDP.2A lo/hi x4", "2 IDP.2A + 2 IDP.4A", "2 IDP.2A + 2 LOP3"};
    for (int m = 0; m < 4; ++m) {
        if (m == 0) k<0><<<148, 1024>>>(d, 0x9E3779B9u);
        if (m == 1) k<1><<<148, 1024>>>(d, 0x9E3779B9u);
        if (m == 2) k<2><<<148, 1024>>>(d, 0x9E3779B9u);
        if (m == 3) k<3><<<148, 1024>>>(d, 0x9E3779B9u);
        cudaDeviceSynchronize();
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int i = 0; i < 32; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("%-24s SM cycles per warp instruction = %.3f\n", names[m], mx / (32.0 * ITER * UNROLL * 4));
    }
    return 0;
}

// host_bitmap_bench.cpp -- can the host cores turn (seq, corrected) into a 1-bit-per-base mismatch map
// faster than PCIe moves the corrected reads?  g++ -O3 -std=c++17 -pthread -o host_bitmap_bench host_bitmap_bench.cpp
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <immintrin.h>

__attribute__((target("avx2"))) static void bitmap_avx2(const uint8_t *a, const uint8_t *b, size_t n32, uint32_t *out) {
    for (size_t i = 0; i < n32; ++i) {
        __m256i x = _mm256_loadu_si256((const __m256i *)(a + 32 * i));
        __m256i y = _mm256_loadu_si256((const __m256i *)(b + 32 * i));
        out[i] = ~(uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(x, y));
    }
}

int main(int argc, char **argv) {
    const size_t n = argc > 1 ? strtoull(argv[1], 0, 10) : 1500000000ull;
    const int T = argc > 2 ? atoi(argv[2]) : (int)std::thread::hardware_concurrency();
    uint8_t *a = (uint8_t *)aligned_alloc(64, n), *b = (uint8_t *)aligned_alloc(64, n);
    uint32_t *out = (uint32_t *)aligned_alloc(64, n / 8 + 64);
    memset(a, 'A', n); memset(b, 'A', n); memset(out, 0, n / 8 + 64);
    for (size_t i = 0; i < n; i += 97) b[i] = 'C';
    printf("threads %d avx2 %d\n", T, __builtin_cpu_supports("avx2"));
    for (int rep = 0; rep < 3; ++rep) {
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> pool;
        const size_t n32 = n / 32;
        for (int t = 0; t < T; ++t)
            pool.emplace_back([=] {
                const size_t lo = n32 * t / T, hi = n32 * (t + 1) / T;
                bitmap_avx2(a + 32 * lo, b + 32 * lo, hi - lo, out + lo);
            });
        for (auto &th : pool) th.join();
        double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        printf("bitmap of %zu bases: %.1f ms, %.1f GB/s read\n", n, s * 1e3, 2.0 * n / s / 1e9);
    }
    // plain copy for reference (memcpy of one array, all threads)
    for (int rep = 0; rep < 2; ++rep) {
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> pool;
        for (int t = 0; t < T; ++t)
            pool.emplace_back([=] { const size_t lo = n * t / T, hi = n * (t + 1) / T; memcpy(b + lo, a + lo, hi - lo); });
        for (auto &th : pool) th.join();
        double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        printf("memcpy %zu bytes: %.1f ms, %.1f GB/s read+write\n", n, s * 1e3, 2.0 * n / s / 1e9);
    }
    size_t c = 0;
    for (size_t i = 0; i < n / 32; ++i) c += __builtin_popcount(out[i]);
    printf("mismatches %zu\n", c);
    return 0;
}

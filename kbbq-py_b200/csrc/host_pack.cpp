// host_pack.cpp -- host-side packing for the PCIe-bound host-buffer entry point.
//
// kbbq_recalibrate_host spends its time on PCIe (4.5 GB up, 1.5 GB down per 10 M x 150 bp, the
// kernels take 2 ms).  The corrected reads are only ever compared with the reads
// (find_corrected_sites, kbbq/recalibrate.py:13-20), so one bit per base carries everything the build
// needs from them: the host cores reduce (seq, corrected) to a mismatch bit map while the copy engine
// moves seq and qual, the map (1/8 of the bytes) follows, and expand_corr_kernel (kbbq_b200.cu) turns
// it back into a byte array that differs from seq exactly where the corrected read did.
// Plain host C++ (no CUDA).  AVX2 when the CPU has it (32 bases per compare + movemask), portable
// 64-bit SWAR otherwise.
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include <thread>
#include <vector>

#include "../../include/kbbq_b200.h"

#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace {

#if defined(__x86_64__)
__attribute__((target("avx2"))) void bits_avx2(const uint8_t *a, const uint8_t *b, size_t words, uint32_t *out) {
    for (size_t i = 0; i < words; ++i) {
        const __m256i x = _mm256_loadu_si256((const __m256i *)(a + 32 * i));
        const __m256i y = _mm256_loadu_si256((const __m256i *)(b + 32 * i));
        out[i] = ~(uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(x, y));
    }
}
#endif

// bit j of out[i] = (a[32 i + j] != b[32 i + j])
void bits_portable(const uint8_t *a, const uint8_t *b, size_t words, uint32_t *out) {
    for (size_t i = 0; i < words; ++i) {
        uint32_t w = 0;
        for (int k = 0; k < 4; ++k) {
            uint64_t x, y;
            memcpy(&x, a + 32 * i + 8 * k, 8);
            memcpy(&y, b + 32 * i + 8 * k, 8);
            uint64_t d = x ^ y;
            // bit 7 of every non-zero byte, then gather the eight flags into one byte
            d = (((d & 0x7F7F7F7F7F7F7F7Full) + 0x7F7F7F7F7F7F7F7Full) | d) & 0x8080808080808080ull;
            w |= (uint32_t)(((d >> 7) * 0x0102040810204080ull) >> 56) << (8 * k);
        }
        out[i] = w;
    }
}

}  // namespace

extern "C" int kbbq_host_mismatch_bits(const uint8_t *seq, const uint8_t *corr, int64_t n, uint32_t *bits, int threads) {
    if (n < 0 || (n > 0 && (!seq || !corr || !bits))) return KBBQ_E_ARG;
    const size_t words = (size_t)n / 32;
    int T = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    if (T < 1) T = 1;
    if ((size_t)T > words / 4096 + 1) T = (int)(words / 4096 + 1);
#if defined(__x86_64__)
    const bool avx2 = __builtin_cpu_supports("avx2");
#else
    const bool avx2 = false;
#endif
    auto run = [&](int t) {
        const size_t lo = words * t / T, hi = words * (t + 1) / T;
#if defined(__x86_64__)
        if (avx2) { bits_avx2(seq + 32 * lo, corr + 32 * lo, hi - lo, bits + lo); return; }
#endif
        bits_portable(seq + 32 * lo, corr + 32 * lo, hi - lo, bits + lo);
    };
    if (T == 1) run(0);
    else {
        std::vector<std::thread> pool;
        pool.reserve(T);
        for (int t = 0; t < T; ++t) pool.emplace_back(run, t);
        for (auto &th : pool) th.join();
    }
    if ((size_t)n % 32) {  // last partial word
        uint32_t w = 0;
        for (size_t j = words * 32; j < (size_t)n; ++j) w |= (uint32_t)(seq[j] != corr[j]) << (j - words * 32);
        bits[words] = w;
    }
    return KBBQ_OK;
}

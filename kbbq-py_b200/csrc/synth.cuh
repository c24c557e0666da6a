// synth.cuh -- counter-based synthetic Illumina-shaped reads (bench / test input only).
//
// Every byte is a pure integer function of (seed, read index, cycle), so any read range can be
// regenerated on any GPU or on the host: kbbq-py_b200/kbbq/synth.py is the numpy twin and
// tests/test_synth.py checks that both produce identical bytes.  Shape: bases iid
// uniform ACGT with P(N) = 1/1024 (N gets Q2); quality = clip(mu_r - 8 (c/L)^2 + noise, 2, 41) with
// mu_r ~ 36 + Binomial(24, .5) - 12 (minus 2 for read 2) and noise ~ Binomial(16, .5) - 8
// (variance 4); 13/256 of the reads end in a Q2 tail of 1 .. L/4 bases; mismatches are iid
// Bernoulli(655/65536 ~ 1 %), the corrected base being one of the other three; reads are
// interleaved pairs (second = index & 1) and the read group is uniform over R per pair.
#pragma once
#include "common.cuh"

namespace kbbq {

__host__ __device__ inline unsigned long long mix64(unsigned long long x) {
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27; x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    return x;
}

__global__ void synth_kernel(unsigned long long seed, long long first, long long n, int L, int R,
                             uint8_t *seq, uint8_t *qual, uint8_t *corr, uint16_t *rg, uint8_t *second) {
    const long long total = n * L;
    const unsigned long long key = seed * 0x9E3779B97F4A7C15ull;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long rl = idx / L;
        const int i = (int)(idx - rl * L);
        const unsigned long long r = (unsigned long long)(first + rl);
        const unsigned long long hr = mix64(key + r * 2 + 1);
        const int sec = (int)(r & 1);
        if (i == 0) {
            const unsigned long long hp = mix64(key + (r >> 1) * 2);
            if (rg) rg[rl] = (uint16_t)((hp >> 8) % (unsigned long long)R);
            if (second) second[rl] = (uint8_t)sec;
        }
        const int mu = 36 + __popcll(hr & 0xFFFFFFull) - 12 - 2 * sec;
        const int l4 = L / 4 > 1 ? L / 4 : 1;
        const int tail = (((hr >> 24) & 0xFF) < 13) ? 1 + (int)(((hr >> 32) & 0xFFFF) % (unsigned)l4) : 0;
        const unsigned long long h = mix64(hr + (unsigned long long)(i + 1) * 0xD6E8FEB86659FD93ull);
        const int b = (int)(h & 3);
        const bool isn = ((h >> 2) & 0x3FF) == 0;
        const int noise = __popcll((h >> 12) & 0xFFFF) - 8;
        const bool iserr = ((h >> 28) & 0xFFFF) < 655;
        const int sh = 1 + (int)(((h >> 44) & 0xFFFF) % 3);
        const int decay = (int)((8ll * i * i + (long long)L * L / 2) / ((long long)L * L));
        int q = mu - decay + noise;
        q = q < 2 ? 2 : (q > 41 ? 41 : q);
        if (i >= L - tail || isn) q = 2;
        const char *acgt = "ACGT";
        const uint8_t s = isn ? 'N' : acgt[b];
        seq[idx] = s;
        qual[idx] = (uint8_t)q;
        corr[idx] = iserr ? acgt[(b + sh) & 3] : s;
    }
}

}  // namespace kbbq

// Host-only check of the ORB "Z" operand encoding of easysfm_b200/csrc/tc_layout.cuh (no GPU needed: nvcc builds it, the CPU runs it).
// Decodes the FP8 (E4M3) bytes the pack kernel / query writers emit and verifies, in exact integer arithmetic, that
//   sum(q_k * t_k) + sum(q'_s * t'_s) == kTcZ0i + 2^15 * hamming + c        for every column index c and a set of bit patterns,
// that every partial result stays below 2^24 (exact in the fp32 accumulator), and that the key decodes back to (hamming, c).
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "../../easysfm_b200/csrc/tc_layout.cuh"

using namespace esfm;

static double e4m3(uint8_t b) {      // OCP FP8 E4M3 (bias 7, no infinities): value of a byte
    const int s = b >> 7, e = (b >> 3) & 15, m = b & 7;
    double v;
    if (e == 0) v = m / 8.0 / 64.0;                  // subnormal: m/8 * 2^-6
    else v = (1.0 + m / 8.0) * (double)(1 << e) / 128.0;
    return s ? -v : v;
}

int main() {
    int bad = 0;
    // 1. digits and augmented blocks: q'.t' == 21 * 448^2 + c for every possible row index
    uint32_t qw[8];
    tcz_query_aug(qw);
    for (uint32_t c = 0; c < (uint32_t)kTcZMaxRows; ++c) {
        uint32_t tw[8];
        tcz_train_aug(c, tw);
        double sum = 0;
        for (int s = 0; s < 32; ++s) {
            const uint8_t qb = (qw[s >> 2] >> (8 * (s & 3))) & 255, tb = (tw[s >> 2] >> (8 * (s & 3))) & 255;
            sum += e4m3(qb) * e4m3(tb);
        }
        if (sum != 21.0 * 448 * 448 + c) { if (bad++ < 5) printf("aug mismatch c=%u sum=%.1f\n", c, sum); }
    }
    // 2. main operands: every nibble expands to +-256 (query) / -+64 (train), bit i in byte i
    for (uint32_t n = 0; n < 16; ++n) {
        const uint32_t q = tcz_expand4_q(n), t = tcz_expand4_t(n);
        for (int i = 0; i < 4; ++i) {
            const int bit = (n >> i) & 1;
            const double qv = e4m3((q >> (8 * i)) & 255), tv = e4m3((t >> (8 * i)) & 255);
            if (qv != (bit ? -256.0 : 256.0) || tv != (bit ? 64.0 : -64.0)) { if (bad++ < 5) printf("expand mismatch n=%u i=%d q=%.0f t=%.0f\n", n, i, qv, tv); }
        }
    }
    // 3. whole keys for a few hamming distances: product sign convention and decode
    for (int h = 0; h <= 256; h += (h < 8 || h > 248) ? 1 : 31) {
        const long long main_sum = (long long)h * (256 * 64) - (long long)(256 - h) * (256 * 64);       // differ: +2^14, agree: -2^14
        for (uint32_t c : {0u, 1u, 127u, 4095u, 4096u, 32767u}) {
            const long long z = main_sum + 21LL * 448 * 448 + c;
            if (z != (long long)kTcZ0i + ((long long)h << kTcZShift) + c || z < 0 || z >= (1LL << 24)) { if (bad++ < 5) printf("key mismatch h=%d c=%u z=%lld\n", h, c, z); }
            const float zf = (float)z;
            if ((long long)zf != z || (((int)zf - kTcZ0i) >> kTcZShift) != h || ((uint32_t)((int)zf - kTcZ0i) & (uint32_t)(kTcZMaxRows - 1)) != c) {
                if (bad++ < 5) printf("decode mismatch h=%d c=%u\n", h, c);
            }
        }
    }
    // 4. the largest partial sums stay exact: |main| <= 2^22, offset 21 * 448^2, digits < 2^15
    if (21LL * 448 * 448 + (1LL << 22) + 32767 >= (1LL << 24)) { ++bad; printf("range overflow\n"); }
    printf(bad ? "TCZ ENCODING FAILED (%d)\n" : "TCZ ENCODING OK\n", bad);
    return bad ? 1 : 0;
}

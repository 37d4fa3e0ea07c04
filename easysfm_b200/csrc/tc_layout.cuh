// tc_layout.cuh -- operand layout and PTX wrappers of the tcgen05 (5th-gen tensor core) L2 sweep.
//
// The SURF distance is ranked as  -1/2 d^2 = q.t - 1/2|q|^2 - 1/2|t|^2.
//   * q.t is a 3xTF32 split product: every fp32 operand x is stored as
//         hi = x rounded to TF32 (10 explicit mantissa bits),  lo = x - hi (exact in fp32; the MMA reads its top 19 bits)
//     and  q.t ~= lo.hi + hi.lo + hi.hi  (24 MMAs of K = 8) accumulates in fp32 in tensor memory (dropped term ~2^-22 |q||t|).
//   * the two half norms are added by ONE more K = 8 MMA over "augmented" columns.  A half norm h is split THREE ways,
//         h = h_h + h_m + h_l,   h_h = tf32(h), h_m = tf32(h - h_h), h_l = h - h_h - h_m   (every part exactly a TF32 number),
//         q' = ( 1,    1,    1,   hq_h | hq_m, hq_l, 0, 0)
//         t' = (-ht_h, -ht_m, -ht_l, -1 |  -1,   -1,  0, 0)
//     so q'.t' = -(ht + hq) with every product exact: 25 MMAs per 128 x 128 tile instead of the 27 of a hi/lo-split aug column.
//
// HBM / shared-memory image ("TC bank"), per 128-row tile (kTcTileBytes = 69632 contiguous bytes = one cp.async.bulk):
//   main  16 x 4096 B : per group of 8 rows [hi k0..31][hi k32..63][lo k0..31][lo k32..63], each a 1024-byte SWIZZLE_128B
//                  K-major atom (8 rows x 128 B; 16-byte chunk c of row r stored at chunk c ^ r) -- what tcgen05.mma's
//                  shared-memory descriptor calls layout_type 2, stride-byte-offset 4096 between 8-row groups;
//   aug   16 x  256 B : per group of 8 rows the 8 augmented columns (train role) in the no-swizzle K-major canonical form
//                  [k-chunk 2][row 8][16 B]  (leading-byte-offset 128, stride-byte-offset 256).
// The query tile's operand is built on the fly by the sweep (fp32 rows -> hi/lo in tensor memory, aug block in shared memory).
// Frames are padded to 128 rows; pad rows are zero with hq (resp. ht) = 1e30, so every accumulator involving them is
// <= -1e30 and can never be selected: no index masking in the kernel.
#pragma once
#include <cmath>
#include <cstring>

#include "esfm_internal.cuh"

namespace esfm {

constexpr int kTcGroupBytes = 4096;      // main image of one 8-row group
constexpr int kTcAugGroupBytes = 256;    // augmented block of one 8-row group
constexpr int kTcMainBytes = 16 * kTcGroupBytes;              // 65536: main image of a 128-row tile
constexpr int kTcAugBytes = 16 * kTcAugGroupBytes;            // 4096: augmented image of a 128-row tile
constexpr int kTcTileBytes = kTcMainBytes + kTcAugBytes;      // 69632 bytes per operand tile, contiguous in HBM
constexpr float kTcPadNorm = 1e30f;

__host__ __device__ __forceinline__ float tc_tf32_hi(float x) {
#ifdef __CUDA_ARCH__
    const uint32_t b = __float_as_uint(x);
    return __uint_as_float((b + 0x1000u) & 0xffffe000u);
#else
    uint32_t b;
    memcpy(&b, &x, 4);
    b = (b + 0x1000u) & 0xffffe000u;
    float r;
    memcpy(&r, &b, 4);
    return r;
#endif
}

// byte offset of element (row rr in [0,8), k in [0,32)) inside one 1024-byte SWIZZLE_128B atom
__host__ __device__ __forceinline__ int tc_sw128_off(int rr, int k) { return rr * 128 + ((((k >> 2) ^ rr) & 7) << 4) + ((k & 3) << 2); }

// byte offset of augmented column j in [0,8) of row rr inside one 256-byte no-swizzle block
__host__ __device__ __forceinline__ int tc_aug_off(int rr, int j) { return (j >> 2) * 128 + rr * 16 + ((j & 3) << 2); }

// three-way TF32 split of a half norm: h == hh + hm + hl exactly, each part a TF32 number
__host__ __device__ __forceinline__ void tc_split3(float h, float& hh, float& hm, float& hl) {
    hh = tc_tf32_hi(h);
    const float r = h - hh;
    hm = tc_tf32_hi(r);
    hl = r - hm;
}

// Host reference of the packing (used by the probe); the device pack kernel in bank.cu computes exactly the same bytes.
// `tile` points at the image of the 128-row tile that holds row r (r = row inside the tile).
inline void tc_pack_row_host(const float* x, bool valid, bool query_role, uint8_t* tile, int r) {
    const int g = r >> 3, rr = r & 7;
    float s = 0.f;
    for (int k = 0; k < kDim; ++k) s = fmaf(x[k], x[k], s);
    const float h = valid ? 0.5f * s : kTcPadNorm;
    for (int k = 0; k < kDim; ++k) {
        const float v = valid ? x[k] : 0.f;
        const float hi = tc_tf32_hi(v), lo = v - hi;
        const int off = g * kTcGroupBytes + (k >> 5) * 1024 + tc_sw128_off(rr, k & 31);
        memcpy(tile + off, &hi, 4);
        memcpy(tile + off + 2048, &lo, 4);
    }
    float hh, hm, hl;
    tc_split3(h, hh, hm, hl);
    float a[8] = {1.f, 1.f, 1.f, hh, hm, hl, 0.f, 0.f};
    if (!query_role) { a[0] = -hh; a[1] = -hm; a[2] = -hl; a[3] = -1.f; a[4] = -1.f; a[5] = -1.f; }
    for (int j = 0; j < 8; ++j) memcpy(tile + kTcMainBytes + g * kTcAugGroupBytes + tc_aug_off(rr, j), &a[j], 4);
}

// ---- SURF on kind::f16: the two-term 16-bit split ("H" layout) -----------------------------------------------------------
// The same idea as 3xTF32 at half the tensor-pipe time: every fp32 operand x is stored as
//     a = fp16(x)  (11 significant bits, round to nearest),   b = fp16(x - a)  (11 more; for |x| < 0.25 the residual is an fp16
//     SUBNORMAL, quantum 2^-24: the tensor core honours them -- csrc/microbench/h16_probe.cu -- so |x - a - b| <= 2^-25)
// and  q.t ~= b_q.a_t + a_q.b_t + a_q.a_t  = 3 terms x 4 MMAs of K = 16 (kind::f16, fp32 accumulation; a 16 x 16-bit k-step is
// the same 32 bytes of a SWIZZLE_128B row as a TF32 K = 8 step).  Dropped term b.b <= 2^-24 |q||t|; measured on unit-norm
// SURF-like rows: max |error| 4.1e-7 on 1/2 d^2, rms 1e-7 -- the same as 3xTF32 (3.6e-7).  (kind::f16 with A = fp16 and B = bf16
// in ONE instruction is an illegal instruction on sm_100a: measured; both operands must share the format.)  The half norms are
// added by the SAME exact K = 8 kind::tf32 MMA over the augmented columns as in the 3xTF32 layout (the kind is per instruction).
// Image per 128-row tile (kTchTileBytes = 36864, the geometry of the FP8 images): 16 groups x 2048 B = [a k0..63][b k0..63],
// each ONE 1024-byte SWIZZLE_128B K-major atom (8 rows x 128 B = 64 16-bit values), then the 16 x 256 B augmented blocks.
constexpr int kTchGroupBytes = 2048;
constexpr int kTchMainBytes = 16 * kTchGroupBytes;            // 32768
constexpr int kTchTileBytes = kTchMainBytes + kTcAugBytes;    // 36864
// fp16 / bf16 codes of a float, round to nearest even (host + device, no cuda_fp16.h types in the interfaces)
__host__ __device__ __forceinline__ uint32_t tch_f2h(float x) {
#ifdef __CUDA_ARCH__
    unsigned short h;
    asm("cvt.rn.f16.f32 %0, %1;" : "=h"(h) : "f"(x));
    return h;
#else
    uint32_t b;
    memcpy(&b, &x, 4);
    const uint32_t sign = (b >> 16) & 0x8000u;
    const int e = (int)((b >> 23) & 0xff) - 127 + 15;
    uint32_t m = b & 0x7fffffu;
    if (((b >> 23) & 0xff) == 0xff) return sign | 0x7c00u | (m ? 0x200u : 0u);
    if (e >= 31) return sign | 0x7c00u;
    if (e <= 0) {
        if (e < -10) return sign;
        m |= 0x800000u;
        const int sh = 14 - e;                       // 14 .. 24
        uint32_t r = m >> sh;
        const uint32_t rem = m & ((1u << sh) - 1u), half = 1u << (sh - 1);
        if (rem > half || (rem == half && (r & 1u))) ++r;
        return sign | r;
    }
    uint32_t r = ((uint32_t)e << 10) | (m >> 13);
    const uint32_t rem = m & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (r & 1u))) ++r;
    return sign | r;
#endif
}
__host__ __device__ __forceinline__ float tch_h2f(uint32_t h) {
#ifdef __CUDA_ARCH__
    float f;
    asm("cvt.f32.f16 %0, %1;" : "=f"(f) : "h"((unsigned short)h));
    return f;
#else
    const uint32_t sign = (h & 0x8000u) << 16;
    uint32_t e = (h >> 10) & 0x1fu, m = h & 0x3ffu, b;
    if (e == 0) {
        if (m == 0) b = sign;
        else {
            int sh = 0;
            while (!(m & 0x400u)) { m <<= 1; ++sh; }
            b = sign | ((uint32_t)(127 - 15 + 1 - sh) << 23) | ((m & 0x3ffu) << 13);
        }
    } else if (e == 31) b = sign | 0x7f800000u | (m << 13);
    else b = sign | ((e + 112u) << 23) | (m << 13);
    float f;
    memcpy(&f, &b, 4);
    return f;
#endif
}
__host__ __device__ __forceinline__ uint32_t tch_f2bf(float x) {
    uint32_t b;
#ifdef __CUDA_ARCH__
    b = __float_as_uint(x);
#else
    memcpy(&b, &x, 4);
#endif
    return (b + 0x7fffu + ((b >> 16) & 1u)) >> 16;            // finite inputs only
}
// (a, b) codes of x: a = fp16(x), b = fp16(x - a)
__host__ __device__ __forceinline__ void tch_split(float x, uint32_t& a, uint32_t& b) {
    a = tch_f2h(x);
    b = tch_f2h(x - tch_h2f(a));
}
// byte offset of 16-bit element k in [0,64) of row rr in [0,8) inside one 1024-byte SWIZZLE_128B atom
__host__ __device__ __forceinline__ int tch_sw128_off(int rr, int k) { return rr * 128 + ((((k >> 3) ^ rr) & 7) << 4) + ((k & 7) << 1); }

// Host reference of the H packing (probe + tests); bank.cu's pack kernel writes the same bytes.  Train role only: the query
// operand is built by the sweep's writer warps in tensor memory.
inline void tch_pack_row_host(const float* x, bool valid, uint8_t* tile, int r) {
    const int g = r >> 3, rr = r & 7;
    float s = 0.f;
    for (int k = 0; k < kDim; ++k) s = fmaf(x[k], x[k], s);
    const float h = valid ? 0.5f * s : kTcPadNorm;
    for (int k = 0; k < kDim; ++k) {
        uint32_t a, b;
        tch_split(valid ? x[k] : 0.f, a, b);
        const uint16_t a16 = (uint16_t)a, b16 = (uint16_t)b;
        const int off = g * kTchGroupBytes + tch_sw128_off(rr, k);
        memcpy(tile + off, &a16, 2);
        memcpy(tile + off + 1024, &b16, 2);
    }
    float hh, hm, hl;
    tc_split3(h, hh, hm, hl);
    const float a[8] = {-hh, -hm, -hl, -1.f, -1.f, -1.f, 0.f, 0.f};
    for (int j = 0; j < 8; ++j) memcpy(tile + kTchMainBytes + g * kTcAugGroupBytes + tc_aug_off(rr, j), &a[j], 4);
}

// ---- ORB / Hamming on the tensor cores: 256 bits -> 256 FP8 (E4M3) values +-1 ----------------------------------------
// bit 0 -> +1.0 (0x38), bit 1 -> -1.0 (0xB8): the dot product of two such rows is 256 - 2 * hamming, an exact small integer in
// the fp32 accumulator.  One more K = 32 MMA over augmented columns adds -256 and pushes pad rows out of reach:
//     q' = (qpad ? 448 : 0,  448,               16, 0 x 29)
//     t' = (-448,            tpad ? -448 : 0,  -16, 0 x 29)         =>  accumulator = -2 * hamming   (pads: < -200000)
// so the selection epilogue is the SURF one unchanged (it ranks "- 1/2 d^2").
// Image per 128-row tile (kTc8TileBytes = 36864): 16 groups x 2048 B = [k 0..127][k 128..255] SWIZZLE_128B atoms of 8 rows x
// 128 B, then 16 x 256 B augmented columns in the same no-swizzle form as the SURF image (K = 32 bytes = 2 chunks of 16).
constexpr int kTc8GroupBytes = 2048;
constexpr int kTc8MainBytes = 16 * kTc8GroupBytes;            // 32768
constexpr int kTc8TileBytes = kTc8MainBytes + kTcAugBytes;    // 36864
constexpr unsigned kFp8Pos448 = 0x7e, kFp8Neg448 = 0xfe, kFp8Pos16 = 0x58, kFp8Neg16 = 0xd8;

// 4 descriptor bits (low nibble of n) -> 4 FP8 bytes, bit i in byte i
__host__ __device__ __forceinline__ uint32_t tc8_expand4(uint32_t n) {
    return ((((n & 0xfu) * 0x00204081u) & 0x01010101u) << 7) | 0x38383838u;
}
// byte offset of element (row rr in [0,8), k in [0,128)) inside one 1024-byte SWIZZLE_128B atom of bytes
__host__ __device__ __forceinline__ int tc8_sw128_off(int rr, int k) { return rr * 128 + ((((k >> 4) ^ rr) & 7) << 4) + (k & 15); }

// ---- ORB "Z" encoding: the MMA itself delivers a packed (distance, column) key ------------------------------------------
// The fp32 accumulation of kind::f8f6f4 is exact for integers up to 2^23 (measured: csrc/microbench/f8_probe.cu,
// profiles/f8_probe_r1.txt), so the operands can be scaled until the accumulator holds
//     z = kTcZ0 + 2^15 * hamming + c,        c = the train row's index inside its frame (< 32768)
// an exact positive integer, smaller = nearer, ties = lower column, and distinct for every column: a query row's two
// nearest neighbours are then the two smallest of the floats it reads from tensor memory -- three FMNMX per element, no
// bound, no slow path, no tie logic.
//   main  q: bit 0 -> +256 (0x78), bit 1 -> -256 (0xF8);   t: bit 0 -> -64 (0xE8), bit 1 -> +64 (0x68)
//         sum = 2^14 * (2 h - 256) = 2^15 h - 2^22
//   aug   q' = (448 x 21, 1, 16, 256, 256, 0 x 7)      t' = (448 x 21, c & 15, (c >> 4) & 15, (c >> 8) & 15, 16 * (c >> 12), 0 x 7)
//         sum = 21 * 448^2 + c = 2^22 + 20480 + c
// Pad rows are all-zero operands (z = 0): the epilogue masks them by index (only the last tile of a frame has any).
// (kTcZShift = 15, kTcZMaxRows, kTcZ0i = 20480: esfm_internal.cuh -- finalize.cu decodes the keys)
constexpr float kTcZNone = 3.0e38f;                          // "no candidate" / masked
// 4 descriptor bits (low nibble of n) -> 4 FP8 bytes: query role +-256, train role -+64
__host__ __device__ __forceinline__ uint32_t tcz_expand4_q(uint32_t n) {
    return ((((n & 0xfu) * 0x00204081u) & 0x01010101u) << 7) | 0x78787878u;
}
__host__ __device__ __forceinline__ uint32_t tcz_expand4_t(uint32_t n) {
    return ((((n & 0xfu) * 0x00204081u) & 0x01010101u) << 7) ^ 0xe8e8e8e8u;
}
// E4M3 code of the integer v in [0, 15] (exact: at most 4 significant bits)
__host__ __device__ __forceinline__ uint32_t tcz_fp8_digit(uint32_t v) {
    if (v == 0) return 0u;
    const uint32_t e = v >= 8 ? 3u : (v >= 4 ? 2u : (v >= 2 ? 1u : 0u));
    return ((e + 7u) << 3) | (((v - (1u << e)) << 3) >> e);
}
// the 32 augmented bytes of a train row with frame index c, as 8 little-endian words
__host__ __device__ __forceinline__ void tcz_train_aug(uint32_t c, uint32_t (&w)[8]) {
    const uint32_t d3 = tcz_fp8_digit((c >> 12) & 7u);
    w[0] = w[1] = w[2] = w[3] = w[4] = 0x7e7e7e7eu;                       // slots 0..19
    w[5] = 0x7eu | (tcz_fp8_digit(c & 15u) << 8) | (tcz_fp8_digit((c >> 4) & 15u) << 16) | (tcz_fp8_digit((c >> 8) & 15u) << 24);   // 20..23
    w[6] = d3 ? d3 + (4u << 3) : 0u;                                      // slot 24: 16 * digit = exponent + 4
    w[7] = 0u;
}
// ... and of every query row: the multipliers
__host__ __device__ __forceinline__ void tcz_query_aug(uint32_t (&w)[8]) {
    w[0] = w[1] = w[2] = w[3] = w[4] = 0x7e7e7e7eu;
    w[5] = 0x7eu | (0x38u << 8) | (0x58u << 16) | (0x78u << 24);          // 448, 1, 16, 256
    w[6] = 0x78u;                                                         // 256
    w[7] = 0u;
}

#ifdef __CUDACC__
// ---- shared-memory matrix descriptors (tcgen05.mma operand A / B), K-major ---------------------------------------
// bits [0,14) start address >> 4, [16,30) leading byte offset >> 4, [32,46) stride byte offset >> 4,
// [46,48) descriptor version (1 on sm_100), [61,64) layout type (0 = no swizzle, 2 = SWIZZLE_128B).
__device__ __forceinline__ uint64_t tc_desc_sw128(uint32_t saddr, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3ffffu) >> 4) | (1ull << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint64_t tc_desc_nosw(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3ffffu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
// instruction descriptor, kind::tf32: D = F32 (bits 4-5 = 1), A = B = TF32 (bits 7-9, 10-12 = 2), both K-major,
// N >> 3 at bits 17-22, M >> 4 at bits 24-28.
__host__ __device__ constexpr uint32_t tc_idesc_tf32(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// kind::f16: D = F32, A / B format 0 = F16, 1 = BF16 (independent), both K-major
__host__ __device__ constexpr uint32_t tc_idesc_f16(int m, int n, int a_bf16, int b_bf16) {
    return (1u << 4) | ((uint32_t)a_bf16 << 7) | ((uint32_t)b_bf16 << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// kind::f8f6f4, A = B = E4M3, D = F16: sums of +-1 products are small integers, exact in fp16; the accumulator still takes one
// 32-bit tensor-memory column per element, tcgen05.ld.pack::16b returns two columns per register
__host__ __device__ constexpr uint32_t tc_idesc_e4m3_h(int m, int n) {
    return ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// kind::f8f6f4 with A = B = E4M3 (format code 0), D = F32
__host__ __device__ constexpr uint32_t tc_idesc_e4m3(int m, int n) {
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void tc_mma_f8(uint32_t d_tmem, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_mma_f8_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t desc_b, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(desc_b), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}

__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
// same, operand A read from tensor memory (lane = row m, one 32-bit column per k) instead of shared memory
__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t desc_b, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(desc_b), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
// kind::f16, operand A from tensor memory (lane = row m, TWO consecutive k per 32-bit column, lower k in the low half)
__device__ __forceinline__ void tc_mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t desc_b, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(desc_b), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
// one lane of the (fully active) warp is elected; the same lane every time, so MMAs and their commits share a thread
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
// all previously issued MMAs of this thread complete -> one arrival on the mbarrier (implies fence::before_thread_sync)
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// tensor-memory allocation: executed by ONE full warp; the base address lands in shared memory
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free(uint32_t addr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
// 32 lanes x 32 consecutive 32-bit columns: thread i of the warp receives lane (base lane + i), columns [col, col+32)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
// 32 lanes x 32 consecutive columns holding 16-bit values (fp16 accumulators), two columns packed per register: register i =
// column 2 i (low half) | column 2 i + 1 (high half)
__device__ __forceinline__ void tmem_ld16_pack(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.pack::16b.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
// registers -> tensor memory: thread i of the warp writes lane (base lane + i), columns [col, col+8)
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]),
                 "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
#endif  // __CUDACC__

}  // namespace esfm

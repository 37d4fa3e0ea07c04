// f8_probe.cu -- how exact is the fp32 accumulation of tcgen05.mma.kind::f8f6f4 at large magnitudes?
//
// The ORB sweep (sweep_l2_tc.cu, KIND = B256) accumulates +-1 products: sums <= 512, trivially exact.  A cheaper
// selection epilogue wants the MMA itself to deliver a packed (distance, column) key
//     z = Z0 + 2^14 * hamming + column            (exact integer < 2^24 in the fp32 accumulator)
// from operands scaled to q = +-128, t = -+64 (products +-2^13, partial sums up to 2^21) plus one augmented K = 32 step
// (11 slots of 448 * 448 for the offset, 4 slots that spell the column index digit by digit).  That only works if the
// tensor core adds FP8 products in full fp32 precision (23 bits), which is not documented.  This probe measures it:
// one 128 x 128 tile, K = 256 + 32, for several operand scales; accumulators read back and compared with exact integers.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o ../../bin/f8_probe f8_probe.cu
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../tc_layout.cuh"

using namespace esfm;

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e = (x);                                                                   \
        if (e != cudaSuccess) {                                                                \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);     \
            exit(2);                                                                           \
        }                                                                                      \
    } while (0)

__global__ void __launch_bounds__(192, 1) f8_probe_kernel(const uint8_t* __restrict__ q_img, const uint8_t* __restrict__ t_img,
                                                          float* __restrict__ out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* Qm = smem;                       // 32 KB main + 4 KB aug
    uint8_t* Tm = Qm + kTc8TileBytes;         // (36864 = 36 x 1024: the second image stays 1024-aligned)
    uint64_t* bars = reinterpret_cast<uint64_t*>(Tm + kTc8TileBytes);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_mbar_init();
    }
    if (warp == 5) tmem_alloc(tmem_slot, 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (warp == 4 && lane == 0) {
        mbar_arrive_expect_tx(&bars[0], 2 * kTc8TileBytes);
        bulk_g2s(Qm, q_img, kTc8TileBytes, &bars[0]);
        bulk_g2s(Tm, t_img, kTc8TileBytes, &bars[0]);
        mbar_wait(&bars[0], 0);
        tc_fence_after();
        const uint32_t idesc = tc_idesc_e4m3(128, 128);
        const uint64_t qd = tc_desc_sw128(smem_u32(Qm), kTc8GroupBytes), td = tc_desc_sw128(smem_u32(Tm), kTc8GroupBytes);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
            const uint64_t off = (uint64_t)(((ks >> 2) * 1024 + (ks & 3) * 32) >> 4);
            tc_mma_f8(tmem, qd + off, td + off, idesc, ks > 0);
        }
        tc_mma_f8(tmem, tc_desc_nosw(smem_u32(Qm) + kTc8MainBytes, 128, kTcAugGroupBytes),
                  tc_desc_nosw(smem_u32(Tm) + kTc8MainBytes, 128, kTcAugGroupBytes), idesc, true);
        tc_commit(&bars[1]);
    }
    if (warp < 4) {
        mbar_wait(&bars[1], 0);
        tc_fence_after();
        const int row = warp * 32 + lane;
        for (int c0 = 0; c0 < 128; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
            tmem_ld_wait();
            for (int j = 0; j < 32; ++j) out[(size_t)row * 128 + c0 + j] = __uint_as_float(v[j]);
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 5) tmem_free(tmem, 128);
}

// E4M3 encoding of +-2^e (e in [-6, 8]) and of the small integers 0..15 and 16 * (0..3)
static uint8_t fp8_pow2(int e, bool neg) { return (uint8_t)((neg ? 0x80 : 0) | ((e + 7) << 3)); }
static uint8_t fp8_int(int v) {      // exact for 0 <= v <= 15 (4 significant bits) and for 16 * c, c <= 7
    if (v == 0) return 0;
    int e = 0;
    while ((1 << (e + 1)) <= v) ++e;
    const int mant8 = (v - (1 << e)) * 8;
    if (mant8 % (1 << e)) { printf("fp8_int(%d) not representable\n", v); exit(2); }
    return (uint8_t)(((e + 7) << 3) | (mant8 >> e));
}

// one configuration: main operands q = +-2^qe, t = -+2^te; aug: n448 slots of 448 * 448, then the column index digits
static int run(int qe, int te, int n448, int col_base, bool with_cols) {
    std::vector<uint8_t> qi(kTc8TileBytes, 0), ti(kTc8TileBytes, 0);
    std::vector<uint32_t> qb(128 * 8), tb(128 * 8);
    srand(1234 + qe * 17 + te);
    for (auto& w : qb) w = ((uint32_t)rand() << 16) ^ (uint32_t)rand();
    for (auto& w : tb) w = ((uint32_t)rand() << 16) ^ (uint32_t)rand();
    for (int w = 0; w < 8; ++w) { tb[3 * 8 + w] = qb[5 * 8 + w]; tb[9 * 8 + w] = ~qb[7 * 8 + w]; }   // hamming 0 and 256
    for (int r = 0; r < 128; ++r) {
        for (int k = 0; k < 256; ++k) {
            const int off = (r >> 3) * kTc8GroupBytes + (k >> 7) * 1024 + tc8_sw128_off(r & 7, k & 127);
            const bool bq = (qb[r * 8 + (k >> 5)] >> (k & 31)) & 1, bt = (tb[r * 8 + (k >> 5)] >> (k & 31)) & 1;
            qi[off] = fp8_pow2(qe, bq);          // bit 0 -> +2^qe, bit 1 -> -2^qe
            ti[off] = fp8_pow2(te, !bt);         // bit 0 -> -2^te, bit 1 -> +2^te   => sum = 2^(qe+te) * (2 h - 256)
        }
        uint8_t qa[32] = {0}, ta[32] = {0};
        for (int s = 0; s < n448; ++s) { qa[s] = 0x7e; ta[s] = 0x7e; }
        if (with_cols) {
            const int col = col_base + r;        // (as a train row: its column index; the query side only carries the multipliers)
            qa[n448 + 0] = fp8_int(1);  ta[n448 + 0] = fp8_int(col & 15);
            qa[n448 + 1] = fp8_pow2(4, false);  ta[n448 + 1] = fp8_int((col >> 4) & 15);
            qa[n448 + 2] = fp8_pow2(8, false);  ta[n448 + 2] = fp8_int((col >> 8) & 15);
            qa[n448 + 3] = fp8_pow2(8, false);  ta[n448 + 3] = fp8_int(16 * ((col >> 12) & 7));
        }
        for (int j = 0; j < 32; ++j) {
            const int off = kTc8MainBytes + (r >> 3) * kTcAugGroupBytes + (j >> 4) * 128 + (r & 7) * 16 + (j & 15);
            qi[off] = qa[j];
            ti[off] = ta[j];
        }
    }
    uint8_t *dq, *dt;
    float* dout;
    CK(cudaMalloc(&dq, qi.size())); CK(cudaMalloc(&dt, ti.size())); CK(cudaMalloc(&dout, 128 * 128 * 4));
    CK(cudaMemcpy(dq, qi.data(), qi.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dt, ti.data(), ti.size(), cudaMemcpyHostToDevice));
    const size_t smem = 2 * kTc8TileBytes + 64;
    CK(cudaFuncSetAttribute(f8_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    f8_probe_kernel<<<1, 192, smem>>>(dq, dt, dout);
    CK(cudaDeviceSynchronize());
    std::vector<float> out(128 * 128);
    CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
    int bad = 0;
    double maxerr = 0, maxabs = 0;
    for (int r = 0; r < 128; ++r)
        for (int c = 0; c < 128; ++c) {
            int h = 0;
            for (int w = 0; w < 8; ++w) h += __builtin_popcount(qb[r * 8 + w] ^ tb[c * 8 + w]);
            const double want = std::ldexp(1.0, qe + te) * (2.0 * h - 256.0) + n448 * 200704.0 + (with_cols ? col_base + c : 0);
            const double got = out[r * 128 + c], err = std::fabs(want - got);
            if (std::fabs(want) > maxabs) maxabs = std::fabs(want);
            if (err > maxerr) maxerr = err;
            if (err != 0) {
                if (bad < 4) printf("    r=%d c=%d h=%d want %.1f got %.1f\n", r, c, h, want, got);
                ++bad;
            }
        }
    printf("q=+-2^%d t=-+2^%d offset slots=%2d column digits=%d (base %5d): max|value| %.0f, max|error| %.1f, inexact %d of 16384\n", qe, te, n448,
           (int)with_cols, col_base, maxabs, maxerr, bad);
    cudaFree(dq); cudaFree(dt); cudaFree(dout);
    return bad;
}

int main() {
    run(0, 0, 0, 0, false);            // the shipped encoding's main part: |sum| <= 256
    run(4, 4, 0, 0, false);            // 2^8 per unit: <= 2^16
    run(7, 6, 0, 0, false);            // 2^13 per unit: <= 2^21
    run(7, 6, 11, 0, false);           // + offset 11 * 448^2: <= 4.3M
    run(7, 6, 11, 0, true);            // + column digits, columns 0..127
    run(7, 6, 11, 16256, true);        // columns 16256..16383 (all four digits in use)
    run(8, 6, 0, 0, false);            // 2^14 per unit: <= 2^22
    run(8, 7, 0, 0, false);            // 2^15 per unit: <= 2^23
    run(8, 6, 21, 0, true);            // the "Z" encoding of tc_layout.cuh: z = 20480 + 2^15 h + column, columns 0..127
    run(8, 6, 21, 32640, true);        // ... columns 32640..32767: values up to 8.44M with unit resolution
    return 0;
}

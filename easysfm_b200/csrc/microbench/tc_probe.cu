// tc_probe.cu -- standalone probe for the tcgen05 (kind::tf32) building blocks of sweep_l2_tc.cu.
//
// 1. correctness: one 128-row query tile x N train rows, K = 64, 3xTF32 split (hi.hi + hi.lo + lo.hi) + ONE augmented
//    K = 8 MMA that adds the two half norms (three-way split, exact), accumulators read back from TMEM and compared
//    with float64 -1/2 d^2;
// 2. pacing: cycles per tcgen05.mma (M=128, N in {64,128,256}, K=8, both operands in shared memory),
//    which decides the train-stage width of the sweep (shared-memory operand reads vs the MMA floor).
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ../../bin/tc_probe tc_probe.cu
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

template <int N>
__global__ void __launch_bounds__(192, 1)
probe_kernel(const uint8_t* __restrict__ q_img, const uint8_t* __restrict__ t_img,     // tile images of tc_layout.cuh
             float* __restrict__ out, int terms, int reps, long long* __restrict__ cycles, const float* __restrict__ q_rows, int ts_mode) {
    extern __shared__ __align__(1024) uint8_t smem[];
    // Q: 128 rows main (16 groups x 4096) + aug (16 x 256); T: N rows likewise (groups of all tiles made contiguous)
    uint8_t* Qm = smem;                               // 64 KB
    uint8_t* Tm = Qm + 16 * kTcGroupBytes;            // N/8 * 4096
    uint8_t* Qa = Tm + (N / 8) * kTcGroupBytes;
    uint8_t* Ta = Qa + 16 * kTcAugGroupBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(Ta + (N / 8) * kTcAugGroupBytes);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_mbar_init();
    }
    if (warp == 5) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (ts_mode && warp < 4) {
        // operand A into tensor memory: lane = row, columns 384.. = hi(k 0..63), 448.. = lo(k 0..63)
        const int row = warp * 32 + lane;
        for (int k0 = 0; k0 < 64; k0 += 8) {
            uint32_t hi[8], lo[8];
            for (int j = 0; j < 8; ++j) {
                const float x = q_rows[row * 64 + k0 + j];
                const float h = tc_tf32_hi(x);
                hi[j] = __float_as_uint(h);
                lo[j] = __float_as_uint(x - h);
            }
            tmem_st8(tmem + ((uint32_t)(warp * 32) << 16) + 384 + k0, hi);
            tmem_st8(tmem + ((uint32_t)(warp * 32) << 16) + 448 + k0, lo);
        }
        tmem_st_wait();
        tc_fence_before();
    }
    if (ts_mode) { __syncthreads(); tc_fence_after(); }

    if (warp == 4 && lane == 0) {
        const uint32_t bytes = (16 + N / 8) * (kTcGroupBytes + kTcAugGroupBytes);
        mbar_arrive_expect_tx(&bars[0], bytes);
        bulk_g2s(Qm, q_img, kTcMainBytes, &bars[0]);
        bulk_g2s(Qa, q_img + kTcMainBytes, kTcAugBytes, &bars[0]);
        for (int r0 = 0; r0 < N; r0 += 128) {
            const int g = (N - r0 < 128 ? N - r0 : 128) / 8;      // groups of this tile that are used
            const uint8_t* img = t_img + (size_t)(r0 / 128) * kTcTileBytes;
            bulk_g2s(Tm + (r0 / 8) * kTcGroupBytes, img, g * kTcGroupBytes, &bars[0]);
            bulk_g2s(Ta + (r0 / 8) * kTcAugGroupBytes, img + kTcMainBytes, g * kTcAugGroupBytes, &bars[0]);
        }
        mbar_wait(&bars[0], 0);
        tc_fence_after();
        const uint32_t idesc = tc_idesc_tf32(128, N);
        const uint32_t qm = smem_u32(Qm), tm = smem_u32(Tm), qa = smem_u32(Qa), ta = smem_u32(Ta);
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            const uint32_t d = tmem + (uint32_t)((r & 1) * N) % 512u;
            bool first = true;
            for (int term = 0; term < terms; ++term) {
                // small terms first -- term 0: lo.hi   1: hi.lo   2: hi.hi     (A part, B part); a 1-term run is hi.hi only
                const int pa = (terms == 3 && term == 0) ? 1 : 0, pb = (terms == 3 && term == 1) ? 1 : 0;
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    const uint64_t da = tc_desc_sw128(qm + pa * 2048 + (ks >> 2) * 1024 + (ks & 3) * 32, kTcGroupBytes);
                    const uint64_t db = tc_desc_sw128(tm + pb * 2048 + (ks >> 2) * 1024 + (ks & 3) * 32, kTcGroupBytes);
                    if (ts_mode) tc_mma_tf32_ts(d, tmem + 384 + pa * 64 + ks * 8, db, idesc, !first);
                    else tc_mma_tf32(d, da, db, idesc, !first);
                    first = false;
                }
            }
            // the half norms: one K = 8 MMA over the three-way split augmented columns
            tc_mma_tf32(d, tc_desc_nosw(qa, 128, kTcAugGroupBytes), tc_desc_nosw(ta, 128, kTcAugGroupBytes), idesc, true);
        }
        tc_commit(&bars[1]);
        mbar_wait(&bars[1], 0);
        const long long t1 = clock64();
        if (cycles) *cycles = t1 - t0;
    }
    if (warp < 4) {
        mbar_wait(&bars[1], 0);
        tc_fence_after();
        const int row = warp * 32 + lane;
        const uint32_t dsel = ((reps - 1) & 1) * N % 512;
        for (int c0 = 0; c0 < N; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + dsel + c0, v);
            tmem_ld_wait();
            for (int j = 0; j < 32; ++j) out[(size_t)row * N + c0 + j] = __uint_as_float(v[j]);
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 5) tmem_free(tmem, 512);
}

static void fill_rows(std::vector<float>& x, int rows, unsigned seed, bool unit) {
    srand(seed);
    x.resize((size_t)rows * 64);
    for (int r = 0; r < rows; ++r) {
        double n = 0;
        for (int k = 0; k < 64; ++k) {
            const float v = (float)rand() / RAND_MAX - 0.5f;
            x[(size_t)r * 64 + k] = v;
            n += (double)v * v;
        }
        if (unit)
            for (int k = 0; k < 64; ++k) x[(size_t)r * 64 + k] = (float)(x[(size_t)r * 64 + k] / std::sqrt(n));
    }
}

template <int N>
static int run(int terms, int reps, bool check, int ts_mode = 0) {
    std::vector<float> q, t;
    fill_rows(q, 128, 1, true);
    fill_rows(t, N, 2, true);
    // make a few near-duplicates so small distances are exercised
    for (int k = 0; k < 64; ++k) t[(size_t)3 * 64 + k] = q[(size_t)5 * 64 + k];
    for (int k = 0; k < 64; ++k) t[(size_t)7 * 64 + k] = q[(size_t)9 * 64 + k] * (1.f + 1e-3f * (k & 1));
    const int t_tiles = (N + 127) / 128;
    std::vector<uint8_t> qm(kTcTileBytes), tm((size_t)t_tiles * kTcTileBytes);
    for (int r = 0; r < 128; ++r) tc_pack_row_host(q.data() + (size_t)r * 64, true, true, qm.data(), r);
    for (int r = 0; r < N; ++r)
        tc_pack_row_host(t.data() + (size_t)r * 64, true, false, tm.data() + (size_t)(r / 128) * kTcTileBytes, r % 128);
    uint8_t *dqm, *dtm;
    float* dout;
    float* dq;
    long long* dcyc;
    CK(cudaMalloc(&dq, q.size() * 4));
    CK(cudaMemcpy(dq, q.data(), q.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&dqm, qm.size())); CK(cudaMalloc(&dtm, tm.size()));
    CK(cudaMalloc(&dout, (size_t)128 * N * 4)); CK(cudaMalloc(&dcyc, 8));
    CK(cudaMemcpy(dqm, qm.data(), qm.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dtm, tm.data(), tm.size(), cudaMemcpyHostToDevice));
    const size_t smem = (16 + N / 8) * (kTcGroupBytes + kTcAugGroupBytes) + 64;
    CK(cudaFuncSetAttribute(probe_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    probe_kernel<N><<<1, 192, smem>>>(dqm, dtm, dout, terms, reps, dcyc, dq, ts_mode);
    CK(cudaDeviceSynchronize());
    long long cyc = 0;
    CK(cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost));
    std::vector<float> out((size_t)128 * N);
    CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
    int bad = 0;
    double maxerr = 0;
    if (check) {
        for (int r = 0; r < 128; ++r)
            for (int c = 0; c < N; ++c) {
                double d2 = 0;
                for (int k = 0; k < 64; ++k) {
                    const double df = (double)q[(size_t)r * 64 + k] - (double)t[(size_t)c * 64 + k];
                    d2 += df * df;
                }
                const double want = -0.5 * d2, got = out[(size_t)r * N + c];
                const double err = std::fabs(want - got);
                if (err > maxerr) maxerr = err;
                if (err > (terms == 3 ? 2e-5 : 2e-3)) {
                    if (bad < 8) printf("  mismatch r=%d c=%d want %.8f got %.8f\n", r, c, want, got);
                    ++bad;
                }
            }
    }
    const int mmas = reps * (terms * 8 + 1);
    printf("%s N=%3d terms=%d reps=%4d: %lld cycles, %.1f cycles/MMA (floor %d)%s max|err|=%.3g bad=%d\n", ts_mode ? "A-in-TMEM" : "A-in-smem", N, terms, reps, cyc,
           (double)cyc / mmas, N / 2, check ? "" : " [timing only]", maxerr, bad);
    cudaFree(dqm); cudaFree(dtm); cudaFree(dout); cudaFree(dcyc);
    return bad;
}

int main() {
    int bad = 0;
    bad += run<64>(1, 1, true);
    bad += run<64>(3, 1, true);
    bad += run<128>(3, 1, true);
    bad += run<256>(3, 1, true);
    run<64>(3, 200, false);
    run<128>(3, 200, false);
    run<256>(3, 100, false);
    bad += run<64>(3, 1, true, 1);
    bad += run<128>(3, 1, true, 1);
    run<64>(3, 200, false, 1);
    run<128>(3, 200, false, 1);
    printf(bad ? "PROBE FAILED\n" : "PROBE OK\n");
    return bad ? 1 : 0;
}

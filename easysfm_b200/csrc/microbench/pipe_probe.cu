// pipe_probe.cu -- what does the accumulator hand-shake of the tensor-core sweeps cost?  (sm_100a)
//
// The sweeps' "MMAs only" probes run at ~1000 (ORB, 9 MMAs) / ~1420 (SURF, 13 MMAs) cycles per 128 x 128 tile although the MMAs themselves
// take 64 cycles each back to back (h16_probe).  This probe isolates the loop: one issuer warp, W drain warps, S accumulator stages of N
// columns; per tile the issuer waits for the stage to be free, issues M tcgen05.mma (kind::f8f6f4, A from tensor memory, B from a fixed
// shared-memory tile), commits to the stage's "full" barrier; the drain warps wait for it, (optionally) tcgen05.ld their part, and arrive on
// the stage's "empty" barrier.  Variants switch single ingredients off.  Every CTA of the grid runs the same loop (grid = 1 or 148).
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o ../../bin/pipe_probe pipe_probe.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

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

struct Cfg {
    int tiles, mmas, stages, ncols, drain_warps;
    int do_ld;          // drain warps read their part of the accumulator
    int wait_empty;     // issuer waits for the stage to be free (0: free-running, commits still issued)
    int commit_every;   // commit after every tile (1) or only after the last one (0; implies no hand-shake)
    int whole_warp;     // issuer: whole warp walks the loop, one elected lane issues (as the sweeps do) / 0: a single thread
    int fence;          // tcgen05.fence::after_thread_sync before every tile's MMAs
    int commits;        // commits per tile (the sweeps issue two: accumulator full + shared-memory stage free)
};

template <int MMAS, int NCOLS>
__global__ void __launch_bounds__(640, 1) pipe_kernel(Cfg c, long long* __restrict__ cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* Tm = smem;                                  // one FP8 tile image worth of (garbage) operand bytes
    uint64_t* bars = reinterpret_cast<uint64_t*>(Tm + kTc8TileBytes);
    uint64_t* full = bars;                               // [8]
    uint64_t* empty = bars + 8;                          // [8]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < kTc8TileBytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(Tm)[i] = 0x38383838u;
    if (threadIdx.x == 0) {
        for (int s = 0; s < 8; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], c.drain_warps); }
        fence_mbar_init();
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 17) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t acol = 448;                           // A operand columns (garbage)
    const long long t0 = clock64();
    if (warp == 16) {
        constexpr uint32_t idesc = tc_idesc_e4m3(128, NCOLS);
        const uint64_t td = tc_desc_sw128(smem_u32(Tm), kTc8GroupBytes);
        if (c.whole_warp || lane == 0) {
            for (int g = 0; g < c.tiles; ++g) {
                const uint32_t s = g % c.stages, ph = (g / c.stages) & 1;
                if (c.wait_empty && c.commit_every) mbar_wait(&empty[s], ph ^ 1);
                if (c.fence) tc_fence_after();
                const uint32_t d = tmem + s * NCOLS;
                if (!c.whole_warp || elect_one()) {
#pragma unroll
                    for (int k = 0; k < MMAS; ++k)
                        tc_mma_f8_ts(d, acol + (k & 7) * 8, td + (uint64_t)((((k & 7) >> 2) * 1024 + (k & 3) * 32) >> 4), idesc, k > 0);
                }
                if (c.whole_warp) __syncwarp();
                if (c.commit_every) {
                    if (!c.whole_warp || elect_one()) { tc_commit(&full[s]); if (c.commits > 1) tc_commit(&full[6]); }
                }
                if (c.whole_warp) __syncwarp();
            }
            if (!c.whole_warp || elect_one()) tc_commit(&full[7]);       // "everything issued has completed"
        }
    } else if (warp < c.drain_warps) {
        const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
        const int part = warp >> 2;
        uint32_t acc = 0;
        if (c.commit_every && c.wait_empty) {
            for (int g = 0; g < c.tiles; ++g) {
                const uint32_t s = g % c.stages, ph = (g / c.stages) & 1;
                mbar_wait(&full[s], ph);
                tc_fence_after();
                if (c.do_ld) {
                    uint32_t v[32];
                    tmem_ld32(tmem + lane_addr + s * NCOLS + (part * 32) % NCOLS, v);
                    tmem_ld_wait();
                    acc ^= v[0] ^ v[31];
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_relaxed(&empty[s]);
            }
        }
        mbar_wait(&full[7], 0);
        if (acc == 0x12345678u) cycles[1] = acc;
    }
    tc_fence_before();
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
    if (warp == 17) tmem_free(tmem, 512);
}

static void run(const char* name, Cfg c, int grid) {
    long long* dcyc;
    CK(cudaMalloc(&dcyc, 16));
    const size_t smem = kTc8TileBytes + 256;
    void (*k)(Cfg, long long*) = c.ncols == 64 ? pipe_kernel<9, 64> : (c.mmas == 13 ? pipe_kernel<13, 128> : (c.mmas == 1 ? pipe_kernel<1, 128> : pipe_kernel<9, 128>));
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, 640, smem>>>(c, dcyc);
    CK(cudaDeviceSynchronize());
    long long cyc = 0;
    CK(cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost));
    printf("%-58s grid %3d: %4d tiles x %2d MMAs (N=%3d), %d stages, %2d drain warps: %7.1f cycles per tile (MMAs alone: %d)\n", name, grid, c.tiles, c.mmas, c.ncols,
           c.stages, c.drain_warps, (double)cyc / c.tiles, c.mmas * c.ncols / 2);
    cudaFree(dcyc);
}

int main(int argc, char** argv) {
    const int grid = argc > 1 ? atoi(argv[1]) : 1;
    const int T = 3000;
    //                                                       tiles mmas st  N  warps ld wait commit warp fence commits
    run("free-running, no fence, 1 thread",                 {T, 9, 3, 128, 16, 0, 0, 0, 0, 0, 1}, grid);
    run("free-running, no fence, whole warp + elect",       {T, 9, 3, 128, 16, 0, 0, 0, 1, 0, 1}, grid);
    run("free-running, fence, 1 thread",                    {T, 9, 3, 128, 16, 0, 0, 0, 0, 1, 1}, grid);
    run("free-running, fence, whole warp + elect",          {T, 9, 3, 128, 16, 0, 0, 0, 1, 1, 1}, grid);
    run("commit per tile, no wait, no fence, 1 thread",     {T, 9, 3, 128, 16, 0, 0, 1, 0, 0, 1}, grid);
    run("2 commits per tile, no wait, no fence, 1 thread",  {T, 9, 3, 128, 16, 0, 0, 1, 0, 0, 2}, grid);
    run("hand-shake + ld, no fence, 1 thread",              {T, 9, 3, 128, 16, 1, 1, 1, 0, 0, 1}, grid);
    run("hand-shake + ld, fence, 1 thread",                 {T, 9, 3, 128, 16, 1, 1, 1, 0, 1, 1}, grid);
    run("hand-shake + ld, fence, whole warp (the sweeps)",  {T, 9, 3, 128, 16, 1, 1, 1, 1, 1, 2}, grid);
    run("hand-shake + ld, no fence, 1 thread, 13 MMAs",     {T, 13, 3, 128, 16, 1, 1, 1, 0, 0, 1}, grid);
    run("hand-shake + ld, no fence, 1 thread, 1 MMA",       {T, 1, 3, 128, 16, 1, 1, 1, 0, 0, 1}, grid);
    run("hand-shake + ld, no fence, 1 thread, 1 stage",     {T, 9, 1, 128, 16, 1, 1, 1, 0, 0, 1}, grid);
    run("hand-shake + ld, no fence, 1 thread, 2 stages",    {T, 9, 2, 128, 16, 1, 1, 1, 0, 0, 1}, grid);
    return 0;
}

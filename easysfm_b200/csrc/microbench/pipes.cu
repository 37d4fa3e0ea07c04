// Pipe-rate microbenchmark for sm_100a (B200): measures per-SM issue rates of the
// instructions the matching kernels are built from, so that the rooflines in
// DESIGN.md / bench.py use MEASURED denominators (SURVEY.md §7 "Hard parts" item 4).
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu
// Run:   ./pipes            (prints one line per test: warp-instr/clk/SM and lane-ops/clk/SM)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int ITERS = 4096;

enum Op { FFMA, FFMA2, POPC, LOP3, IADD3, IMAD, FMNMX, VIMNMX, FSETP_SEL, REDUX, POPC_LOP3, POPC_LOP3x2, FFMA_LDS, LDS128_BCAST, LDS128_QBCAST, LDS128_FULL, FADD, POPC_IMAD, HAM8, HAM8_CSA };

template <int OP>
__global__ void __launch_bounds__(1024) kern(uint32_t* out, long long* cycles, uint32_t seed) {
    __shared__ __align__(16) float sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = (float)i;
    __syncthreads();
    uint32_t r[8], s[8];
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { r[i] = seed * (threadIdx.x + 1) + i * 7919u; f[i] = (float)(r[i] & 1023) * 1e-3f; s[i] = r[i] * 31u + 5u; }
    float fa = (float)(seed & 7) * 1e-4f + 1.0f, fb = 1e-7f;
    uint32_t ua = seed | 1u, ub = seed * 3u + 1u;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        if (OP == FFMA) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(f[i]) : "f"(fa), "f"(fb));
        } else if (OP == FADD) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < 8; ++i) asm volatile("add.rn.f32 %0, %1, %0;" : "+f"(f[i]) : "f"(fb));
        } else if (OP == FFMA2) {
            // 16 packed FMAs = 32 scalar FMAs
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    asm volatile("{ .reg .b64 a, b, c; mov.b64 a, {%2, %2}; mov.b64 b, {%3, %3}; mov.b64 c, {%0, %1};\n"
                                 "fma.rn.f32x2 c, a, b, c; mov.b64 {%0, %1}, c; }"
                                 : "+f"(f[2 * i]), "+f"(f[2 * i + 1]) : "f"(fa), "f"(fb));
                }
        } else if (OP == POPC) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < 8; ++i) asm volatile("popc.b32 %0, %0;" : "+r"(r[i]));
        } else if (OP == LOP3) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < 8; ++i) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(ua), "r"(ub));
        } else if (OP == IADD3) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < 8; ++i) asm volatile("add.s32 %0, %0, %1;" : "+r"(r[i]) : "r"(ua));
        } else if (OP == IMAD) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < 8; ++i) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(ua), "r"(ub));
        } else if (OP == FMNMX) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < 8; ++i) asm volatile("min.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(fa));
        } else if (OP == VIMNMX) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < 8; ++i) asm volatile("min.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(ua));
        } else if (OP == FSETP_SEL) {
            // compare + OR-accumulate into a predicate-like register (setp + selp = 2 instr)
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    asm volatile("{ .reg .pred p; setp.lt.f32 p, %1, %2; @p add.s32 %0, %0, 1; }" : "+r"(r[i]) : "f"(f[i]), "f"(fa));
        } else if (OP == REDUX) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < 8; ++i) asm volatile("redux.sync.min.u32 %0, %0, 0xffffffff;" : "+r"(r[i]));
        } else if (OP == POPC_LOP3) {
            // 1 POPC : 1 LOP3 (the naive Hamming mix) -- do they overlap?
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    asm volatile("popc.b32 %0, %0;" : "+r"(r[i]));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(s[i]) : "r"(ua), "r"(ub));
                }
        } else if (OP == POPC_LOP3x2) {
            // 1 POPC : 4 LOP3 (CSA-compressed Hamming mix)
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    asm volatile("popc.b32 %0, %0;" : "+r"(r[i]));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[4 + i]) : "r"(ua), "r"(ub));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(s[i]) : "r"(ua), "r"(ub));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(s[4 + i]) : "r"(ua), "r"(ub));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0xe8;" : "+r"(ua) : "r"(ub), "r"(r[4 + i]));
                }
        } else if (OP == POPC_IMAD) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    asm volatile("popc.b32 %0, %0;" : "+r"(r[i]));
                    asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(s[i]) : "r"(ua), "r"(ub));
                }
        } else if (OP == FFMA_LDS) {
            // 64 FFMA + 4 broadcast LDS.128 per step (the SGEMM k-step mix)
            float4 a0 = *reinterpret_cast<float4*>(&sm[((it * 4 + 0) * 8 + (threadIdx.x & 7)) * 4 & 4095]);
            float4 a1 = *reinterpret_cast<float4*>(&sm[((it * 4 + 1) * 8 + ((threadIdx.x >> 3) & 3)) * 4 & 4095]);
            float4 a2 = *reinterpret_cast<float4*>(&sm[((it * 4 + 2) * 8 + (threadIdx.x & 7)) * 4 & 4095]);
            float4 a3 = *reinterpret_cast<float4*>(&sm[((it * 4 + 3) * 8 + ((threadIdx.x >> 3) & 3)) * 4 & 4095]);
            float av[8] = {a0.x, a0.y, a0.z, a0.w, a2.x, a2.y, a2.z, a2.w};
            float bv[8] = {a1.x, a1.y, a1.z, a1.w, a3.x, a3.y, a3.z, a3.w};
            // 8x8 outer product onto 8 accumulators reused 8 times (dependency distance 8)
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(f[j]) : "f"(av[i]), "f"(bv[j]));
        } else if (OP == LDS128_BCAST || OP == LDS128_QBCAST || OP == LDS128_FULL) {
            int lane = threadIdx.x & 31;
            int idx;
            if (OP == LDS128_BCAST) idx = 0;                 // every lane same 16 B
            else if (OP == LDS128_QBCAST) idx = (lane & 7);   // 8 distinct chunks, replicated over the 4 quarter-warps
            else idx = lane;                                  // 32 distinct chunks = 512 B
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                float4 v;
                uint32_t addr = (uint32_t)__cvta_generic_to_shared(&sm[((it * 8 + u) * 32 + idx) * 4 & 4095]);
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
                f[u] += v.x + v.y + v.z + v.w;
            }
        } else if (OP == HAM8 || OP == HAM8_CSA) {
            // one full 256-bit Hamming distance per iteration step x4: r[] is the query, train words derived from ua/ub
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                uint32_t x[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = r[i] ^ (ua + i * ub + u);
                uint32_t d;
                if (OP == HAM8) {
                    d = __popc(x[0]) + __popc(x[1]) + __popc(x[2]) + __popc(x[3]) + __popc(x[4]) + __popc(x[5]) + __popc(x[6]) + __popc(x[7]);
                } else {
                    uint32_t s0 = x[0] ^ x[1] ^ x[2], c0 = (x[0] & x[1]) | (x[2] & (x[0] ^ x[1]));
                    uint32_t s1 = x[3] ^ x[4] ^ x[5], c1 = (x[3] & x[4]) | (x[5] & (x[3] ^ x[4]));
                    uint32_t s2 = s0 ^ s1 ^ x[6],     c2 = (s0 & s1) | (x[6] & (s0 ^ s1));
                    uint32_t s3 = c0 ^ c1 ^ c2,       c3 = (c0 & c1) | (c2 & (c0 ^ c1));
                    d = __popc(s2) + __popc(x[7]) + 2 * __popc(s3) + 4 * __popc(c3);
                }
                ua += d;
            }
        }
    }
    long long t1 = clock64();
    uint32_t acc = ua;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += r[i] + s[i] + __float_as_uint(f[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
int run(const char* name, double instr_per_iter, int threads, uint32_t* d_out, long long* d_cyc, int nsm) {
    kern<OP><<<nsm, threads>>>(d_out, d_cyc, 12345u);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    kern<OP><<<nsm, threads>>>(d_out, d_cyc, 12345u);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    static long long h[1024];
    CK(cudaMemcpy(h, d_cyc, nsm * sizeof(long long), cudaMemcpyDeviceToHost));
    double avg = 0; for (int i = 0; i < nsm; ++i) avg += h[i]; avg /= nsm;
    double warps = threads / 32.0;
    double winstr = instr_per_iter * ITERS * warps;   // warp-instructions per SM
    printf("%-14s threads=%4d  cycles=%10.0f  warp-instr/clk/SM=%7.3f  lane-ops/clk/SM=%8.2f  ms=%.3f  eff_clk_MHz=%.0f\n",
           name, threads, avg, winstr / avg, winstr * 32 / avg, ms, avg / (ms * 1e3));
    return 0;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("device=%s sm=%d.%d SMs=%d clockRate_kHz=%d smem/SM=%zu regs/SM=%d L2=%d\n", p.name, p.major, p.minor, p.multiProcessorCount, clk,
           p.sharedMemPerMultiprocessor, p.regsPerMultiprocessor, p.l2CacheSize);
    int nsm = p.multiProcessorCount;
    uint32_t* d_out; long long* d_cyc;
    CK(cudaMalloc(&d_out, sizeof(uint32_t) * nsm * 1024)); CK(cudaMalloc(&d_cyc, sizeof(long long) * 1024));
    for (int threads : {128, 256, 512, 1024}) {
        run<FFMA>("FFMA", 32, threads, d_out, d_cyc, nsm);
        run<FFMA2>("FFMA2(x2)", 16, threads, d_out, d_cyc, nsm);
        run<FADD>("FADD", 32, threads, d_out, d_cyc, nsm);
        run<POPC>("POPC", 32, threads, d_out, d_cyc, nsm);
        run<LOP3>("LOP3", 32, threads, d_out, d_cyc, nsm);
        run<IADD3>("IADD", 32, threads, d_out, d_cyc, nsm);
        run<IMAD>("IMAD", 32, threads, d_out, d_cyc, nsm);
        run<FMNMX>("FMNMX", 32, threads, d_out, d_cyc, nsm);
        run<VIMNMX>("VIMNMX", 32, threads, d_out, d_cyc, nsm);
        run<FSETP_SEL>("FSETP+@IADD", 64, threads, d_out, d_cyc, nsm);
        run<REDUX>("REDUX", 32, threads, d_out, d_cyc, nsm);
        run<POPC_LOP3>("POPC+LOP3", 64, threads, d_out, d_cyc, nsm);
        run<POPC_LOP3x2>("POPC+4LOP3", 80, threads, d_out, d_cyc, nsm);
        run<POPC_IMAD>("POPC+IMAD", 64, threads, d_out, d_cyc, nsm);
        run<FFMA_LDS>("64FFMA+4LDS128", 68, threads, d_out, d_cyc, nsm);
        run<LDS128_BCAST>("LDS128 bcast", 8, threads, d_out, d_cyc, nsm);
        run<LDS128_QBCAST>("LDS128 8chunk", 8, threads, d_out, d_cyc, nsm);
        run<LDS128_FULL>("LDS128 full", 8, threads, d_out, d_cyc, nsm);
        run<HAM8>("HAM256 naive", 4, threads, d_out, d_cyc, nsm);      // "instr" here = one 256-bit comparison
        run<HAM8_CSA>("HAM256 csa", 4, threads, d_out, d_cyc, nsm);
        printf("\n");
    }
    return 0;
}

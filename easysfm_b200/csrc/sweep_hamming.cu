// sweep_hamming.cu -- ORB / Hamming distance sweep for sm_100a: the POPC-pipe hot kernel.
//
// Replaces the Fq x Ft x 256-bit XOR/popcount loop + per-row top-2 insertion inside OpenCV's
// BFMatcher("BruteForce-Hamming")::knnMatch that the reference calls at
// cpp_code/src/feature_matching.cpp:74,80 (matchFeaturesORB), and the reverse nearest-neighbour pass a
// mutual cross-check needs (python_code/feature_match.py:26-27).
//
// One persistent CTA walks work units (image pair x range of 1024-row query blocks):
//   producer warp    : one lane streams 256-row train tiles (8 KB, row-major 32 B descriptors) into a
//                      4-stage shared-memory ring with cp.async.bulk + mbarriers (TMA, UBLKCP);
//   8 consumer warps : every thread keeps kHamRQ = 4 query descriptors in registers (8 x 32-bit words
//                      each) and walks the tile; a train descriptor is two broadcast LDS.128.
//                      Per comparison: 8 LOP3 (xor) + 8 POPC + 4 IADD3.  Integer distances are packed as
//                      dist << 20 | index, so plain unsigned min/max give (distance, lowest index) order:
//                        row top-2   : 3 VIMNMX per comparison, branch-free, in registers;
//                        column top-1: min over the thread's 4 rows, one warp REDUX.MIN per train row,
//                                      one shared-memory atomicMin by lane 0.
// Results land in the same 64-bit key scratch the L2 sweep uses (distance in the high word, as an integer).
#include <cstdlib>

#include "esfm_internal.cuh"

namespace esfm {

namespace {

constexpr int kHamThreads = kConsumerThreads + 32;
constexpr int kHamQBlock = kConsumerThreads * kHamRQ;  // 1024 query rows per block
constexpr uint32_t kHamIdxMask = (1u << kHamIdxBits) - 1;

struct HamUnit {
    int pair, q_frame, t_frame;
    int fq, ft;        // rows
    int ntt;           // train tiles
    int qb0, qb1;      // query block range
};

__device__ __forceinline__ HamUnit decode_ham_unit(const SweepParams& p, int unit) {
    HamUnit u;
    u.pair = unit / p.units_per_pair;
    const int part = unit - u.pair * p.units_per_pair;
    const PairDesc pd = p.pairs[u.pair];
    u.q_frame = pd.q_frame;
    u.t_frame = pd.t_frame;
    u.fq = p.frame_rows[pd.q_frame];
    u.ft = p.frame_rows[pd.t_frame];
    u.ntt = (u.ft + kHamTile - 1) / kHamTile;
    const int nqb = (u.fq + kHamQBlock - 1) / kHamQBlock;
    u.qb0 = (int)((long long)nqb * part / p.units_per_pair);
    u.qb1 = (int)((long long)nqb * (part + 1) / p.units_per_pair);
    if (u.ft < 1) u.qb1 = u.qb0;
    return u;
}

// Keys at or above this value mean "no candidate": 0xffffffff is the initial value, and rows past the end of the
// query frame carry the index kHamInvalidRow so that dist * 2^20 + index can never win a column minimum
// (largest real key < 2^29; 0x7ff00000 + 256 * 2^20 still fits in 32 bits).
constexpr uint32_t kHamInvalidRow = 0x7ff00000u;

__device__ __forceinline__ u64 expand_key(uint32_t k) {
    return k >= kHamInvalidRow ? kKeyInit : make_key(k >> kHamIdxBits, k & kHamIdxMask);
}

// 256-bit Hamming distance of a query (8 words in registers) and a train descriptor (two 128-bit words).
//   plain  : 8 LOP3 (xor) + 8 POPC + 4 IADD3                                  -> POPC-pipe bound (16 lanes/clk/SM)
//   CSA    : 8 LOP3 (xor) + 4 carry-save adders (2 LOP3 each: xor3 0x96, majority 0xe8) compress the 8 words to
//            ones, ones, twos, fours -> 4 POPC + 3 shift-adds: halves the load on the scarce POPC pipe by moving
//            work to the 4x wider LOP3 pipe (Harley-Seal).  Bit-exact either way.
// carry-save adder on 32 bit-lanes: sum = a ^ b ^ c (LOP3 0x96), carry = majority(a, b, c) (LOP3 0xe8)
__device__ __forceinline__ void csa(uint32_t& sum, uint32_t& carry, uint32_t a, uint32_t b, uint32_t c) {
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(sum) : "r"(a), "r"(b), "r"(c));
    asm("lop3.b32 %0, %1, %2, %3, 0xe8;" : "=r"(carry) : "r"(a), "r"(b), "r"(c));
}

template <bool CSA>
__device__ __forceinline__ uint32_t hamming256(const uint32_t (&q)[8], const uint4& x0, const uint4& x1) {
    const uint32_t w0 = q[0] ^ x0.x, w1 = q[1] ^ x0.y, w2 = q[2] ^ x0.z, w3 = q[3] ^ x0.w;
    const uint32_t w4 = q[4] ^ x1.x, w5 = q[5] ^ x1.y, w6 = q[6] ^ x1.z, w7 = q[7] ^ x1.w;
    if (!CSA) {
        return __popc(w0) + __popc(w1) + __popc(w2) + __popc(w3) + __popc(w4) + __popc(w5) + __popc(w6) + __popc(w7);
    } else {
        uint32_t s0, c0, s1, c1, s2, c2, s3, c3;
        csa(s0, c0, w0, w1, w2);
        csa(s1, c1, w3, w4, w5);
        csa(s2, c2, s0, s1, w6);
        csa(s3, c3, c0, c1, c2);
        return __popc(s2) + __popc(w7) + 2u * __popc(s3) + 4u * __popc(c3);
    }
}

}  // namespace

template <bool CSA>
__global__ void __launch_bounds__(kHamThreads, 2) sweep_hamming_kernel(const SweepParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint4* Ts = reinterpret_cast<uint4*>(smem_raw);                                   // kHamStages x 256 rows x 2 uint4
    uint64_t* bars = reinterpret_cast<uint64_t*>(Ts + kHamStages * kHamTile * 2);
    uint64_t* fullT = bars;
    uint64_t* emptyT = bars + kHamStages;
    uint32_t* colmin = reinterpret_cast<uint32_t*>(bars + 2 * kHamStages);            // col_cap packed keys

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_units = p.n_pairs * p.units_per_pair;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kHamStages; ++s) {
            mbar_init(&fullT[s], 1);
            mbar_init(&emptyT[s], kConsumerThreads / 32);
        }
        fence_mbar_init();
    }
    __syncthreads();

    if (warp == kConsumerThreads / 32) {
        // ===================== producer warp =====================
        if (lane == 0) {
            uint32_t g = 0;
            for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
                const HamUnit u = decode_ham_unit(p, unit);
                const uint4* tbase = p.rows_b256 + (size_t)p.frame_row_off[u.t_frame] * 2;
                for (int qb = u.qb0; qb < u.qb1; ++qb) {
                    for (int tt = 0; tt < u.ntt; ++tt, ++g) {
                        const uint32_t st = g % kHamStages, ph = (g / kHamStages) & 1;
                        const int n = min(kHamTile, u.ft - tt * kHamTile);
                        mbar_wait_backoff(&emptyT[st], ph ^ 1);
                        mbar_arrive_expect_tx(&fullT[st], (uint32_t)n * 32u);
                        bulk_g2s(Ts + (size_t)st * kHamTile * 2, tbase + (size_t)tt * kHamTile * 2, (uint32_t)n * 32u, &fullT[st]);
                    }
                }
            }
        }
        return;
    }

    // ===================== consumer warps =====================
    uint32_t g = 0;
    for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        const HamUnit u = decode_ham_unit(p, unit);
        consumer_sync();
        for (int x = threadIdx.x; x < u.ft; x += kConsumerThreads) colmin[x] = 0xffffffffu;
        consumer_sync();
        u64* rk1 = p.keys + (size_t)u.pair * 4 * p.stride;
        u64* rk2 = rk1 + p.stride;
        u64* ck1 = rk2 + p.stride;
        const uint4* qbase = p.rows_b256 + (size_t)p.frame_row_off[u.q_frame] * 2;

        for (int qb = u.qb0; qb < u.qb1; ++qb) {
            uint32_t q[kHamRQ][8], qidx[kHamRQ], m1[kHamRQ], m2[kHamRQ];
#pragma unroll
            for (int r = 0; r < kHamRQ; ++r) {
                const int row = qb * kHamQBlock + r * kConsumerThreads + threadIdx.x;
                const bool valid = row < u.fq;
                uint4 a = make_uint4(0, 0, 0, 0), b = a;
                if (valid) {
                    a = __ldg(qbase + (size_t)row * 2);
                    b = __ldg(qbase + (size_t)row * 2 + 1);
                }
                q[r][0] = a.x; q[r][1] = a.y; q[r][2] = a.z; q[r][3] = a.w;
                q[r][4] = b.x; q[r][5] = b.y; q[r][6] = b.z; q[r][7] = b.w;
                qidx[r] = valid ? (uint32_t)row : kHamInvalidRow;  // => this row never wins a column minimum
                m1[r] = 0xffffffffu;
                m2[r] = 0xffffffffu;
            }
            for (int tt = 0; tt < u.ntt; ++tt, ++g) {
                const uint32_t st = g % kHamStages, ph = (g / kHamStages) & 1;
                const int n = min(kHamTile, u.ft - tt * kHamTile);
                mbar_wait(&fullT[st], ph);
                const uint4* T4 = Ts + (size_t)st * kHamTile * 2;
                uint32_t* cm = colmin + tt * kHamTile;
                const uint32_t tg0 = (uint32_t)(tt * kHamTile);
#pragma unroll 2
                for (int t = 0; t < n; ++t) {
                    const uint4 x0 = T4[2 * t], x1 = T4[2 * t + 1];
                    uint32_t cmin = 0xffffffffu;
#pragma unroll
                    for (int r = 0; r < kHamRQ; ++r) {
                        const uint32_t d = hamming256<CSA>(q[r], x0, x1);
                        // packed keys: multiply-adds run on the FMA pipe, which this kernel leaves idle
                        const uint32_t key = d * (1u << kHamIdxBits) + (tg0 + (uint32_t)t);
                        m2[r] = min(m2[r], max(m1[r], key));
                        m1[r] = min(m1[r], key);
                        cmin = min(cmin, d * (1u << kHamIdxBits) + qidx[r]);
                    }
                    const uint32_t cw = __reduce_min_sync(0xffffffffu, cmin);
                    if (lane == 0) atomicMin(cm + t, cw);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&emptyT[st]);
            }
            // each query row is owned by exactly one thread of one unit: plain stores
#pragma unroll
            for (int r = 0; r < kHamRQ; ++r) {
                if (qidx[r] != kHamInvalidRow) {
                    rk1[qidx[r]] = expand_key(m1[r]);
                    rk2[qidx[r]] = expand_key(m2[r]);
                }
            }
        }
        consumer_sync();  // all warps finished updating colmin for this unit
        if (u.qb1 > u.qb0) {
            for (int x = threadIdx.x; x < u.ft; x += kConsumerThreads) {
                const uint32_t k = colmin[x];
                if (k < kHamInvalidRow) atomicMin(ck1 + x, expand_key(k));
            }
        }
    }
}

size_t sweep_hamming_smem_bytes(int col_cap) {
    return (size_t)kHamStages * kHamTile * 32 + 2 * kHamStages * 8 + (size_t)col_cap * 4;
}

int sweep_hamming_max_rows() {
    // the launch sizes the column minima by the tile-padded row count (plan_chunks: stride rounded up to kTile), so the limit
    // is rounded DOWN to a whole tile: a frame the bank accepts always launches
    const size_t cap = (232448 - sweep_hamming_smem_bytes(0)) / 4 / kTile * kTile;
    const size_t lim = (1u << kHamIdxBits) - 1;
    return (int)(cap < lim ? cap : lim);
}

cudaError_t launch_sweep_hamming(const SweepParams& p, int sm_count, cudaStream_t s) {
    const int n_units = p.n_pairs * p.units_per_pair;
    if (n_units <= 0) return cudaSuccess;
    const size_t smem = sweep_hamming_smem_bytes(p.col_cap);
    if (smem > 232448) return cudaErrorInvalidValue;
    // two CTAs per SM (4 warps per scheduler) whenever their shared memory fits side by side
    const int ctas_per_sm = (2 * (smem + 1024) <= 232448) ? 2 : 1;
    const int grid = n_units < sm_count * ctas_per_sm ? n_units : sm_count * ctas_per_sm;
    // ESFM_HAMMING_PLAIN=1 selects the 8-POPC form (kept for A/B measurements; results are identical)
    static const bool plain = [] { const char* e = getenv("ESFM_HAMMING_PLAIN"); return e && e[0] == '1'; }();
    cudaError_t e;
    if (plain) {
        e = cudaFuncSetAttribute(sweep_hamming_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        sweep_hamming_kernel<false><<<grid, kHamThreads, smem, s>>>(p);
    } else {
        e = cudaFuncSetAttribute(sweep_hamming_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        sweep_hamming_kernel<true><<<grid, kHamThreads, smem, s>>>(p);
    }
    return cudaGetLastError();
}

}  // namespace esfm

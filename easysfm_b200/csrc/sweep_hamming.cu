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

__device__ __forceinline__ u64 expand_key(uint32_t k) {
    return k == 0xffffffffu ? kKeyInit : make_key(k >> kHamIdxBits, k & kHamIdxMask);
}

}  // namespace

__global__ void __launch_bounds__(kHamThreads, 1) sweep_hamming_kernel(const SweepParams p) {
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
                qidx[r] = valid ? (uint32_t)row : 0xffffffffu;  // all-ones => this row never wins a column minimum
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
                        const uint32_t d = __popc(q[r][0] ^ x0.x) + __popc(q[r][1] ^ x0.y) + __popc(q[r][2] ^ x0.z) +
                                           __popc(q[r][3] ^ x0.w) + __popc(q[r][4] ^ x1.x) + __popc(q[r][5] ^ x1.y) +
                                           __popc(q[r][6] ^ x1.z) + __popc(q[r][7] ^ x1.w);
                        const uint32_t d20 = d << kHamIdxBits;
                        const uint32_t key = d20 + (tg0 + (uint32_t)t);
                        m2[r] = min(m2[r], max(m1[r], key));
                        m1[r] = min(m1[r], key);
                        cmin = min(cmin, d20 | qidx[r]);
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
                if (qidx[r] != 0xffffffffu) {
                    rk1[qidx[r]] = expand_key(m1[r]);
                    rk2[qidx[r]] = expand_key(m2[r]);
                }
            }
        }
        consumer_sync();  // all warps finished updating colmin for this unit
        if (u.qb1 > u.qb0) {
            for (int x = threadIdx.x; x < u.ft; x += kConsumerThreads) {
                const uint32_t k = colmin[x];
                if (k != 0xffffffffu) atomicMin(ck1 + x, expand_key(k));
            }
        }
    }
}

size_t sweep_hamming_smem_bytes(int col_cap) {
    return (size_t)kHamStages * kHamTile * 32 + 2 * kHamStages * 8 + (size_t)col_cap * 4;
}

int sweep_hamming_max_rows() {
    const size_t cap = (232448 - sweep_hamming_smem_bytes(0)) / 4;
    const size_t lim = (1u << kHamIdxBits) - 1;
    return (int)(cap < lim ? cap : lim);
}

cudaError_t launch_sweep_hamming(const SweepParams& p, int sm_count, cudaStream_t s) {
    const int n_units = p.n_pairs * p.units_per_pair;
    if (n_units <= 0) return cudaSuccess;
    const int grid = n_units < sm_count ? n_units : sm_count;
    const size_t smem = sweep_hamming_smem_bytes(p.col_cap);
    if (smem > 232448) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(sweep_hamming_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    sweep_hamming_kernel<<<grid, kHamThreads, smem, s>>>(p);
    return cudaGetLastError();
}

}  // namespace esfm

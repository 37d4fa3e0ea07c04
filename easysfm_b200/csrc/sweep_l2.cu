// sweep_l2.cu -- SURF / L2 distance sweep for sm_100a: the FP32-pipe hot kernel.
//
// Replaces the Fq x Ft x 64 distance loop + per-row top-K insertion inside OpenCV's
// BFMatcher::knnMatchImpl -> cv::batchDistance that the reference reaches from
// python_code/feature_match.py:33-34 (BFMatcher(NORM_L2).knnMatch(k=2)) and, for the cross-check,
// feature_match.py:26-27; the C++ twin of the call site is cpp_code/src/feature_matching.cpp:125.
//
// One persistent CTA per SM walks work units (image pair x range of 256-row query blocks):
//   producer warp : one elected lane streams k-major 128-row tiles (33,280 B each: 64x128 operands +
//                   128 half squared norms) global -> shared with cp.async.bulk (TMA, UBLKCP) through a
//                   STAGES-deep full/empty mbarrier ring; the 256-row query block has its own barrier pair.
//   8 consumer warps : each thread owns an 8-row x 16-column register tile of
//                   acc = 1/2|q|^2 + 1/2|t|^2 - q.t  (= 1/2 d^2, exact FP32 FMAs, issued as packed FFMA2
//                   so the FMA pipe saturates at half the issue slots; operands via LDS.128, 6 per k-step).
//   epilogue      : acc <= threshold compares against the running second-best of the row (shared memory,
//                   per query block) and of the column (shared memory, per pair).  Hits are rare
//                   (~2 ln F per row/column); they go through 64-bit atomicMin on packed
//                   (distance bits << 32 | index) keys, whose unsigned order is (distance, lowest index) --
//                   OpenCV's tie-break.  The distance matrix never leaves registers.
// The keys hold the two best candidates per query row AND per train row (for the mutual cross-check);
// finalize.cu re-evaluates them in direct form sum (a-b)^2 before the ratio test.
#include "esfm_internal.cuh"

namespace esfm {

namespace {

__device__ __forceinline__ float4 lds128(const float* p) { return *reinterpret_cast<const float4*>(p); }

__device__ __forceinline__ uint4 lds128_volatile_u32(const uint32_t* p) {
    uint4 v;
    asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_u32(p)));
    return v;
}

struct UnitInfo {
    int pair, q_frame, t_frame;
    int nqt, ntt;      // tiles in query / train frame
    int qb0, qb1;      // query-block range of this unit
};

__device__ __forceinline__ UnitInfo decode_unit(const SweepParams& p, int unit) {
    UnitInfo u;
    u.pair = unit / p.units_per_pair;
    const int part = unit - u.pair * p.units_per_pair;
    const PairDesc pd = p.pairs[u.pair];
    u.q_frame = pd.q_frame;
    u.t_frame = pd.t_frame;
    u.nqt = p.frame_tile_off[pd.q_frame + 1] - p.frame_tile_off[pd.q_frame];
    u.ntt = p.frame_tile_off[pd.t_frame + 1] - p.frame_tile_off[pd.t_frame];
    const int nqb = (u.nqt + kQTiles - 1) / kQTiles;
    u.qb0 = (int)((long long)nqb * part / p.units_per_pair);
    u.qb1 = (int)((long long)nqb * (part + 1) / p.units_per_pair);
    if (p.frame_rows[pd.t_frame] < 1) u.qb1 = u.qb0;  // nothing to compare against
    return u;
}

}  // namespace

template <int STAGES>
__global__ void __launch_bounds__(kSweepThreads, 1) sweep_l2_kernel(const SweepParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* Qs = reinterpret_cast<float*>(smem_raw);                       // kQTiles tiles
    float* Ts = Qs + kQTiles * kTileFloats;                               // STAGES tiles
    uint32_t* taur = reinterpret_cast<uint32_t*>(Ts + STAGES * kTileFloats);  // 256 row thresholds (float bits)
    uint64_t* bars = reinterpret_cast<uint64_t*>(taur + kQTiles * kTile);
    uint64_t* fullQ = bars;
    uint64_t* emptyQ = bars + 1;
    uint64_t* fullT = bars + 2;
    uint64_t* emptyT = bars + 2 + STAGES;
    uint32_t* tauc = reinterpret_cast<uint32_t*>(bars + 2 + 2 * STAGES);  // col thresholds, col_cap entries

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_units = p.n_pairs * p.units_per_pair;

    if (threadIdx.x == 0) {
        mbar_init(fullQ, 1);
        mbar_init(emptyQ, kConsumerThreads / 32);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&fullT[s], 1);
            mbar_init(&emptyT[s], kConsumerThreads / 32);
        }
        fence_mbar_init();
    }
    __syncthreads();

    if (warp >= kConsumerThreads / 32) {
        // ============ producer warpgroup: one lane drives the TMA bulk copies, the rest idle ============
        reg_dealloc<40>();
        if (threadIdx.x == kConsumerThreads) {
            uint32_t g = 0, qseq = 0;
            for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
                const UnitInfo u = decode_unit(p, unit);
                const float* qbase = p.kmajor + (size_t)p.frame_tile_off[u.q_frame] * kTileFloats;
                const float* tbase = p.kmajor + (size_t)p.frame_tile_off[u.t_frame] * kTileFloats;
                for (int qb = u.qb0; qb < u.qb1; ++qb) {
                    const int ntq = min(kQTiles, u.nqt - qb * kQTiles);
                    mbar_wait(emptyQ, (qseq & 1) ^ 1);
                    mbar_arrive_expect_tx(fullQ, (uint32_t)ntq * kTileBytes);
                    bulk_g2s(Qs, qbase + (size_t)qb * kQTiles * kTileFloats, (uint32_t)ntq * kTileBytes, fullQ);
                    ++qseq;
                    for (int tt = 0; tt < u.ntt; ++tt, ++g) {
                        const uint32_t st = g % STAGES, ph = (g / STAGES) & 1;
                        mbar_wait(&emptyT[st], ph ^ 1);
                        mbar_arrive_expect_tx(&fullT[st], kTileBytes);
                        bulk_g2s(Ts + (size_t)st * kTileFloats, tbase + (size_t)tt * kTileFloats, kTileBytes, &fullT[st]);
                    }
                }
            }
        }
        return;
    }

    // ================================ consumer warps ================================
    reg_alloc<232>();
    const int qt = warp >> 2;                          // which of the 2 query tiles this warp works on
    const int ty = ((warp & 3) << 2) | (lane >> 3);    // 0..15 : row group inside the tile
    const int tx = lane & 7;                           // 0..7  : column group
    // rows of this thread inside the tile: r(i) = (i>>2)*64 + ty*4 + (i&3), i < 8
    // cols of this thread inside the tile: c(j) = (j>>2)*32 + tx*4 + (j&3), j < 16
    uint32_t g = 0, qseq = 0;
    for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        const UnitInfo u = decode_unit(p, unit);
        consumer_sync();  // everyone is done with the previous unit's column thresholds
        for (int x = threadIdx.x; x < u.ntt * kTile; x += kConsumerThreads) tauc[x] = kFltMaxBits;
        consumer_sync();
        u64* rk1 = p.keys + (size_t)u.pair * 4 * p.stride;
        u64* rk2 = rk1 + p.stride;
        u64* ck1 = rk2 + p.stride;
        u64* ck2 = ck1 + p.stride;

        for (int qb = u.qb0; qb < u.qb1; ++qb) {
            const int ntq = min(kQTiles, u.nqt - qb * kQTiles);
            const bool active = qt < ntq;
            mbar_wait(fullQ, qseq & 1);
            ++qseq;
            {   // reset the thresholds of the 32 rows this warp owns
                const int r = (lane < 16) ? (((warp & 3) << 4) + lane) : (64 + ((warp & 3) << 4) + (lane - 16));
                taur[qt * kTile + r] = kFltMaxBits;
            }
            __syncwarp();
            const float* Qt = Qs + qt * kTileFloats;
            float hq[8];
            if (active) {
                const float4 h0 = lds128(Qt + kDim * kTile + ty * 4), h1 = lds128(Qt + kDim * kTile + 64 + ty * 4);
                hq[0] = h0.x; hq[1] = h0.y; hq[2] = h0.z; hq[3] = h0.w;
                hq[4] = h1.x; hq[5] = h1.y; hq[6] = h1.z; hq[7] = h1.w;
            }
            const int qrow0 = qb * (kQTiles * kTile) + qt * kTile;  // frame row of this warp's tile row 0

            for (int tt = 0; tt < u.ntt; ++tt, ++g) {
                const uint32_t st = g % STAGES, ph = (g / STAGES) & 1;
                mbar_wait(&fullT[st], ph);
                if (active) {
                    const float* Tt = Ts + (size_t)st * kTileFloats;
                    float2 acc[8][8];
                    {   // acc = hq_i + ht_j   (one packed FMA per pair: hq * 1 + ht)
                        float ht[16];
#pragma unroll
                        for (int m = 0; m < 4; ++m) {
                            const float4 h = lds128(Tt + kDim * kTile + m * 32 + tx * 4);
                            ht[m * 4 + 0] = h.x; ht[m * 4 + 1] = h.y; ht[m * 4 + 2] = h.z; ht[m * 4 + 3] = h.w;
                        }
                        const float2 one = make_float2(1.f, 1.f);
#pragma unroll
                        for (int i = 0; i < 8; ++i)
#pragma unroll
                            for (int jp = 0; jp < 8; ++jp)
                                acc[i][jp] = __ffma2_rn(make_float2(hq[i], hq[i]), one, make_float2(ht[2 * jp], ht[2 * jp + 1]));
                    }
                    const float* qp = Qt + ty * 4;
                    const float* tp = Tt + tx * 4;
#pragma unroll 4
                    for (int k = 0; k < kDim; ++k) {
                        const float4 a0 = lds128(qp + k * kTile), a1 = lds128(qp + k * kTile + 64);
                        const float4 b0 = lds128(tp + k * kTile), b1 = lds128(tp + k * kTile + 32);
                        const float4 b2 = lds128(tp + k * kTile + 64), b3 = lds128(tp + k * kTile + 96);
                        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                        const float2 b[8] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w), make_float2(b1.x, b1.y),
                                             make_float2(b1.z, b1.w), make_float2(b2.x, b2.y), make_float2(b2.z, b2.w),
                                             make_float2(b3.x, b3.y), make_float2(b3.z, b3.w)};
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float2 na = make_float2(-a[i], -a[i]);
#pragma unroll
                            for (int jp = 0; jp < 8; ++jp) acc[i][jp] = __ffma2_rn(na, b[jp], acc[i][jp]);
                        }
                    }

                    // ---------------- epilogue: threshold compares (fast path) ----------------
                    uint32_t rowmask = 0, colmask = 0;
                    {   // row minima (3-input FMNMX trees) against the 8 row thresholds
                        const uint4 t0 = lds128_volatile_u32(taur + qt * kTile + ty * 4);
                        const uint4 t1 = lds128_volatile_u32(taur + qt * kTile + 64 + ty * 4);
                        const float tr[8] = {__uint_as_float(t0.x), __uint_as_float(t0.y), __uint_as_float(t0.z), __uint_as_float(t0.w),
                                             __uint_as_float(t1.x), __uint_as_float(t1.y), __uint_as_float(t1.z), __uint_as_float(t1.w)};
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            float m = fminf(acc[i][0].x, acc[i][0].y);
#pragma unroll
                            for (int jp = 1; jp < 8; ++jp) m = fminf(fminf(m, acc[i][jp].x), acc[i][jp].y);
                            if (m <= tr[i]) rowmask |= 1u << i;
                        }
                    }
                    {   // column minima against the 16 column thresholds
                        float tc[16];
#pragma unroll
                        for (int m = 0; m < 4; ++m) {
                            const uint4 t = lds128_volatile_u32(tauc + tt * kTile + m * 32 + tx * 4);
                            tc[m * 4 + 0] = __uint_as_float(t.x); tc[m * 4 + 1] = __uint_as_float(t.y);
                            tc[m * 4 + 2] = __uint_as_float(t.z); tc[m * 4 + 3] = __uint_as_float(t.w);
                        }
#pragma unroll
                        for (int jp = 0; jp < 8; ++jp) {
                            float mx = fminf(acc[0][jp].x, acc[1][jp].x), my = fminf(acc[0][jp].y, acc[1][jp].y);
#pragma unroll
                            for (int i = 2; i < 8; i += 2) {
                                mx = fminf(fminf(mx, acc[i][jp].x), acc[i + 1][jp].x);
                                my = fminf(fminf(my, acc[i][jp].y), acc[i + 1][jp].y);
                            }
                            if (mx <= tc[2 * jp]) colmask |= 1u << (2 * jp);
                            if (my <= tc[2 * jp + 1]) colmask |= 2u << (2 * jp);
                        }
                    }

                    // ---------------- slow path: rare candidate inserts ----------------
                    if (rowmask | colmask) {
                        float vals[128];  // dynamic indexing below => local memory (L1); only touched on a hit
#pragma unroll
                        for (int i = 0; i < 8; ++i)
#pragma unroll
                            for (int jp = 0; jp < 8; ++jp) {
                                vals[i * 16 + 2 * jp] = acc[i][jp].x;
                                vals[i * 16 + 2 * jp + 1] = acc[i][jp].y;
                            }
                        const int col0 = tt * kTile + tx * 4;
                        const int row0 = qrow0 + ty * 4;
#pragma unroll 1
                        for (int i = 0; i < 8; ++i) {
                            if (!((rowmask >> i) & 1)) continue;
                            const int rl = (i >> 2) * 64 + (i & 3);           // tile-local row minus ty*4
                            uint32_t* thp = taur + qt * kTile + ty * 4 + rl;
                            const int grow = row0 + rl;
                            float th = __uint_as_float(*reinterpret_cast<volatile uint32_t*>(thp));
#pragma unroll 1
                            for (int j = 0; j < 16; ++j) {
                                const float v = vals[i * 16 + j];
                                if (v <= th) {
                                    insert_candidate(rk1 + grow, rk2 + grow, thp, v, (uint32_t)(col0 + (j >> 2) * 32 + (j & 3)));
                                    th = __uint_as_float(*reinterpret_cast<volatile uint32_t*>(thp));
                                }
                            }
                        }
#pragma unroll 1
                        for (int j = 0; j < 16; ++j) {
                            if (!((colmask >> j) & 1)) continue;
                            const int gcol = col0 + (j >> 2) * 32 + (j & 3);
                            uint32_t* thp = tauc + gcol;
                            float th = __uint_as_float(*reinterpret_cast<volatile uint32_t*>(thp));
#pragma unroll 1
                            for (int i = 0; i < 8; ++i) {
                                const float v = vals[i * 16 + j];
                                if (v <= th) {
                                    insert_candidate(ck1 + gcol, ck2 + gcol, thp, v, (uint32_t)(row0 + (i >> 2) * 64 + (i & 3)));
                                    th = __uint_as_float(*reinterpret_cast<volatile uint32_t*>(thp));
                                }
                            }
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&emptyT[st]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(emptyQ);
        }
    }
}

size_t sweep_l2_smem_bytes(int col_cap, int stages) {
    return (size_t)(kQTiles + stages) * kTileBytes + kQTiles * kTile * 4 + (2 + 2 * stages) * 8 + (size_t)col_cap * 4;
}

int sweep_l2_max_rows() {
    // column thresholds must fit next to the 2-stage ring inside 227 KB of shared memory
    const size_t fixed = sweep_l2_smem_bytes(0, 2);
    const size_t cap = (232448 - fixed) / 4;
    return (int)(cap / kTile) * kTile;
}

cudaError_t launch_sweep_l2(const SweepParams& p, int sm_count, cudaStream_t s) {
    const int n_units = p.n_pairs * p.units_per_pair;
    if (n_units <= 0) return cudaSuccess;
    const int grid = n_units < sm_count ? n_units : sm_count;
    int stages = 3;
    size_t smem = sweep_l2_smem_bytes(p.col_cap, 3);
    if (smem > 232448) {
        stages = 2;
        smem = sweep_l2_smem_bytes(p.col_cap, 2);
    }
    if (smem > 232448) return cudaErrorInvalidValue;
    cudaError_t e;
    if (stages == 3) {
        e = cudaFuncSetAttribute(sweep_l2_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        sweep_l2_kernel<3><<<grid, kSweepThreads, smem, s>>>(p);
    } else {
        e = cudaFuncSetAttribute(sweep_l2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        sweep_l2_kernel<2><<<grid, kSweepThreads, smem, s>>>(p);
    }
    return cudaGetLastError();
}

}  // namespace esfm

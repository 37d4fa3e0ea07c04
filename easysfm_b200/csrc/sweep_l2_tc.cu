// sweep_l2_tc.cu -- SURF / L2 distance sweep on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// Same contract as sweep_l2.cu (the FP32-FFMA engine): for every image pair, the two best train rows of every
// query row and the best query row of every train row, as packed (1/2 d^2 bits << 32 | index) keys, ranked with
// OpenCV's lowest-index tie-break (BFMatcher(NORM_L2).knnMatch(k=2), python_code/feature_match.py:33-34, and the
// crossCheck of :26-27; C++ call site cpp_code/src/feature_matching.cpp:125).  finalize.cu then re-evaluates the
// candidates in direct form, so what this kernel must get right is the RANKING; its arithmetic is the 3xTF32 split
// product of tc_layout.cuh (|error| ~ 2e-6 on 1/2 d^2, measured by csrc/microbench/tc_probe.cu).
//
// One persistent CTA per SM, 22 warps, four pipelines (TMA -> shared memory -> tensor memory -> registers).  A CTA works
// on a QUERY BLOCK of two 128-row tiles at a time and streams the train frame past it once: every 68 KB train tile feeds
// two 128 x 128 accumulators, so the L2 -> shared-memory traffic per comparison is half of a one-tile block (at one tile
// per block the 148 SMs asked L2 for 6300 B/clk, exactly the measured L2 slice throughput cap).
//   warp 16 (one lane)  TMA producer: one cp.async.bulk per 128-row train tile image (main 64 KB + augmented 4 KB) into a
//                       3-stage full/empty mbarrier ring; the 128 running column thresholds of the tile ride in their own ring.
//   warp 17             MMA issuer: per (train tile, query tile) 25 x tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=128, K=8):
//                       3 terms (lo.hi, hi.lo, hi.hi) x 8 k-steps with the query operand read from TENSOR MEMORY and the train
//                       operand from SWIZZLE_128B shared-memory atoms, + 1 augmented k-step that adds both half norms exactly;
//                       -1/2 d^2 accumulates in one of 2 tensor-memory stages (128 columns each);
//                       tcgen05.commit publishes the accumulator stage and releases the shared-memory stage / query tile.
//   warps 18-21         query writers: fp32 rows -> (hi, lo) TF32 operand of query tile h in tensor-memory columns
//                       256 + 128 h .. (tcgen05.st) + its augmented block in shared memory; tile h of the NEXT block is
//                       written while the MMAs of the other tile still run.
//   warps 0-15          epilogue: warp w owns TMEM lanes 32*(w%4).. (= query rows) and column quarter w/4.  tcgen05.ld
//                       gives each THREAD one query row x 32 train columns, so the row's running top-2 is thread-private
//                       (no shuffles); column minima go through a warp REDUX + one fire-and-forget atomicMin per hit.
//                       The accumulator stage is released as soon as it is in registers.
#include "tc_layout.cuh"

namespace esfm {

namespace {

constexpr int kTcColParts = 4;                          // column quarters of a train tile, one per epilogue warp of a lane quarter
constexpr int kTcEpiWarps = 4 * kTcColParts;            // 16 epilogue warps: enough to hide the epilogue's dependent-issue latency
constexpr int kTcEpiThreads = kTcEpiWarps * 32;
constexpr int kTcThreads = kTcEpiThreads + 64 + 128;    // + TMA producer warp + MMA issuer warp + 4 query-writer warps
constexpr int kTcWriterWarp0 = kTcEpiWarps + 2;         // warps 18..21: warp % 4 covers the four TMEM lane quarters
constexpr int kTcQTiles = 2;                            // query tiles per block: each train tile in shared memory is used twice
constexpr uint32_t kTcACol0 = 256;                      // tensor-memory columns of query tile h: 256 + 128 h (hi 64 | lo 64)
constexpr int kTcPartCols = kTile / kTcColParts;        // 32 columns per epilogue thread and stage
constexpr int kTcStages = 3;                 // shared-memory train stages (68 KB each)
constexpr int kTcAccStages = 2;              // tensor-memory accumulator stages (128 columns each; columns 256..511 hold the query operands)
constexpr int kTcThrBytes = kTile * 4;                  // 512: column thresholds riding with a train tile
constexpr int kTcThrStages = 4;                         // threshold snapshots have their own (deeper) ring
constexpr uint32_t kTcBoundBits = 0x6f6f6f6fu;          // 7.4e28f: "no bound yet" (what memset(0x6f) writes); pads are 1e30

struct TcUnit {
    int pair, q_frame, t_frame;
    int nqt, ntt;
    int qb0, qb1;        // query TILES [qb0, qb1) of this unit; walked in blocks of kTcQTiles
};

__device__ __forceinline__ TcUnit tc_decode_unit(const SweepParams& p, int unit) {
    TcUnit u;
    u.pair = unit / p.units_per_pair;
    const int part = unit - u.pair * p.units_per_pair;
    const PairDesc pd = p.pairs[u.pair];
    u.q_frame = pd.q_frame;
    u.t_frame = pd.t_frame;
    u.nqt = p.frame_tile_off[pd.q_frame + 1] - p.frame_tile_off[pd.q_frame];
    u.ntt = p.frame_tile_off[pd.t_frame + 1] - p.frame_tile_off[pd.t_frame];
    u.qb0 = (int)((long long)u.nqt * part / p.units_per_pair);
    u.qb1 = (int)((long long)u.nqt * (part + 1) / p.units_per_pair);
    if (p.frame_rows[pd.t_frame] < 1) u.qb1 = u.qb0;
    return u;
}

struct RowTop2 {
    float v1, v2;        // 1/2 d^2 of the best / second best so far
    uint32_t i1, i2;
};

// Wait of the single producer / MMA-issuing lanes: each shares an SM sub-partition with four epilogue warps, so a tight
// try_wait loop (measured: 3 issue slots every ~7 cycles) steals a large part of their issue bandwidth, while the 2 us
// back-off of the FFMA sweep's producer is longer than a whole tile here (~1 us) and starves the tensor pipe.  Poll every
// few tens of nanoseconds instead.
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity), "r"(20000u)
            : "memory");
        if (ok) break;
        __nanosleep(40);
    }
}

// Wait of the epilogue warps: let the hardware park the warp (suspend-time hint) instead of spinning on try_wait -- the
// spin loop was 21 % of all executed instructions, taken from the issue slots of the warps that had work.
__device__ __forceinline__ void mbar_wait_parked(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity), "r"(100000u)
            : "memory");
    } while (!ok);
}

}  // namespace

__global__ void __launch_bounds__(kTcThreads, 1) sweep_l2_tc_kernel(const SweepParams p) {
    extern __shared__ unsigned char smem_raw[];
    // SWIZZLE_128B atoms need 1024-byte alignment
    unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* Ts = base;                                   // kTcStages train tile images (each 68 x 1024 B: atoms stay aligned)
    unsigned char* Qa = Ts + kTcStages * kTcTileBytes;          // augmented blocks (1, 1, 1, hq_h | hq_m, hq_l, 0, 0) of the two query tiles
    unsigned char* Thr = Qa + kTcQTiles * kTcAugBytes;          // kTcThrStages x 128 column thresholds
    u64* mkey = reinterpret_cast<u64*>(Thr + kTcThrStages * kTcThrBytes);          // [tile h][best, second][128] merged row keys
    uint32_t* sbound = reinterpret_cast<uint32_t*>(mkey + kTcQTiles * 2 * kTile);  // [tile h][128] row bound shared by a row's parts
    uint64_t* bars = reinterpret_cast<uint64_t*>(sbound + kTcQTiles * kTile);
    uint64_t* fullQ = bars;                      // [kTcQTiles]
    uint64_t* emptyQ = fullQ + kTcQTiles;        // [kTcQTiles]
    uint64_t* fullT = emptyQ + kTcQTiles;
    uint64_t* emptyT = fullT + kTcStages;
    uint64_t* accFull = emptyT + kTcStages;
    uint64_t* accEmpty = accFull + kTcAccStages;
    uint64_t* thrFull = accEmpty + kTcAccStages;
    uint64_t* thrEmpty = thrFull + kTcThrStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(thrEmpty + kTcThrStages);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_units = p.n_pairs * p.units_per_pair;

    if (threadIdx.x == 0) {
        for (int h = 0; h < kTcQTiles; ++h) {
            mbar_init(&fullQ[h], 4);                  // the 4 query-writer warps
            mbar_init(&emptyQ[h], 1);                 // MMA commit
        }
        for (int s = 0; s < kTcStages; ++s) {
            mbar_init(&fullT[s], 1);
            mbar_init(&emptyT[s], 1);                 // MMA commit
        }
        for (int s = 0; s < kTcThrStages; ++s) {
            mbar_init(&thrFull[s], 1);
            mbar_init(&thrEmpty[s], kTcEpiWarps);
        }
        for (int s = 0; s < kTcAccStages; ++s) {
            mbar_init(&accFull[s], 1);
            mbar_init(&accEmpty[s], kTcEpiWarps);
        }
        fence_mbar_init();
    }
    if (warp == kTcEpiWarps + 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == kTcEpiWarps) {
        // ======================= TMA producer =======================
        if (lane == 0) {
            uint32_t g = 0;
            for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
                const TcUnit u = tc_decode_unit(p, unit);
                const unsigned char* timg = p.tc_main + (size_t)p.frame_tile_off[u.t_frame] * kTcTileBytes;
                const uint32_t* tauc = p.col_thr + (size_t)u.pair * p.stride;
                for (int qt = u.qb0; qt < u.qb1; qt += kTcQTiles) {
                    for (int tt = 0; tt < u.ntt; ++tt, ++g) {
                        const uint32_t st = g % kTcStages, ph = (g / kTcStages) & 1;
                        mbar_wait_relaxed(&emptyT[st], ph ^ 1);
                        mbar_arrive_expect_tx(&fullT[st], kTcTileBytes);
                        if (p.debug_flags & 4)
                            asm volatile("mbarrier.complete_tx.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&fullT[st])), "r"(kTcTileBytes) : "memory");
                        else
                            bulk_g2s(Ts + (size_t)st * kTcTileBytes, timg + (size_t)tt * kTcTileBytes, kTcTileBytes, &fullT[st]);
                        // the running column thresholds of this tile ride along in their own ring (a snapshot a few tiles
                        // old is fine: a stale threshold is only looser, never wrong)
                        const uint32_t ts = g % kTcThrStages, tph = (g / kTcThrStages) & 1;
                        mbar_wait_relaxed(&thrEmpty[ts], tph ^ 1);
                        mbar_arrive_expect_tx(&thrFull[ts], kTcThrBytes);
                        bulk_g2s(Thr + ts * kTcThrBytes, tauc + (size_t)tt * kTile, kTcThrBytes, &thrFull[ts]);
                    }
                }
            }
        }
    } else if (warp == kTcEpiWarps + 1) {
        // ======================= MMA issuer =======================
        // The WHOLE warp walks the loops and waits on the barriers (warp-uniform control flow, so the shared-memory
        // descriptors live in uniform registers); one elected lane issues the tcgen05.mma / tcgen05.commit instructions.
        // (Issuing from inside `if (lane == 0)` made the compiler wrap every MMA in an ELECT / R2UR.BROADCAST waterfall
        // loop: 16 dependent instructions and ~90 cycles per MMA, longer than the MMA itself.)
        constexpr uint32_t idesc = tc_idesc_tf32(128, 128);
        const uint64_t qad0 = tc_desc_nosw(smem_u32(Qa), 128, kTcAugGroupBytes);
        const uint64_t td0 = tc_desc_sw128(smem_u32(Ts), kTcGroupBytes);
        const uint64_t tad0 = tc_desc_nosw(smem_u32(Ts) + kTcMainBytes, 128, kTcAugGroupBytes);
        uint32_t g = 0, a = 0;               // train tiles streamed, accumulator jobs issued
        uint32_t qn0 = 0, qn1 = 0;           // how many times query slot 0 / 1 has been filled (slot 1 is absent from an odd tail block)
        for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
            const TcUnit u = tc_decode_unit(p, unit);
            for (int qt = u.qb0; qt < u.qb1; qt += kTcQTiles) {
                const int nh = min(kTcQTiles, u.qb1 - qt);
                for (int tt = 0; tt < u.ntt; ++tt, ++g) {
                    const uint32_t st = g % kTcStages, ph = (g / kTcStages) & 1;
                    mbar_wait_relaxed(&fullT[st], ph);
                    // descriptor start addresses are in 16-byte units: adding (bytes >> 4) to the low word moves the window
                    const uint64_t td = td0 + (uint64_t)(st * (kTcTileBytes >> 4));
                    const uint64_t tad = tad0 + (uint64_t)(st * (kTcTileBytes >> 4));
#pragma unroll 1
                    for (int h = 0; h < nh; ++h, ++a) {
                        const uint32_t as = a % kTcAccStages, aph = (a / kTcAccStages) & 1;
                        if (tt == 0) mbar_wait_relaxed(&fullQ[h], (h == 0 ? qn0 : qn1) & 1);   // the writers have filled slot h for this block
                        mbar_wait_relaxed(&accEmpty[as], aph ^ 1);
                        tc_fence_after();
                        const uint32_t d = tmem + as * 128;
                        const uint32_t acol = tmem + kTcACol0 + h * 128;
                        if (!(p.debug_flags & 2) && elect_one()) {
                            bool first = true;
#pragma unroll
                            for (int term = 0; term < 3; ++term) {
                                // (A part, B part): lo.hi, hi.lo first (small terms), hi.hi last
                                const int pa = term == 0 ? 1 : 0, pb = term == 1 ? 1 : 0;
#pragma unroll
                                for (int ks = 0; ks < 8; ++ks) {
                                    const uint32_t off = ((ks >> 2) * 1024 + (ks & 3) * 32) >> 4;
                                    // A (query) from tensor memory: lane = row, column = k.  With both operands in shared memory the
                                    // 8 KB of operand reads per MMA plus the TMA writes saturate the 128 B/clk shared-memory port
                                    // (measured 87 cycles per MMA instead of 64).
                                    tc_mma_tf32_ts(d, acol + pa * 64 + ks * 8, td + (uint64_t)(pb * (2048 >> 4) + off), idesc, !first);
                                    first = false;
                                }
                            }
                            // - 1/2|q|^2 - 1/2|t|^2, exact (three-way split half norms against ones)
                            tc_mma_tf32(d, qad0 + (uint64_t)(h * (kTcAugBytes >> 4)), tad, idesc, true);
                        }
                        __syncwarp();
                        if (elect_one()) {
                            tc_commit(&accFull[as]);                        // accumulator stage ready for the epilogue
                            if (tt == u.ntt - 1) tc_commit(&emptyQ[h]);     // query tile h may be overwritten once these MMAs retire
                            if (h == nh - 1) tc_commit(&emptyT[st]);        // shared-memory stage consumed
                        }
                        __syncwarp();
                    }
                }
                ++qn0;
                if (nh == kTcQTiles) ++qn1;
            }
        }
    } else if (warp >= kTcWriterWarp0) {
        // ======================= query writers: fp32 rows -> (hi, lo) TF32 operand in tensor memory =======================
        const int quarter = warp & 3;
        const int trow = quarter * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        uint32_t quse[kTcQTiles] = {0, 0};
        for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
            const TcUnit u = tc_decode_unit(p, unit);
            const int fq = p.frame_rows[u.q_frame];
            const float4* qrows = reinterpret_cast<const float4*>(p.rows_f32 + (size_t)p.frame_row_off[u.q_frame] * kDim);
            for (int qt = u.qb0; qt < u.qb1; ++qt) {
                const int h = (qt - u.qb0) & (kTcQTiles - 1);
                const int r = qt * kTile + trow;
                float4 x[16];
#pragma unroll
                for (int m = 0; m < 16; ++m) x[m] = r < fq ? __ldg(qrows + (size_t)r * 16 + m) : make_float4(0.f, 0.f, 0.f, 0.f);
                float hs = 0.f;     // 1/2|q|^2 in the summation order of bank.cu's pack kernels
#pragma unroll
                for (int m = 0; m < 16; ++m) {
                    hs = __fmaf_rn(x[m].x, x[m].x, hs); hs = __fmaf_rn(x[m].y, x[m].y, hs);
                    hs = __fmaf_rn(x[m].z, x[m].z, hs); hs = __fmaf_rn(x[m].w, x[m].w, hs);
                }
                const float hq = r < fq ? 0.5f * hs : kTcPadNorm;   // pad rows can never win a column
                float hqh, hqm, hql;
                tc_split3(hq, hqh, hqm, hql);
                const uint32_t use = h == 0 ? quse[0] : quse[1];
                mbar_wait(&emptyQ[h], (use & 1) ^ 1);      // every MMA reading the previous tile in this slot has retired
                if (h == 0) ++quse[0]; else ++quse[1];
                tc_fence_after();
                unsigned char* qa = Qa + h * kTcAugBytes + (trow >> 3) * kTcAugGroupBytes + (trow & 7) * 16;
                *reinterpret_cast<float4*>(qa) = make_float4(1.f, 1.f, 1.f, hqh);
                *reinterpret_cast<float4*>(qa + 128) = make_float4(hqm, hql, 0.f, 0.f);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the MMA's async reads
                const uint32_t acol = tmem + lane_addr + kTcACol0 + h * 128;
#pragma unroll
                for (int m = 0; m < 8; ++m) {
                    const float xs[8] = {x[2 * m].x, x[2 * m].y, x[2 * m].z, x[2 * m].w, x[2 * m + 1].x, x[2 * m + 1].y, x[2 * m + 1].z, x[2 * m + 1].w};
                    uint32_t hi[8], lo[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float hv = tc_tf32_hi(xs[j]);
                        hi[j] = __float_as_uint(hv);
                        lo[j] = __float_as_uint(xs[j] - hv);
                    }
                    tmem_st8(acol + m * 8, hi);
                    tmem_st8(acol + 64 + m * 8, lo);
                }
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&fullQ[h]);
            }
        }
    } else {
        // ======================= epilogue warps =======================
        const int quarter = warp & 3, part = warp >> 2;
        const int trow = quarter * 32 + lane;               // row inside a 128-row query tile (= TMEM lane)
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        if (part < kTcQTiles) {
            sbound[part * kTile + trow] = kTcBoundBits;
            mkey[(part * 2) * kTile + trow] = kKeyInit;
            mkey[(part * 2 + 1) * kTile + trow] = kKeyInit;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kTcEpiThreads) : "memory");
        uint32_t g = 0, a = 0;
        for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
            const TcUnit u = tc_decode_unit(p, unit);
            u64* rk1 = p.keys + (size_t)u.pair * 4 * p.stride;
            u64* rk2 = rk1 + p.stride;
            u64* ck1 = rk2 + p.stride;
            uint32_t* tauc = p.col_thr + (size_t)u.pair * p.stride;
            for (int qt = u.qb0; qt < u.qb1; qt += kTcQTiles) {
                const int nh = min(kTcQTiles, u.qb1 - qt);
                // running top-2 of this thread's row in query tile 0 (t) and tile 1 (to); the two swap after every accumulator
                RowTop2 t, to;
                t.v1 = t.v2 = to.v1 = to.v2 = __uint_as_float(kTcBoundBits);
                t.i1 = t.i2 = to.i1 = to.i2 = 0xffffffffu;
                for (int tt = 0; tt < u.ntt; ++tt, ++g) {
                    const uint32_t ts = g % kTcThrStages, tph = (g / kTcThrStages) & 1;
                    mbar_wait_parked(&thrFull[ts], tph);
                    const float4* tp = reinterpret_cast<const float4*>(Thr + ts * kTcThrBytes) + part * (kTcPartCols / 4);
#pragma unroll 1
                    for (int h = 0; h < nh; ++h, ++a) {
                        const uint32_t as = a % kTcAccStages, aph = (a / kTcAccStages) & 1;
                        const uint32_t qrow = (uint32_t)((qt + h) * kTile + trow);     // frame row of this thread
                        uint32_t* sb = &sbound[h * kTile + trow];
                        mbar_wait_parked(&accFull[as], aph);
                        tc_fence_after();
                        const uint32_t taddr = tmem + lane_addr + as * 128 + part * kTcPartCols;
                        // All 32 columns of this thread go to registers at once and the accumulator stage is released right
                        // away: with only two stages the tensor pipe must never wait for the selection logic below.
                        uint32_t vb[16], vn[16];
                        tmem_ld16(taddr, vb);
                        tmem_ld16(taddr + 16, vn);
                        // The row's running second best over ALL column parts (each part keeps a private top-2; the shared
                        // bound only filters, with '>=' so equal values still reach the private strict-'<' insertion).
                        float nb = -__uint_as_float(*reinterpret_cast<volatile uint32_t*>(sb));
                        tmem_ld_wait();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&accEmpty[as]);
                        // columns in chunks of 16: a real loop, so the epilogue body stays small enough for the instruction
                        // cache (a fully unrolled 64-column body was 64 KB of SASS and stalled on instruction fetch)
#pragma unroll 1
                        for (int ch = 0; ch < kTcPartCols / 16; ++ch) {
                            if (ch > 0) {
#pragma unroll
                                for (int c = 0; c < 16; ++c) vb[c] = vn[c];     // second half of the columns
                            }
                            float thr[16];
#pragma unroll
                            for (int m = 0; m < 4; ++m) {
                                const float4 x = tp[ch * 4 + m];
                                thr[4 * m] = x.x; thr[4 * m + 1] = x.y; thr[4 * m + 2] = x.z; thr[4 * m + 3] = x.w;
                            }
                            if (ch == kTcPartCols / 16 - 1 && h == nh - 1) {      // last read of this threshold snapshot
                                __syncwarp();
                                if (lane == 0) mbar_arrive(&thrEmpty[ts]);
                            }
                            float v[16];
#pragma unroll
                            for (int c = 0; c < 16; ++c) v[c] = __uint_as_float(vb[c]);   // v = -1/2 d^2
                            if (p.debug_flags & 1) continue;
                            // ---- fast path (~40 instructions): one row test + 16 column tests + ONE vote ----
                            float gmx[4];
#pragma unroll
                            for (int gq = 0; gq < 4; ++gq)
                                gmx[gq] = fmaxf(fmaxf(fmaxf(v[4 * gq], v[4 * gq + 1]), v[4 * gq + 2]), v[4 * gq + 3]);
                            const bool rflag = fmaxf(fmaxf(fmaxf(gmx[0], gmx[1]), gmx[2]), gmx[3]) >= nb;
                            bool cflag = false;
#pragma unroll
                            for (int j = 0; j < 16; ++j) cflag |= (v[j] >= -thr[j]);
                            if (!__any_sync(0xffffffffu, rflag || cflag)) continue;

                            // ---- slow path: ~2 ln F hits per row and ~ln F per column over a whole sweep ----
                            const uint32_t col0 = (uint32_t)(tt * kTile + part * kTcPartCols + ch * 16);
                            if (rflag) {
                                bool ins = false;
#pragma unroll
                                for (int gq = 0; gq < 4; ++gq) {
                                    if (gmx[gq] >= nb) {
#pragma unroll
                                        for (int j = 0; j < 4; ++j) {
                                            const float d = -v[4 * gq + j];
                                            if (d < t.v2) {     // ascending column order + strict '<' keeps the lowest index on ties
                                                const uint32_t idx = col0 + 4 * gq + j;
                                                if (d < t.v1) {
                                                    t.v2 = t.v1; t.i2 = t.i1;
                                                    t.v1 = d;    t.i1 = idx;
                                                } else {
                                                    t.v2 = d;    t.i2 = idx;
                                                }
                                                ins = true;
                                            }
                                        }
                                    }
                                }
                                if (ins) {
                                    atomicMin(sb, __float_as_uint(fmaxf(t.v2, 0.f)));
                                    nb = fmaxf(nb, -t.v2);
                                }
                            }
                            if (__any_sync(0xffffffffu, cflag)) {
#pragma unroll
                                for (int gq = 0; gq < 2; ++gq) {
                                    bool any = false;
#pragma unroll
                                    for (int j = 0; j < 8; ++j) any |= (v[8 * gq + j] >= -thr[8 * gq + j]);
                                    if (__any_sync(0xffffffffu, any)) {
#pragma unroll
                                        for (int j = 0; j < 8; ++j) {
                                            const bool hit = v[8 * gq + j] >= -thr[8 * gq + j];
                                            const uint32_t bal = __ballot_sync(0xffffffffu, hit);
                                            if (bal) {     // warp-uniform
                                                const uint32_t bits = hit ? __float_as_uint(fmaxf(-v[8 * gq + j], 0.f)) : 0xffffffffu;
                                                const uint32_t mn = __reduce_min_sync(0xffffffffu, bits);
                                                const uint32_t win = __ballot_sync(0xffffffffu, bits == mn);
                                                if (lane == __ffs(win) - 1) {     // lowest lane = lowest query row among equals
                                                    uint32_t gcol = col0 + 8 * gq + j;
                                                    asm volatile("" : "+r"(gcol));   // keep the 64-bit address arithmetic inside this (rare) branch
                                                    atomicMin(ck1 + gcol, make_key(mn, qrow));
                                                    atomicMin(tauc + gcol, mn);
                                                }
                                            }
                                        }
                                    }
                                }
                            }
                        }
                        if (nh == kTcQTiles) {      // next accumulator belongs to the other query tile
                            const RowTop2 x = t;
                            t = to;
                            to = x;
                        }
                    }
                }
                // ---- end of the sweep for this query block: merge the column parts of every row, publish ----
                // 64-bit shared-memory atomics on packed keys: the smallest key ends in mkey[.][0], the smallest of all the
                // "losers" (displaced old minimum, or the newcomer if it did not win) in mkey[.][1] = the second smallest overall.
                // (an even number of swaps per train tile: t is tile 0's state again, `to` tile 1's)
#pragma unroll
                for (int h = 0; h < kTcQTiles; ++h) {
                    if (h < nh) {
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const RowTop2& s = h ? to : t;
                            const uint32_t idx = e ? s.i2 : s.i1;
                            if (idx != 0xffffffffu) {
                                const u64 k = make_key(__float_as_uint(fmaxf(e ? s.v2 : s.v1, 0.f)), idx);
                                const u64 old = atomicMin(&mkey[(h * 2) * kTile + trow], k);
                                atomicMin(&mkey[(h * 2 + 1) * kTile + trow], old > k ? old : k);
                            }
                        }
                    }
                }
                asm volatile("bar.sync 1, %0;" ::"n"(kTcEpiThreads) : "memory");
                if (part < nh) {       // column part h publishes (and resets) query tile h
                    const uint32_t qrow = (uint32_t)((qt + part) * kTile + trow);
                    rk1[qrow] = mkey[(part * 2) * kTile + trow];
                    rk2[qrow] = mkey[(part * 2 + 1) * kTile + trow];
                    mkey[(part * 2) * kTile + trow] = kKeyInit;
                    mkey[(part * 2 + 1) * kTile + trow] = kKeyInit;
                    sbound[part * kTile + trow] = kTcBoundBits;
                }
                asm volatile("bar.sync 1, %0;" ::"n"(kTcEpiThreads) : "memory");   // mkey[] / sbound[] are reused by the next query block
            }
        }
    }

    // ---- teardown: everything issued has been consumed (the epilogue waited on every accumulator stage) ----
    tc_fence_before();
    __syncthreads();
    if (warp == kTcEpiWarps + 1) tmem_free(tmem, 512);
}

size_t sweep_l2_tc_smem_bytes() {
    return 1024 + (size_t)kTcStages * kTcTileBytes + (size_t)kTcQTiles * kTcAugBytes + (size_t)kTcThrStages * kTcThrBytes +
           (size_t)kTcQTiles * 2 * kTile * sizeof(u64) + (size_t)kTcQTiles * kTile * 4 +
           (2 * kTcQTiles + 2 * kTcStages + 2 * kTcAccStages + 2 * kTcThrStages) * 8 + 16;
}

cudaError_t launch_sweep_l2_tc(const SweepParams& p, int sm_count, cudaStream_t s) {
    const int n_units = p.n_pairs * p.units_per_pair;
    if (n_units <= 0) return cudaSuccess;
    const int grid = n_units < sm_count ? n_units : sm_count;
    const size_t smem = sweep_l2_tc_smem_bytes();
    cudaError_t e = cudaFuncSetAttribute(sweep_l2_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    sweep_l2_tc_kernel<<<grid, kTcThreads, smem, s>>>(p);
    return cudaGetLastError();
}

}  // namespace esfm

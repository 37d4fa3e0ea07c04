// sweep_l2_tc.cu -- SURF / L2 distance sweep on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// Same contract as sweep_l2.cu (the FP32-FFMA engine): for every image pair, the two best train rows of every
// query row and the best query row of every train row, as packed (1/2 d^2 bits << 32 | index) keys, ranked with
// OpenCV's lowest-index tie-break (BFMatcher(NORM_L2).knnMatch(k=2), python_code/feature_match.py:33-34, and the
// crossCheck of :26-27; C++ call site cpp_code/src/feature_matching.cpp:125).  finalize.cu then re-evaluates the
// candidates in direct form, so what this kernel must get right is the RANKING; its arithmetic is the 3xTF32 split
// product of tc_layout.cuh (|error| ~ 2e-6 on 1/2 d^2, measured by csrc/microbench/tc_probe.cu).
//
// One persistent CTA per SM, 10 warps, three pipelines (TMA -> shared memory -> tensor memory -> registers):
//   warp 8 (one lane)  TMA producer: 128-row operand images (main 64 KB + augmented hi/lo 2 x 4 KB) with cp.async.bulk;
//                      the query tile once per query block, train tiles through a 2-stage full/empty mbarrier ring;
//                      the 128 running column thresholds of a train tile ride along with it.
//   warp 9 (one lane)  MMA issuer: per train tile 27 x tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=128, K=8):
//                      8 k-steps over SWIZZLE_128B atoms + 1 augmented k-step (the norms), for hi.hi, hi.lo, lo.hi,
//                      accumulating -1/2 d^2 in one of 4 tensor-memory stages (128 columns each);
//                      tcgen05.commit releases the shared-memory stage and publishes the accumulator stage.
//   warps 0-7          epilogue: warp w owns TMEM lanes 32*(w%4).. (= query rows) and column half w/4.  tcgen05.ld
//                      gives each THREAD one query row x 64 train columns, so the row's running top-2 is thread-private
//                      (no shuffles); column minima go through a warp REDUX + one fire-and-forget atomicMin per hit.
//                      The accumulator stage is released as soon as it is in registers.
#include "tc_layout.cuh"

namespace esfm {

namespace {

constexpr int kTcColParts = 4;                          // column quarters of a train tile, one per epilogue warp of a lane quarter
constexpr int kTcEpiWarps = 4 * kTcColParts;            // 16 epilogue warps: enough to hide the epilogue's dependent-issue latency
constexpr int kTcEpiThreads = kTcEpiWarps * 32;
constexpr int kTcThreads = kTcEpiThreads + 64 + 128;    // + TMA producer warp + MMA issuer warp + 4 query-writer warps
constexpr int kTcWriterWarp0 = kTcEpiWarps + 2;         // warps 18..21: warp % 4 covers the four TMEM lane quarters
constexpr uint32_t kTcAColHi = 384, kTcAColLo = 448;    // tensor-memory columns of the query operand (hi, lo: 64 each)
constexpr int kTcPartCols = kTile / kTcColParts;        // 32 columns per epilogue thread and stage
constexpr int kTcStages = 3;                 // shared-memory train stages (2 starve the tensor pipe: a 72 KB tile takes longer to
                                             // land than one tile's MMAs while those MMAs are reading the same shared memory)
constexpr int kTcQaBytes = 16 * 128 + 128;   // one (hi | lo) block of the query tile's augmented columns: 16 B per row + pad
constexpr int kTcAccStages = 3;              // tensor-memory accumulator stages (128 columns each; columns 384..511 hold the query operand)
constexpr int kTcMainBytes = 16 * kTcGroupBytes;       // 65536: main image of a 128-row tile
constexpr int kTcAugBytes = 16 * kTcAugGroupBytes;     // 4096: one (role, part) augmented image of a tile
constexpr int kTcTileBytes = kTcMainBytes + 2 * kTcAugBytes;   // 73728 bytes per operand tile in shared memory
constexpr int kTcThrBytes = kTile * 4;                  // 512: column thresholds riding with a train tile
constexpr int kTcThrStages = 4;                         // threshold snapshots have their own (deeper) ring
constexpr uint32_t kTcBoundBits = 0x6f6f6f6fu;          // 7.4e28f: "no bound yet" (what memset(0x6f) writes); pads are 1e30

struct TcUnit {
    int pair, q_frame, t_frame;
    int nqt, ntt;
    int qb0, qb1;
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

}  // namespace

__global__ void __launch_bounds__(kTcThreads, 1) sweep_l2_tc_kernel(const SweepParams p) {
    extern __shared__ unsigned char smem_raw[];
    // SWIZZLE_128B atoms need 1024-byte alignment
    unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* Ts = base;                                   // kTcStages train tile images (each 72 x 1024 B: atoms stay aligned)
    unsigned char* Qa = Ts + kTcStages * kTcTileBytes;          // augmented columns (1, 1/2|q|^2, 0, 0) of the query tile: hi, lo
    unsigned char* Thr = Qa + 2 * kTcQaBytes;                   // kTcThrStages x 128 column thresholds
    u64* mkey = reinterpret_cast<u64*>(Thr + kTcThrStages * kTcThrBytes);          // [2][128] merged best / second-best row keys
    uint32_t* sbound = reinterpret_cast<uint32_t*>(mkey + 2 * kTile);              // [128] row bound shared by a row's parts
    uint64_t* bars = reinterpret_cast<uint64_t*>(sbound + kTile);
    uint64_t* fullQ = bars;
    uint64_t* emptyQ = bars + 1;
    uint64_t* fullT = bars + 2;
    uint64_t* emptyT = fullT + kTcStages;
    uint64_t* accFull = emptyT + kTcStages;
    uint64_t* accEmpty = accFull + kTcAccStages;
    uint64_t* thrFull = accEmpty + kTcAccStages;
    uint64_t* thrEmpty = thrFull + kTcThrStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(thrEmpty + kTcThrStages);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_units = p.n_pairs * p.units_per_pair;

    if (threadIdx.x == 0) {
        mbar_init(fullQ, 4);                          // the 4 query-writer warps
        mbar_init(emptyQ, 1);
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
            const size_t aug_part = (size_t)p.tc_groups * kTcAugGroupBytes;   // bytes of one (role, part) augmented array
            uint32_t g = 0;
            for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
                const TcUnit u = tc_decode_unit(p, unit);
                const size_t tg0 = (size_t)p.frame_tile_off[u.t_frame] * 16;
                const uint32_t* tauc = p.col_thr + (size_t)u.pair * p.stride;
                for (int qb = u.qb0; qb < u.qb1; ++qb) {
                    for (int tt = 0; tt < u.ntt; ++tt, ++g) {
                        const uint32_t st = g % kTcStages, ph = (g / kTcStages) & 1;
                        mbar_wait_relaxed(&emptyT[st], ph ^ 1);
                        mbar_arrive_expect_tx(&fullT[st], kTcTileBytes);
                        unsigned char* dst = Ts + (size_t)st * kTcTileBytes;
                        const size_t tg = tg0 + (size_t)tt * 16;
                        if (p.debug_flags & 4) { mbar_arrive_expect_tx(&fullT[st], 0); asm volatile("mbarrier.complete_tx.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&fullT[st])), "r"(kTcTileBytes) : "memory"); } else {
                        bulk_g2s(dst, p.tc_main + tg * kTcGroupBytes, kTcMainBytes, &fullT[st]);
                        bulk_g2s(dst + kTcMainBytes, p.tc_aug + 2 * aug_part + tg * kTcAugGroupBytes, kTcAugBytes, &fullT[st]);
                        bulk_g2s(dst + kTcMainBytes + kTcAugBytes, p.tc_aug + 3 * aug_part + tg * kTcAugGroupBytes, kTcAugBytes, &fullT[st]);
                        }
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
        // query augmented columns: 16 B per row ([group][8 rows][16 B], group stride 128 B).  Only k-chunk 0 exists; the
        // descriptor's k-chunk 1 aliases the NEXT group's rows, which is harmless because the train side's k-chunk 1 is all
        // zeros and every aliased value is finite.
        const uint64_t qad = tc_desc_nosw(smem_u32(Qa), 128, 128);
        const uint64_t td0 = tc_desc_sw128(smem_u32(Ts), kTcGroupBytes);
        const uint64_t tad0 = tc_desc_nosw(smem_u32(Ts) + kTcMainBytes, 128, kTcAugGroupBytes);
        uint32_t g = 0, qseq = 0;
        for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
            const TcUnit u = tc_decode_unit(p, unit);
            for (int qb = u.qb0; qb < u.qb1; ++qb) {
                mbar_wait_relaxed(fullQ, qseq & 1);
                ++qseq;
                for (int tt = 0; tt < u.ntt; ++tt, ++g) {
                    const uint32_t st = g % kTcStages, ph = (g / kTcStages) & 1;
                    const uint32_t as = g % kTcAccStages, aph = (g / kTcAccStages) & 1;
                    mbar_wait_relaxed(&fullT[st], ph);
                    mbar_wait_relaxed(&accEmpty[as], aph ^ 1);
                    tc_fence_after();
                    // descriptor start addresses are in 16-byte units: adding (bytes >> 4) to the low word moves the window
                    const uint64_t td = td0 + (uint64_t)(st * (kTcTileBytes >> 4));
                    const uint64_t tad = tad0 + (uint64_t)(st * (kTcTileBytes >> 4));
                    const uint32_t d = tmem + as * 128;
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
                                tc_mma_tf32_ts(d, tmem + (pa ? kTcAColLo : kTcAColHi) + ks * 8, td + (uint64_t)(pb * (2048 >> 4) + off), idesc, !first);
                                first = false;
                            }
                            tc_mma_tf32(d, qad + (uint64_t)(pa * (kTcQaBytes >> 4)), tad + (uint64_t)(pb * (kTcAugBytes >> 4)), idesc, true);
                        }
                    }
                    __syncwarp();
                    if (elect_one()) {
                        tc_commit(&emptyT[st]);     // shared-memory stage consumed once these MMAs retire
                        tc_commit(&accFull[as]);    // accumulator stage ready for the epilogue
                    }
                    __syncwarp();
                }
                if (elect_one()) tc_commit(emptyQ);
                __syncwarp();
            }
        }
    } else if (warp >= kTcWriterWarp0) {
        // ======================= query writers: fp32 rows -> (hi, lo) TF32 operand in tensor memory =======================
        const int quarter = warp & 3;
        const int trow = quarter * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        uint32_t qseq = 0;
        for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
            const TcUnit u = tc_decode_unit(p, unit);
            const int fq = p.frame_rows[u.q_frame];
            const float4* qrows = reinterpret_cast<const float4*>(p.rows_f32 + (size_t)p.frame_row_off[u.q_frame] * kDim);
            for (int qb = u.qb0; qb < u.qb1; ++qb) {
                const int r = qb * kTile + trow;
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
                const float hqh = tc_tf32_hi(hq);
                mbar_wait(emptyQ, (qseq & 1) ^ 1);      // every MMA reading the previous query operand has retired
                ++qseq;
                tc_fence_after();
                *reinterpret_cast<float4*>(Qa + trow * 16) = make_float4(1.f, hqh, 0.f, 0.f);
                *reinterpret_cast<float4*>(Qa + kTcQaBytes + trow * 16) = make_float4(0.f, hq - hqh, 0.f, 0.f);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the MMA's async reads
#pragma unroll
                for (int m = 0; m < 8; ++m) {
                    const float xs[8] = {x[2 * m].x, x[2 * m].y, x[2 * m].z, x[2 * m].w, x[2 * m + 1].x, x[2 * m + 1].y, x[2 * m + 1].z, x[2 * m + 1].w};
                    uint32_t hi[8], lo[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float h = tc_tf32_hi(xs[j]);
                        hi[j] = __float_as_uint(h);
                        lo[j] = __float_as_uint(xs[j] - h);
                    }
                    tmem_st8(tmem + lane_addr + kTcAColHi + m * 8, hi);
                    tmem_st8(tmem + lane_addr + kTcAColLo + m * 8, lo);
                }
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(fullQ);
            }
        }
    } else {
        // ======================= epilogue warps =======================
        const int quarter = warp & 3, part = warp >> 2;
        const int trow = quarter * 32 + lane;               // row inside the 128-row query tile (= TMEM lane)
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        if (part == 0) { sbound[trow] = kTcBoundBits; mkey[trow] = kKeyInit; mkey[kTile + trow] = kKeyInit; }
        if (part == 1 && trow < 8) {      // the 128-byte pads behind the two augmented blocks (aliased k-chunk of the last group)
            *reinterpret_cast<float4*>(Qa + 16 * 128 + trow * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<float4*>(Qa + kTcQaBytes + 16 * 128 + trow * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kTcEpiThreads) : "memory");
        uint32_t g = 0;
        for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
            const TcUnit u = tc_decode_unit(p, unit);
            u64* rk1 = p.keys + (size_t)u.pair * 4 * p.stride;
            u64* rk2 = rk1 + p.stride;
            u64* ck1 = rk2 + p.stride;
            uint32_t* tauc = p.col_thr + (size_t)u.pair * p.stride;
            for (int qb = u.qb0; qb < u.qb1; ++qb) {
                const uint32_t qrow = (uint32_t)(qb * kTile + trow);     // frame row of this thread
                RowTop2 t;
                t.v1 = t.v2 = __uint_as_float(kTcBoundBits);
                t.i1 = t.i2 = 0xffffffffu;
                for (int tt = 0; tt < u.ntt; ++tt, ++g) {
                    const uint32_t as = g % kTcAccStages, aph = (g / kTcAccStages) & 1;
                    const uint32_t ts = g % kTcThrStages, tph = (g / kTcThrStages) & 1;
                    mbar_wait(&thrFull[ts], tph);
                    mbar_wait(&accFull[as], aph);
                    tc_fence_after();
                    const float4* tp = reinterpret_cast<const float4*>(Thr + ts * kTcThrBytes) + part * (kTcPartCols / 4);
                    const uint32_t taddr = tmem + lane_addr + as * 128 + part * kTcPartCols;
                    // The row's running second best over ALL column parts (each part keeps a private top-2; the shared
                    // bound only filters, with '>=' so equal values still reach the private strict-'<' insertion).
                    float nb = -__uint_as_float(*reinterpret_cast<volatile uint32_t*>(&sbound[trow]));
                    // columns in chunks of 16: a real loop, so the epilogue body stays small enough for the instruction
                    // cache (a fully unrolled 64-column body was 64 KB of SASS and stalled on instruction fetch)
#pragma unroll 1
                    for (int ch = 0; ch < kTcPartCols / 16; ++ch) {
                        uint32_t vb[16];
                        tmem_ld16(taddr + ch * 16, vb);
                        float thr[16];
#pragma unroll
                        for (int m = 0; m < 4; ++m) {
                            const float4 x = tp[ch * 4 + m];
                            thr[4 * m] = x.x; thr[4 * m + 1] = x.y; thr[4 * m + 2] = x.z; thr[4 * m + 3] = x.w;
                        }
                        tmem_ld_wait();
                        if (ch == kTcPartCols / 16 - 1) {      // everything this warp needs from the two rings is in registers
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) { mbar_arrive(&accEmpty[as]); mbar_arrive(&thrEmpty[ts]); }
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
                                atomicMin(&sbound[trow], __float_as_uint(fmaxf(t.v2, 0.f)));
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
                }
                // ---- end of the sweep for this query block: merge the column parts of every row, publish ----
                // 64-bit shared-memory atomics on packed keys: the smallest key ends in mkey[0], the smallest of all the
                // "losers" (displaced old minimum, or the newcomer if it did not win) in mkey[1] = the second smallest overall.
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const uint32_t idx = e ? t.i2 : t.i1;
                    if (idx != 0xffffffffu) {
                        const u64 k = make_key(__float_as_uint(fmaxf(e ? t.v2 : t.v1, 0.f)), idx);
                        const u64 old = atomicMin(&mkey[trow], k);
                        atomicMin(&mkey[kTile + trow], old > k ? old : k);
                    }
                }
                asm volatile("bar.sync 1, %0;" ::"n"(kTcEpiThreads) : "memory");
                if (part == 0) {
                    rk1[qrow] = mkey[trow];
                    rk2[qrow] = mkey[kTile + trow];
                    mkey[trow] = kKeyInit;
                    mkey[kTile + trow] = kKeyInit;
                    sbound[trow] = kTcBoundBits;
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
    return 1024 + (size_t)2 * kTcQaBytes + (size_t)kTcStages * kTcTileBytes + (size_t)kTcThrStages * kTcThrBytes + 2 * kTile * sizeof(u64) + kTile * 4 +
           (2 + 2 * kTcStages + 2 * kTcAccStages + 2 * kTcThrStages) * 8 + 16;
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

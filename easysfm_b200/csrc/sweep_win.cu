// sweep_win.cu -- round-2 tensor-core sweeps with WINDOW keys (sm_100a): half the tensor-pipe time and half the operand bytes of
// sweep_l2_tc.cu, and a selection epilogue that carries no column indices.
//
// Same contract towards the caller as the other sweeps (cpp_code/src/feature_matching.cpp:71-92 / :115-137 as exact brute force,
// python_code/feature_match.py:26-27,33-39): per image pair the two nearest train rows of every query row and the nearest query
// row of every train row, with OpenCV's lowest-index tie-break.  What differs is HOW a row's neighbours are reported:
//     row key    = value bits << 32 | SLICE,    slice = (first column / 8) << 1 | wide: the 8 (SURF, best value) or 32 train columns
//                                               the value was found in
//     column key = value bits << 32 | query row                                                    (unchanged)
// finalize.cu finds the column inside the window by re-evaluating its 32 distances exactly (direct-form L2 / XOR + POPC), which it
// only does for the rows that can pass the ratio test.  The epilogue therefore needs VALUES only:
//   * SURF (KIND = ESFM_KIND_F32X64, "H" split of tc_layout.cuh): 12 x tcgen05.mma.kind::f16 (a = fp16(x), b = fp16(x - a);
//     b_q.a_t + a_q.b_t + a_q.a_t, fp32 accumulation; fp16 subnormals are honoured by the tensor core: h16_probe.cu) + the exact
//     kind::tf32 half-norm step = 13 MMA slots per 128 x 128 tile instead of 25, tile image 36 KB instead of 68 KB.  Epilogue:
//     a branch-free tournament over the 32 accumulators of a pass gives the pass's two largest -1/2 d^2, eleven selects merge them
//     into the running (k1, w1, k2, w2).
//   * ORB (KIND = ESFM_KIND_B256): the +-1 FP8 operands of sweep_l2_tc.cu with FP16 ACCUMULATORS (-2 hamming, exact; pad rows
//     overflow to -inf): tcgen05.ld.pack::16b delivers two columns per register and the whole epilogue runs on packed halves
//     (HMNMX2 / VHMNMX tournament, HFMA2.RELU threshold tests): ~1/2 the instructions per element of the fp32 epilogue.
// Column minima keep the running-threshold scheme of sweep_l2_tc.cu (tc_sweep_common.cuh).
//
// Pipeline: one persistent CTA per SM, 24 warps -- TMA producer (5-stage ring of 36 KB tile images + threshold snapshots),
// MMA issuer (query operand in tensor memory, 3 accumulator stages), 4 query writers, 16 epilogue warps (warp w: TMEM lane quarter
// w % 4 = 32 query rows, column part w / 4 = 32 train columns of every tile).
#include <cuda_fp16.h>

#include "tc_sweep_common.cuh"

namespace esfm {

namespace {

constexpr int kWinStages = 5;                       // shared-memory train stages (36 KB each)
// MMA issuers: the tcgen05.mma queue is shallow, so the tensor pipe runs dry while ONE issuer thread does a tile's book-keeping (two barrier
// waits, two commits: ~500 clk against 577 / 833 clk of MMAs per tile; csrc/microbench/pipe_probe.cu).  Three warps take the tiles in turn:
// while one is blocked issuing its MMAs the next has already passed its waits.  Every barrier must keep ONE issuer (a parity wait may not
// skip a phase): the accumulator ring has 3 stages = 3 issuers; the load ring indexes its barriers by tile % lcm(stages, issuers).
constexpr int kWinIssuers = 3;
constexpr int kWinLoadBars = 15;                    // lcm(kWinStages, kWinIssuers)
static_assert(kWinLoadBars % kWinStages == 0 && kWinLoadBars % kWinIssuers == 0, "load-ring barriers");
constexpr int kWinAccStages = 3;                    // (512 - 64) / 128 accumulator stages
constexpr uint32_t kWinACol0 = 128u * kWinAccStages; // query operand: 64 tensor-memory columns (SURF: a 32 | b 32; ORB: 256 FP8)
constexpr int kWinTileBytes = kTchTileBytes;        // == kTc8TileBytes
constexpr int kWinMainBytes = kTchMainBytes;
constexpr int kWinGroupBytes = kTchGroupBytes;
static_assert(kTchTileBytes == kTc8TileBytes && kTchMainBytes == kTc8MainBytes && kTchGroupBytes == kTc8GroupBytes, "the two kinds share one geometry");
constexpr float kWinNone = -3.0e38f;                // "no candidate": below every real value and every pad
// ORB thresholds are fp16 "S" values: a column is hit iff 2 hamming < S  <=>  v + S > 0 (v = -2 hamming).  Start: 560 (0x6060,
// a repeated byte for cudaMemsetAsync) > 512 >= every real 2 * hamming; pads are -inf.
constexpr int kWinThrBytesB256 = kTile * 2;

__device__ __forceinline__ uint32_t hmax2(uint32_t a, uint32_t b) { uint32_t r; asm("max.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t hmin2(uint32_t a, uint32_t b) { uint32_t r; asm("min.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
// relu(v + s) per half: > 0 iff the column beats its threshold
__device__ __forceinline__ uint32_t hadd2_relu(uint32_t v, uint32_t s) {
    uint32_t r;
    asm("fma.rn.relu.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(v), "r"(0x3c003c00u), "r"(s));
    return r;
}
__device__ __forceinline__ float h_lo(uint32_t x) { return __low2float(*reinterpret_cast<const __half2*>(&x)); }
__device__ __forceinline__ float h_hi(uint32_t x) { return __high2float(*reinterpret_cast<const __half2*>(&x)); }

// One register = two adjacent ORB columns (2 j, 2 j + 1): every half that beat its threshold (e half != 0) posts its key
// (float bits of 2 * hamming, query row) with RED.MIN.64, and ONE vector RED.MIN (two fp16) lowers both thresholds (+inf in a half
// leaves it alone).  s_add: 0 = the published threshold excludes equal distances (one CTA per pair: a snapshot only holds minima of
// lower query rows), 1.0 in both halves = equal distances still get through (pair split over several CTAs).
template <int J>
__device__ __forceinline__ void win_col_post2(uint32_t v, uint32_t e, u64* ck, unsigned short* tau, uint32_t qrow, uint32_t s_add) {
    if (e == 0u) return;
    const float y0 = fabsf(h_lo(v)), y1 = fabsf(h_hi(v));      // v <= 0; |v| also turns a -0 into +0 (keys order as unsigned bits)
    uint32_t sn;                                                 // new thresholds: 2 hamming (+ s_add)
    asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(sn) : "r"(v), "r"(0xbc00bc00u), "r"(s_add));
    const bool p0 = (e & 0xffffu) != 0u, p1 = (e >> 16) != 0u;
    const uint32_t s0 = p0 ? (sn & 0xffffu) : 0x7c00u, s1 = p1 ? (sn >> 16) : 0x7c00u;
    if (p0) atomicMin(ck + 2 * J, make_key(__float_as_uint(y0), qrow));
    if (p1) atomicMin(ck + 2 * J + 1, make_key(__float_as_uint(y1), qrow));
    asm volatile("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %1;\n\tred.global.v2.f16.min.noftz [%0], {lo, hi};\n\t}" ::"l"(tau + 2 * J), "r"(s0 | (s1 << 16)) : "memory");
}

// The many-rows case (first query block of a pair): 4 columns parked in the warp's scratch as floats y = 2 * hamming, one segmented
// warp reduction (8 lanes per column), the 4 leaders post.  Thresholds are fp16 S values at thr_addr (shared window), 2 bytes each.
__device__ __forceinline__ void win_col_group_h(uint32_t sc_addr, uint32_t thr_addr, u64* ck, unsigned short* tau, uint32_t qrow0, float s_add) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t col = lane >> 3, r4 = (lane & 7u) * 4u;
    float y0, y1, y2, y3;
    unsigned short sh;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(y0), "=f"(y1), "=f"(y2), "=f"(y3) : "r"(sc_addr + col * 128u + r4 * 4u));
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(sh) : "r"(thr_addr + col * 2u));
    const float s = __half2float(__ushort_as_half(sh));
    float best = y0;
    uint32_t br = r4;
    if (y1 < best) { best = y1; br = r4 + 1; }
    if (y2 < best) { best = y2; br = r4 + 2; }
    if (y3 < best) { best = y3; br = r4 + 3; }
#pragma unroll
    for (int d = 1; d <= 4; d <<= 1) {
        const float oy = __shfl_xor_sync(0xffffffffu, best, d);
        const uint32_t orow = __shfl_xor_sync(0xffffffffu, br, d);
        const bool take = oy < best || (oy == best && orow < br);
        best = take ? oy : best;
        br = take ? orow : br;
    }
    if ((lane & 7u) == 0 && best < s) {
        atomicMin(ck + col, make_key(__float_as_uint(best), qrow0 + br));
        const uint32_t sn = (uint32_t)__half_as_ushort(__float2half_rn(best + s_add));
        const uint32_t pk = (col & 1u) ? (0x7c00u | (sn << 16)) : (sn | 0x7c000000u);
        asm volatile("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %1;\n\tred.global.v2.f16.min.noftz [%0], {lo, hi};\n\t}" ::"l"(tau + (col & 2u)), "r"(pk) : "memory");
    }
}

// ---- ORB: compact, never-inlined event handlers: ONE copy of each in the instruction stream (the unrolled per-chain copies were 23 KB
//      of rarely executed code; "no instruction" was 31 % of the stalls inside them, profiles/ncu_orb_tc16_r2.txt).  The SURF epilogue
//      keeps its handlers inline: a call while 32 accumulators are live costs more in register moves than the code size saves
//      (measured: 22.8 -> 24.2 ms) ----
// ORB: one chain of 8 columns = 4 registers of the calling lane (only lanes with a hit call this)
__device__ __noinline__ void win_chain_post_h(uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t thr_addr, u64* ck, unsigned short* tau,
                                              uint32_t qrow, uint32_t s_add) {
    uint32_t s0, s1, s2, s3;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(s0), "=r"(s1), "=r"(s2), "=r"(s3) : "r"(thr_addr));
    win_col_post2<0>(x0, hadd2_relu(x0, s0), ck, tau, qrow, s_add);
    win_col_post2<1>(x1, hadd2_relu(x1, s1), ck, tau, qrow, s_add);
    win_col_post2<2>(x2, hadd2_relu(x2, s2), ck, tau, qrow, s_add);
    win_col_post2<3>(x3, hadd2_relu(x3, s3), ck, tau, qrow, s_add);
}
// ORB: one group of 4 columns = 2 registers of EVERY lane (whole warp): park, reduce, post
__device__ __noinline__ void win_group_h(uint32_t x0, uint32_t x1, uint32_t sc_addr, uint32_t thr_addr, u64* ck, unsigned short* tau, uint32_t qrow0, float s_add) {
    const uint32_t lane = threadIdx.x & 31u;
    const float y[4] = {fabsf(h_lo(x0)), fabsf(h_hi(x0)), fabsf(h_lo(x1)), fabsf(h_hi(x1))};      // pads: +inf, never a winner
#pragma unroll
    for (int c = 0; c < 4; ++c) asm volatile("st.shared.f32 [%0], %1;" ::"r"(sc_addr + (uint32_t)(c * 128) + lane * 4u), "f"(y[c]) : "memory");
    __syncwarp();
    win_col_group_h(sc_addr, thr_addr, ck, tau, qrow0, s_add);
    __syncwarp();
}
// merge the pass's two largest values (p1 >= p2, found in the column slices id1 / id2) into the running (k1, w1, k2, w2): strict
// comparisons, so of equal values the one found first (the lower slice = the lower column) stays in front
__device__ __forceinline__ void win_merge(float p1, int id1, float p2, int id2, float& k1, int& w1, float& k2, int& w2) {
    const bool c1 = p1 > k1;
    const float nk2 = fmaxf(c1 ? p2 : k2, fminf(k1, p1));
    const int wa = p2 > k1 ? id2 : w1;           // (c1)  second = p2 (this pass) or the demoted old best
    const int wb = p1 > k2 ? id1 : w2;           // (!c1) second = p1 (this pass) or unchanged
    w2 = c1 ? wa : wb;
    k2 = nk2;
    w1 = c1 ? id1 : w1;
    k1 = fmaxf(k1, p1);
}

// work decoding: a normal sweep takes (query frame, train frame) from the pair list; the verification sweep of the two-phase
// cross-check swaps the roles and reads its query rows through the pair's gather list
struct WinUnit {
    int pair, q_frame, t_frame;
    int fq;              // query rows of this unit's pair (gathered rows in a verification sweep)
    int ntt;
    int qb0, qb1;        // query TILES [qb0, qb1) of this unit
};
__device__ __forceinline__ WinUnit win_decode_unit(const SweepParams& p, int unit) {
    WinUnit u;
    u.pair = unit / p.units_per_pair;
    const int part = unit - u.pair * p.units_per_pair;
    const PairDesc pd = p.pairs[u.pair];
    int nqt;
    if (p.gather) {
        u.q_frame = pd.t_frame;
        u.t_frame = pd.q_frame;
        u.fq = p.gather_cnt[u.pair];
        nqt = (u.fq + kTile - 1) / kTile;
    } else {
        u.q_frame = pd.q_frame;
        u.t_frame = pd.t_frame;
        u.fq = p.frame_rows[pd.q_frame];
        nqt = p.frame_tile_off[pd.q_frame + 1] - p.frame_tile_off[pd.q_frame];
    }
    u.ntt = p.frame_tile_off[u.t_frame + 1] - p.frame_tile_off[u.t_frame];
    u.qb0 = (int)((long long)nqt * part / p.units_per_pair);
    u.qb1 = (int)((long long)nqt * (part + 1) / p.units_per_pair);
    if (p.frame_rows[u.t_frame] < 1) u.qb1 = u.qb0;
    return u;
}

}  // namespace

template <int KIND>
__global__ void __launch_bounds__(kTcThreads, 1) sweep_win_kernel(const SweepParams p) {
    constexpr bool kOrb = KIND != ESFM_KIND_F32X64;
    constexpr int kThrBytes = kOrb ? kWinThrBytesB256 : kTcThrBytes;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);       // SWIZZLE_128B atoms: 1024-byte alignment
    unsigned char* Ts = base;                                   // kWinStages train tile images (36 x 1024 B each)
    unsigned char* Qa = Ts + kWinStages * kWinTileBytes;        // augmented block of the query tile
    unsigned char* Thr = Qa + kTcAugBytes;                      // kTcThrStages threshold snapshots
    u64* mkey = reinterpret_cast<u64*>(Thr + kTcThrStages * kTcThrBytes);          // [best, second][128] merged row keys
    float* colsc = reinterpret_cast<float*>(mkey + 2 * kTile);                    // [epilogue warp][4 columns][32 lanes] column-event scratch
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(colsc) + kTcScBytes);
    uint64_t* fullQ = bars;
    uint64_t* fullT = fullQ + 1;
    uint64_t* emptyT = fullT + kWinLoadBars;
    uint64_t* accFull = emptyT + kWinLoadBars;
    uint64_t* accEmpty = accFull + kWinAccStages;
    uint64_t* thrFull = accEmpty + kWinAccStages;
    uint64_t* thrEmpty = thrFull + kTcThrStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(thrEmpty + kTcThrStages);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_units = p.n_pairs * p.units_per_pair;
    const bool nosleep = (p.debug_flags & 128) != 0;      // probe: the producer / issuer poll their barriers without sleeping

    if (threadIdx.x == 0) {
        mbar_init(&fullQ[0], 4);                      // the 4 query-writer warps
        for (int s = 0; s < kWinLoadBars; ++s) {
            mbar_init(&fullT[s], 1);
            mbar_init(&emptyT[s], 1);                 // MMA commit
        }
        for (int s = 0; s < kTcThrStages; ++s) {
            mbar_init(&thrFull[s], 1);
            mbar_init(&thrEmpty[s], kTcEpiWarps);
        }
        for (int s = 0; s < kWinAccStages; ++s) {
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

    if (warp >= kTcEpiWarps) {
    reg_dealloc<kTcServiceRegs>();
    if (warp == kTcEpiWarps) {
        // ======================= TMA producer =======================
        if (lane == 0) {
            uint32_t g = 0;
            for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
                const WinUnit u = win_decode_unit(p, unit);
                const unsigned char* timg = p.tc_main + (size_t)p.frame_tile_off[u.t_frame] * kWinTileBytes;
                const unsigned char* tauc = reinterpret_cast<const unsigned char*>(p.col_thr) + (size_t)u.pair * p.stride * (kOrb ? 2 : 4);
                for (int qt = u.qb0; qt < u.qb1; ++qt) {
                    for (int tt = 0; tt < u.ntt; ++tt, ++g) {
                        // stage g % 5; its barrier pair g % 15; the stage was last used by tile g - 5, whose barrier is (g - 5) % 15
                        const uint32_t st = g % kWinStages, bi = g % kWinLoadBars, ph = (g / kWinLoadBars) & 1;
                        if (g >= kWinStages) {
                            const uint32_t pg = g - kWinStages, pbi = pg % kWinLoadBars, pph = (pg / kWinLoadBars) & 1;
                            if (nosleep) mbar_wait_sleep<0>(&emptyT[pbi], pph); else mbar_wait_sleep<kTcSleepProducer>(&emptyT[pbi], pph);
                        }
                        (void)ph;
                        mbar_arrive_expect_tx(&fullT[bi], kWinTileBytes);
                        if (p.debug_flags & 4)
                            asm volatile("mbarrier.complete_tx.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&fullT[bi])), "r"(kWinTileBytes) : "memory");
                        else
                            bulk_g2s(Ts + (size_t)st * kWinTileBytes, timg + (size_t)tt * kWinTileBytes, kWinTileBytes, &fullT[bi]);
                        // the running column thresholds of this tile ride along in their own ring (a stale snapshot is only looser)
                        if (p.need_cols) {
                            const uint32_t ts = g % kTcThrStages, tph = (g / kTcThrStages) & 1;
                            if (nosleep) mbar_wait_sleep<0>(&thrEmpty[ts], tph ^ 1); else mbar_wait_sleep<kTcSleepProducer>(&thrEmpty[ts], tph ^ 1);
                            mbar_arrive_expect_tx(&thrFull[ts], kThrBytes);
                            bulk_g2s(Thr + ts * kTcThrBytes, tauc + (size_t)tt * kThrBytes, kThrBytes, &thrFull[ts]);
                        }
                    }
                }
            }
        }
    } else if (warp >= kTcEpiWarps + 1 && warp <= kTcEpiWarps + kWinIssuers) {
        // ======================= MMA issuers (warp-uniform control flow, one elected lane issues): tile g belongs to issuer g % 3 =======================
        const uint32_t me = (uint32_t)(warp - (kTcEpiWarps + 1));
        constexpr uint32_t id_main = kOrb ? tc_idesc_e4m3_h(128, 128) : tc_idesc_f16(128, 128, 0, 0);
        constexpr uint32_t id_aug = kOrb ? tc_idesc_e4m3_h(128, 128) : tc_idesc_tf32(128, 128);
        const uint64_t qad = tc_desc_nosw(smem_u32(Qa), 128, kTcAugGroupBytes);
        const uint64_t td0 = tc_desc_sw128(smem_u32(Ts), kWinGroupBytes);
        const uint64_t tad0 = tc_desc_nosw(smem_u32(Ts) + kWinMainBytes, 128, kTcAugGroupBytes);
        uint32_t g = 0, qn = 0;
        for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
            const WinUnit u = win_decode_unit(p, unit);
            for (int qt = u.qb0; qt < u.qb1; ++qt, ++qn) {
                // every issuer observes every fill of the query slot (a parity wait may not skip a phase), also in a block none of whose
                // tiles are its own
                mbar_wait_sleep<0>(&fullQ[0], qn & 1);
                for (int tt = 0; tt < u.ntt; ++tt, ++g) {
                    if (g % kWinIssuers != me) continue;
                    const uint32_t st = g % kWinStages, bi = g % kWinLoadBars, ph = (g / kWinLoadBars) & 1;
                    if (nosleep) mbar_wait_sleep<0>(&fullT[bi], ph); else mbar_wait_sleep<kTcSleepIssuer>(&fullT[bi], ph);
                    const uint64_t td = td0 + (uint64_t)(st * (kWinTileBytes >> 4));
                    const uint64_t tad = tad0 + (uint64_t)(st * (kWinTileBytes >> 4));
                    const uint32_t as = g % kWinAccStages, aph = (g / kWinAccStages) & 1;
                    if (nosleep) mbar_wait_sleep<0>(&accEmpty[as], aph ^ 1); else mbar_wait_sleep<kTcSleepIssuer>(&accEmpty[as], aph ^ 1);
                    tc_fence_after();
                    const uint32_t d = tmem + as * 128;
                    const uint32_t acol = tmem + kWinACol0;
                    if (!(p.debug_flags & 2) && elect_one()) {
                        if (kOrb) {
#pragma unroll
                            for (int ks = 0; ks < 8; ++ks)          // 256 FP8 values per row = 8 k-steps of K = 32 (32 bytes each)
                                tc_mma_f8_ts(d, acol + ks * 8, td + (uint64_t)(((ks >> 2) * 1024 + (ks & 3) * 32) >> 4), id_main, ks > 0);
                            // - 256 and the pad-row penalties: only the column thresholds need them (rows-only sweeps mask the pad
                            // columns of a frame's last tile in the epilogue and report 256 - v: one MMA slot of nine saved)
                            if (p.need_cols) tc_mma_f8(d, qad, tad, id_aug, true);
                        } else {
                            // small terms first: b_q.a_t, a_q.b_t, then a_q.a_t; a k-step = 16 halves = 32 bytes of the 128-byte row
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) tc_mma_f16_ts(d, acol + 32 + ks * 8, td + (uint64_t)((ks * 32) >> 4), id_main, ks > 0);
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) tc_mma_f16_ts(d, acol + ks * 8, td + (uint64_t)((1024 + ks * 32) >> 4), id_main, true);
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) tc_mma_f16_ts(d, acol + ks * 8, td + (uint64_t)((ks * 32) >> 4), id_main, true);
                            tc_mma_tf32(d, qad, tad, id_aug, true);   // - 1/2|q|^2 - 1/2|t|^2, exact
                        }
                    }
                    __syncwarp();
                    if (elect_one()) {
                        tc_commit(&accFull[as]);        // accumulator stage ready for the epilogue
                        tc_commit(&emptyT[bi]);         // shared-memory stage consumed
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp >= kTcWriterWarp0) {
        // ======================= query writers: rows of the query tile -> operand columns of tensor memory =======================
        const int quarter = warp & 3;
        const int trow = quarter * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        const uint32_t acol = tmem + lane_addr + kWinACol0;
        unsigned char* qa = Qa + (trow >> 3) * kTcAugGroupBytes + (trow & 7) * 16;
        uint32_t quse = 0;
        for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
            const WinUnit u = win_decode_unit(p, unit);
            const int fq = u.fq;
            const int* gl = p.gather ? p.gather + (size_t)u.pair * p.stride : nullptr;
            for (int qt = u.qb0; qt < u.qb1; ++qt) {
                const int slot = qt * kTile + trow;
                const bool valid = slot < fq;
                const int r = (gl && valid) ? __ldg(gl + slot) : slot;      // verification sweep: the slot's row of the (original) train frame
                if (kOrb) {
                    const uint4* qbits = p.rows_b256 + (size_t)p.frame_row_off[u.q_frame] * 2;
                    uint4 w0 = make_uint4(0u, 0u, 0u, 0u), w1 = w0;
                    if (valid) { w0 = __ldg(qbits + (size_t)r * 2); w1 = __ldg(qbits + (size_t)r * 2 + 1); }
                    // The slot is free once every MMA that read the previous tile has retired: the writers block on NAMED BARRIER 2,
                    // which epilogue warp 0 arrives at when it has seen the block's last accumulator complete.
                    if (quse > 0) named_bar_sync(2, 128 + 32);
                    ++quse;
                    tc_fence_after();
                    *reinterpret_cast<uint4*>(qa) = make_uint4((valid ? 0u : kFp8Pos448) | (kFp8Pos448 << 8) | (kFp8Pos16 << 16), 0u, 0u, 0u);
                    *reinterpret_cast<uint4*>(qa + 128) = make_uint4(0u, 0u, 0u, 0u);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    const uint32_t ws[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                    for (int m = 0; m < 8; ++m) {       // word m = elements 32 m .. 32 m + 31 = 8 columns
                        uint32_t c[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) c[j] = valid ? tc8_expand4(ws[m] >> (4 * j)) : 0u;
                        tmem_st8(acol + m * 8, c);
                    }
                } else {
                    const float4* xr = reinterpret_cast<const float4*>(p.rows_f32 + (size_t)p.frame_row_off[u.q_frame] * kDim) + (size_t)(valid ? r : 0) * 16;
                    // pass 1: 1/2|q|^2 in the summation order of bank.cu's pack kernels
                    float hs = 0.f;
#pragma unroll 4
                    for (int m = 0; m < 16; ++m) {
                        const float4 x = __ldg(xr + m);
                        hs = __fmaf_rn(x.x, x.x, hs); hs = __fmaf_rn(x.y, x.y, hs);
                        hs = __fmaf_rn(x.z, x.z, hs); hs = __fmaf_rn(x.w, x.w, hs);
                    }
                    const float hq = valid ? 0.5f * hs : kTcPadNorm;       // pad rows can never win a column
                    float hqh, hqm, hql;
                    tc_split3(hq, hqh, hqm, hql);
                    if (quse > 0) named_bar_sync(2, 128 + 32);
                    ++quse;
                    tc_fence_after();
                    *reinterpret_cast<float4*>(qa) = make_float4(1.f, 1.f, 1.f, hqh);
                    *reinterpret_cast<float4*>(qa + 128) = make_float4(hqm, hql, 0.f, 0.f);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    // pass 2: the row again (L1/L2 hit), 16 dims at a time -> 8 columns of a (fp16 pairs) and 8 of b
#pragma unroll 1
                    for (int m = 0; m < 4; ++m) {
                        uint32_t ca[8], cb[8];
#pragma unroll
                        for (int h = 0; h < 4; ++h) {
                            float4 x = __ldg(xr + 4 * m + h);
                            if (!valid) x = make_float4(0.f, 0.f, 0.f, 0.f);
                            const __half2 a0 = __floats2half2_rn(x.x, x.y), a1 = __floats2half2_rn(x.z, x.w);
                            const float2 f0 = __half22float2(a0), f1 = __half22float2(a1);
                            const __half2 b0 = __floats2half2_rn(x.x - f0.x, x.y - f0.y), b1 = __floats2half2_rn(x.z - f1.x, x.w - f1.y);
                            ca[2 * h] = *reinterpret_cast<const uint32_t*>(&a0); ca[2 * h + 1] = *reinterpret_cast<const uint32_t*>(&a1);
                            cb[2 * h] = *reinterpret_cast<const uint32_t*>(&b0); cb[2 * h + 1] = *reinterpret_cast<const uint32_t*>(&b1);
                        }
                        tmem_st8(acol + m * 8, ca);
                        tmem_st8(acol + 32 + m * 8, cb);
                    }
                }
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&fullQ[0]);
            }
        }
        if (quse > 0) named_bar_sync(2, 128 + 32);      // the epilogue arrives once per block; the last arrival has no refill waiting
    }
    } else {
        // ======================= epilogue warps =======================
        reg_alloc<kTcEpiRegs>();
        const int quarter = warp & 3, part = warp >> 2;
        const int trow = quarter * 32 + lane;               // row inside the 128-row query tile (= TMEM lane)
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        if (part == 0) {
            mkey[trow] = kKeyInit;
            mkey[kTile + trow] = kKeyInit;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kTcEpiThreads) : "memory");
        const uint32_t sc_addr = smem_u32(colsc) + (uint32_t)warp * (kTcScCols * 32 * 4);
        uint32_t g = 0;
        for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
            const WinUnit u = win_decode_unit(p, unit);
            u64* rk1 = p.keys + (size_t)u.pair * 4 * p.stride;
            u64* rk2 = rk1 + p.stride;
            u64* ck1 = rk2 + p.stride;
            for (int qt = u.qb0; qt < u.qb1; ++qt) {
                const uint32_t qrow = (uint32_t)(qt * kTile + trow);
                const uint32_t qrow0 = qrow - (uint32_t)lane;
                const int ft = p.frame_rows[u.t_frame];
                const float voff = (kOrb && !p.need_cols) ? 256.f : 0.f;       // ORB rows-only: v = 256 - 2 hamming (no augmented MMA)
                float k1 = kWinNone, k2 = kWinNone;
                int w1 = -1, w2 = -1;
                for (int tt = 0; tt < u.ntt; ++tt, ++g) {
                    const uint32_t ts = g % kTcThrStages, tph = (g / kTcThrStages) & 1;
                    if (p.need_cols) mbar_wait_sleep<kTcSleepEpilogue>(&thrFull[ts], tph);       // (no cross-check: no thresholds travel)
                    const uint32_t as = g % kWinAccStages, aph = (g / kWinAccStages) & 1;
                    mbar_wait_sleep<kTcSleepEpilogue>(&accFull[as], aph);
                    if (warp == 0 && tt == u.ntt - 1) named_bar_arrive(2, 128 + 32);     // the query slot may be refilled (see the writers)
                    tc_fence_after();
                    const int wide_id = ((tt * (kTile / 8) + part * (kTcPartCols / 8)) << 1) | 1;     // this pass's 32 columns as a slice id
                    const uint32_t col0 = (uint32_t)(tt * kTile + part * kTcPartCols);
                    if constexpr (kOrb) {
                        // ---------------- ORB: 16 registers = 32 fp16 accumulators v = -2 hamming (pads: -inf) ----------------
                        uint32_t hb[16];
                        tmem_ld16_pack(tmem + lane_addr + as * 128 + part * kTcPartCols, hb);
                        const uint4* tp = reinterpret_cast<const uint4*>(Thr + ts * kTcThrBytes + part * (kTcPartCols * 2));
                        tmem_ld_wait();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_relaxed(&accEmpty[as]);
                        if (!(p.debug_flags & 1)) {
                            if (!p.need_cols) {
                                // no augmented MMA: v = 256 - 2 hamming, and the pad rows of the train frame (all-zero operands, only in
                                // its last tile) read as 0 instead of -inf: mask them here
                                const int rem = ft - (int)col0;
                                if (rem < kTcPartCols) {
#pragma unroll
                                    for (int c = 0; c < 16; ++c)
                                        hb[c] = 2 * c + 1 < rem ? hb[c] : (2 * c < rem ? ((hb[c] & 0xffffu) | 0xfc000000u) : 0xfc00fc00u);
                                }
                            }
                            if (!(p.debug_flags & 16)) {
                                // rows: tournament on packed halves (even columns in the low halves, odd in the high ones)
                                uint32_t H[8], L[8];
#pragma unroll
                                for (int j = 0; j < 8; ++j) { H[j] = hmax2(hb[2 * j], hb[2 * j + 1]); L[j] = hmin2(hb[2 * j], hb[2 * j + 1]); }
#pragma unroll
                                for (int n = 8; n > 1; n >>= 1) {
#pragma unroll
                                    for (int j = 0; j < n / 2; ++j) {
                                        const uint32_t hh = hmax2(H[2 * j], H[2 * j + 1]);
                                        const uint32_t ll = hmax2(hmax2(hmin2(H[2 * j], H[2 * j + 1]), L[2 * j]), L[2 * j + 1]);
                                        H[j] = hh; L[j] = ll;
                                    }
                                }
                                const float a1 = h_lo(H[0]), b1 = h_hi(H[0]), a2 = h_lo(L[0]), b2 = h_hi(L[0]);
                                const float p1 = fmaxf(a1, b1), p2 = fmaxf(fmaxf(fminf(a1, b1), a2), b2);
                                win_merge(p1, wide_id, p2, wide_id, k1, w1, k2, w2);
                            }
                            if (p.need_cols) {
                                // columns: e = relu(v + S) per half; 4 chains of 8 columns; ONE vote
                                uint32_t e[16], cf[4];
#pragma unroll
                                for (int cq = 0; cq < 4; ++cq) {
                                    const uint4 s = tp[cq];
                                    e[4 * cq] = hadd2_relu(hb[4 * cq], s.x); e[4 * cq + 1] = hadd2_relu(hb[4 * cq + 1], s.y);
                                    e[4 * cq + 2] = hadd2_relu(hb[4 * cq + 2], s.z); e[4 * cq + 3] = hadd2_relu(hb[4 * cq + 3], s.w);
                                    cf[cq] = (e[4 * cq] | e[4 * cq + 1]) | (e[4 * cq + 2] | e[4 * cq + 3]);
                                }
                                const uint32_t cfa = (cf[0] | cf[1]) | (cf[2] | cf[3]);
                                const uint32_t hm = __ballot_sync(0xffffffffu, cfa != 0u);
                                if (!(p.debug_flags & 8) && hm != 0u) {
                                    // Column events: ~0.7 per pass in the steady state, 32 per pass in the first query block of a pair.
                                    u64* ckb = ck1 + col0;
                                    unsigned short* taub = reinterpret_cast<unsigned short*>(p.col_thr) + (size_t)u.pair * p.stride + col0;
                                    const uint32_t thr_addr = smem_u32(tp);
                                    const bool strict = p.units_per_pair == 1;
                                    if (__popc(hm) <= 6) {
                                        // few rows beat a threshold (the steady state): each posts its own keys, no further votes
                                        if (cfa != 0u) {
                                            const uint32_t s_add = strict ? 0u : 0x3c003c00u;
#pragma unroll
                                            for (int cq = 0; cq < 4; ++cq)
                                                if (cf[cq] != 0u)
                                                    win_chain_post_h(hb[4 * cq], hb[4 * cq + 1], hb[4 * cq + 2], hb[4 * cq + 3], thr_addr + 16u * cq, ckb + 8 * cq, taub + 8 * cq, qrow, s_add);
                                        }
                                    } else {
                                        // many rows at once (the first query block of a pair): one winner per column and warp
#pragma unroll 1
                                        for (int gq = 0; gq < 8; ++gq) {
                                            uint32_t x0 = hb[0], x1 = hb[1], cfg = cf[0];
#pragma unroll
                                            for (int k = 1; k < 8; ++k)
                                                if (gq == k) { x0 = hb[2 * k]; x1 = hb[2 * k + 1]; cfg = cf[k >> 1]; }
                                            if (__ballot_sync(0xffffffffu, cfg != 0u) == 0u) continue;
                                            win_group_h(x0, x1, sc_addr, thr_addr + 8u * gq, ckb + 4 * gq, taub + 4 * gq, qrow0, strict ? 0.f : 1.f);
                                        }
                                    }
                                }
                            }
                        }
                    } else {
                        // ---------------- SURF: 32 fp32 accumulators v = -1/2 d^2 (pads: <= -1e30) ----------------
                        uint32_t vb[32];
                        tmem_ld32(tmem + lane_addr + as * 128 + part * kTcPartCols, vb);
                        const float4* tp = reinterpret_cast<const float4*>(Thr + ts * kTcThrBytes) + part * (kTcPartCols / 4);
                        tmem_ld_wait();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_relaxed(&accEmpty[as]);
                        if (!(p.debug_flags & 1)) {
                            float v[32];
#pragma unroll
                            for (int c = 0; c < 32; ++c) v[c] = __uint_as_float(vb[c]);
                            if (!(p.debug_flags & 16)) {
                                // rows: tournament -> the pass's two largest values
                                float H[16], L[16];
#pragma unroll
                                for (int j = 0; j < 16; ++j) { H[j] = fmaxf(v[2 * j], v[2 * j + 1]); L[j] = fminf(v[2 * j], v[2 * j + 1]); }
#pragma unroll
                                for (int n = 16; n > 4; n >>= 1) {
#pragma unroll
                                    for (int j = 0; j < n / 2; ++j) {
                                        const float hh = fmaxf(H[2 * j], H[2 * j + 1]);
                                        const float ll = fmaxf(fmaxf(fminf(H[2 * j], H[2 * j + 1]), L[2 * j]), L[2 * j + 1]);
                                        H[j] = hh; L[j] = ll;
                                    }
                                }
                                // H[0..3] / L[0..3]: the two largest of each chain of 8 columns
                                const float h01 = fmaxf(H[0], H[1]), l01 = fmaxf(fmaxf(fminf(H[0], H[1]), L[0]), L[1]);
                                const float h23 = fmaxf(H[2], H[3]), l23 = fmaxf(fmaxf(fminf(H[2], H[3]), L[2]), L[3]);
                                const float p1 = fmaxf(h01, h23), p2 = fmaxf(fmaxf(fminf(h01, h23), l01), l23);
                                // the best value's chain of 8 columns (the lowest on ties): finalize evaluates 8 candidates instead of 32
                                const int sub = H[0] == p1 ? 0 : (H[1] == p1 ? 1 : (H[2] == p1 ? 2 : 3));
                                win_merge(p1, (wide_id & ~1) + 2 * sub, p2, wide_id, k1, w1, k2, w2);
                            }
                            if (p.need_cols) {
                                // columns: 4 chains of 8 threshold tests (thresholds = float bits of 1/2 d^2), ONE vote
                                bool cf[4];
#pragma unroll
                                for (int cq = 0; cq < 4; ++cq) {
                                    const float4 x0 = tp[2 * cq], x1 = tp[2 * cq + 1];
                                    cf[cq] = (v[8 * cq] >= -x0.x) | (v[8 * cq + 1] >= -x0.y) | (v[8 * cq + 2] >= -x0.z) | (v[8 * cq + 3] >= -x0.w) |
                                             (v[8 * cq + 4] >= -x1.x) | (v[8 * cq + 5] >= -x1.y) | (v[8 * cq + 6] >= -x1.z) | (v[8 * cq + 7] >= -x1.w);
                                }
                                if (!(p.debug_flags & 8) && __any_sync(0xffffffffu, cf[0] | cf[1] | cf[2] | cf[3])) {
                                    const uint32_t thr_addr = smem_u32(tp);
                                    u64* ckb = ck1 + col0;
                                    uint32_t* taub = p.col_thr + (size_t)u.pair * p.stride + col0;
                                    asm volatile("" : "+l"(ckb), "+l"(taub));
#pragma unroll
                                    for (int cq = 0; cq < 4; ++cq) {
                                        const uint32_t hm = __ballot_sync(0xffffffffu, cf[cq]);
                                        if (hm == 0) continue;
                                        if (__popc(hm) <= 4) {
                                            const float4 x0 = tp[2 * cq], x1 = tp[2 * cq + 1];
                                            u64* ckc = ckb + 8 * cq;
                                            uint32_t* tac = taub + 8 * cq;
                                            tc_col_post<0>(fmaxf(-v[8 * cq], 0.f), x0.x, 1, ckc, tac, qrow, 0.f);
                                            tc_col_post<1>(fmaxf(-v[8 * cq + 1], 0.f), x0.y, 1, ckc, tac, qrow, 0.f);
                                            tc_col_post<2>(fmaxf(-v[8 * cq + 2], 0.f), x0.z, 1, ckc, tac, qrow, 0.f);
                                            tc_col_post<3>(fmaxf(-v[8 * cq + 3], 0.f), x0.w, 1, ckc, tac, qrow, 0.f);
                                            tc_col_post<4>(fmaxf(-v[8 * cq + 4], 0.f), x1.x, 1, ckc, tac, qrow, 0.f);
                                            tc_col_post<5>(fmaxf(-v[8 * cq + 5], 0.f), x1.y, 1, ckc, tac, qrow, 0.f);
                                            tc_col_post<6>(fmaxf(-v[8 * cq + 6], 0.f), x1.z, 1, ckc, tac, qrow, 0.f);
                                            tc_col_post<7>(fmaxf(-v[8 * cq + 7], 0.f), x1.w, 1, ckc, tac, qrow, 0.f);
                                            continue;
                                        }
#pragma unroll
                                        for (int g2 = 0; g2 < 2; ++g2) {
                                            const int gq = 2 * cq + g2;
#pragma unroll
                                            for (int c = 0; c < 4; ++c)     // keys: 1/2 d^2 clamped at 0 (pad query rows: 1e30, never a winner)
                                                asm volatile("st.shared.f32 [%0], %1;" ::"r"(sc_addr + (uint32_t)(c * 128) + (uint32_t)lane * 4u), "f"(fmaxf(-v[4 * gq + c], 0.f)) : "memory");
                                            __syncwarp();
                                            tc_col_group<false>(sc_addr, thr_addr + 16u * gq, ckb + 4 * gq, taub + 4 * gq, qrow0, 0.f, p.debug_flags);
                                            __syncwarp();
                                        }
                                    }
                                }
                            }
                        }
                    }
                    __syncwarp();
                    if (p.need_cols && lane == 0) mbar_arrive_relaxed(&thrEmpty[ts]);      // last read of this threshold snapshot
                }
                // ---- end of the sweep for this query block: merge the four column parts of every row (64-bit shared-memory atomics
                //      on packed keys, "smaller = nearer": the smallest ends in mkey[0], the smallest of all the losers in mkey[1]) ----
#pragma unroll
                for (int e2 = 0; e2 < 2; ++e2) {
                    const float kv = e2 ? k2 : k1;
                    const int kw = e2 ? w2 : w1;
                    if (kw >= 0 && kv > (kOrb ? -600.f : -1.0e29f)) {          // a real column (pads: SURF <= -1e30, ORB -inf / -65504)
                        const u64 k = make_key(__float_as_uint(fmaxf(voff - kv, 0.f) + 0.f), (uint32_t)kw);   // (+ 0.f: a -0 becomes +0)
                        const u64 old = atomicMin(&mkey[trow], k);
                        atomicMin(&mkey[kTile + trow], old > k ? old : k);
                    }
                }
                asm volatile("bar.sync 1, %0;" ::"n"(kTcEpiThreads) : "memory");
                if (part == 0) {
                    if (p.gather) {
                        rk1[3 * (size_t)p.stride + qrow] = mkey[trow];      // verification sweep: the best (value, slice) of gathered row `qrow`
                    } else {
                        rk1[qrow] = mkey[trow];
                        rk2[qrow] = mkey[kTile + trow];
                    }
                    mkey[trow] = kKeyInit;
                    mkey[kTile + trow] = kKeyInit;
                }
                asm volatile("bar.sync 1, %0;" ::"n"(kTcEpiThreads) : "memory");   // mkey[] is reused by the next query block
            }
        }
    }

    // ---- teardown: everything issued has been consumed (the epilogue waited on every accumulator stage) ----
    tc_fence_before();
    __syncthreads();
    if (warp == kTcEpiWarps + 1) tmem_free(tmem, 512);
}

size_t sweep_win_smem_bytes() {
    return 1024 + (size_t)kWinStages * kWinTileBytes + kTcAugBytes + (size_t)kTcThrStages * kTcThrBytes + 2 * kTile * sizeof(u64) + (size_t)kTcScBytes +
           (1 + 2 * kWinLoadBars + 2 * kWinAccStages + 2 * kTcThrStages) * 8 + 16;
}

// kind = p.tc_kind: ESFM_KIND_F32X64 (tc_main = the H images of launch_pack_tch) or ESFM_KIND_B256 (tc_main = the +-1 FP8 images)
cudaError_t launch_sweep_win(const SweepParams& p, int sm_count, cudaStream_t s) {
    const int n_units = p.n_pairs * p.units_per_pair;
    if (n_units <= 0) return cudaSuccess;
    const int grid = n_units < sm_count ? n_units : sm_count;
    const size_t smem = sweep_win_smem_bytes();
    void (*kern)(const SweepParams) = p.tc_kind == ESFM_KIND_F32X64 ? sweep_win_kernel<ESFM_KIND_F32X64> : sweep_win_kernel<ESFM_KIND_B256>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, kTcThreads, smem, s>>>(p);
    return cudaGetLastError();
}

}  // namespace esfm

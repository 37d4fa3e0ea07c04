// orb.cu -- ORB key points and descriptors on the device: the step that PRODUCES the B256 bank.
//
// Replaces FeatureMatching::detectFeaturesORB (cpp_code/src/feature_matching.cpp:14-41): cv::ORB::create(max_num)->detect + ->compute with
// OpenCV's defaults.  The algorithm is OpenCV 4.13.0's (features2d orb.cpp / fast.cpp / fast_score.cpp / keypoint.cpp, imgproc resize.cpp
// INTER_LINEAR_EXACT, filter.simd.hpp, color_rgb, core fastAtan2); oracle/orb_oracle.py restates it stage by stage and is pinned to cv2 bit
// for bit (tests/golden/orb_extract.npz).  Everything that touches pixels runs here:
//
//   orb_gray_kernel      BGR -> gray, 15-bit fixed point                                                      1 launch
//   orb_resize_kernel    level l from level l - 1, 8.8 x 8.8 fixed-point bilinear, one rounding              7 launches (a chain)
//   orb_fast_kernel      FAST-9/16 corner test + corner score for every pixel of every level                 1 launch, all levels
//   orb_nms_kernel       strict 8-neighbour maxima inside the 31-pixel edge band: one bit per pixel + per-row counts
//   orb_scan_kernel      exclusive scan of the row counts (raster order is the order cv::FAST emits corners in)
//   orb_pos_kernel       one warp per row: the row's corners, left to right, into their slots
//   orb_cand_kernel      one warp per corner: Harris response (7 x 7 block, integer gradients), intensity-centroid moments over the
//                        radius-15 disc, fastAtan2
//   orb_blur_kernel      7 x 7 sigma-2 blur of every level that has key points: float32 separable, fused multiply-adds in the order
//                        OpenCV's AVX2 build executes them, round-half-even
//   orb_desc_kernel      one warp per key point, one descriptor byte per lane: 16 rotated samples, 8 comparisons
//
// The host between orb_cand_kernel and orb_desc_kernel does what cannot be restated as a parallel selection without changing the result:
// KeyPointsFilter::retainBest is std::nth_element + std::partition, and the order those leave the survivors in is the row order of the
// frame's descriptors.  It is O(corners) per level on device-computed responses; cos / sin of the angle are evaluated there too (the same
// libm call OpenCV makes).  All float expressions use round-to-nearest intrinsics so that nvcc cannot contract them differently from the
// x86 code they mirror.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/esfm_match.h"
#include "esfm_internal.cuh"
#include "host_internal.h"
#include "orb_host.h"

namespace esfm {

namespace {

constexpr int kOrbLevels = 8;
constexpr int kOrbEdge = 31;
constexpr int kOrbHalfPatch = 15;
constexpr int kOrbFastThr = 20;
constexpr int kFastTileW = 32, kFastTileH = 16;
constexpr int kBlurTile = 32;

struct OrbLevels {
    int n;
    int w[kOrbLevels], h[kOrbLevels], pitch[kOrbLevels];
    long long off[kOrbLevels];        // byte offset of the level inside the pyramid buffer
    int row0[kOrbLevels + 1];         // first global row index (rows of all levels back to back)
    int tile0[kOrbLevels + 1];        // first tile index of the batched tile kernels
    int words[kOrbLevels];            // 32-pixel words per row of the corner bit map
    long long word0[kOrbLevels + 1];  // first word of the level in the bit map
    unsigned blur_mask;               // levels orb_blur_kernel works on
};

struct OrbCand {
    int xy;            // x | y << 14 | level << 28, level coordinates
    float score;       // FAST corner score
    float harris;
    float angle;
};

struct OrbKp {
    int cx, cy, level;
    float a, b;        // cos / sin of the angle
};

__constant__ signed char c_orb_pattern[1024] = {
#include "orb_pattern.inc"
};

// half-widths of the rows of the radius-15 disc (orb.cpp umax)
__constant__ int c_umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};

__global__ void orb_gray_kernel(const uint8_t* __restrict__ src, size_t stride, int channels, int w, int h, uint8_t* __restrict__ dst, int pitch) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const uint8_t* p = src + (size_t)y * stride + (size_t)x * channels;
    int v = p[0];
    if (channels == 3) v = (p[0] * 3735 + p[1] * 19235 + p[2] * 9798 + (1 << 14)) >> 15;
    dst[(size_t)y * pitch + x] = (uint8_t)v;
}

// resize.cpp, INTER_LINEAR_EXACT for 8-bit: source index and an 8.8 weight per axis from (d + 0.5) * (src / dst) - 0.5 in double
__device__ __forceinline__ void linear_coeff(int d, int src, double scale, int& i0, int& i1, int& c) {
    const double fx = ((double)d + 0.5) * scale - 0.5;
    const double fl = floor(fx);
    int sx = (int)fl;
    double f = fx - fl;
    if (sx < 0) { sx = 0; f = 0.0; }
    if (sx >= src - 1) { sx = src - 1; f = 0.0; }
    i0 = sx;
    i1 = min(sx + 1, src - 1);
    c = __double2int_rn(f * 256.0);
}

__global__ void orb_resize_kernel(const uint8_t* __restrict__ src, int sw, int sh, int spitch, uint8_t* __restrict__ dst, int dw, int dh, int dpitch) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dw || y >= dh) return;
    int x0, x1, cx, y0, y1, cy;
    linear_coeff(x, sw, (double)sw / (double)dw, x0, x1, cx);
    linear_coeff(y, sh, (double)sh / (double)dh, y0, y1, cy);
    const uint8_t* r0 = src + (size_t)y0 * spitch;
    const uint8_t* r1 = src + (size_t)y1 * spitch;
    const int h0 = r0[x0] * (256 - cx) + r0[x1] * cx;
    const int h1 = r1[x0] * (256 - cx) + r1[x1] * cx;
    const int v = h0 * (256 - cy) + h1 * cy;
    dst[(size_t)y * dpitch + x] = (uint8_t)((v + (1 << 15)) >> 16);
}

__device__ __forceinline__ int level_of_tile(const OrbLevels& L, int tile) {
    int l = 0;
    while (l + 1 < L.n && tile >= L.tile0[l + 1]) ++l;
    return l;
}

// FAST-9/16 (fast.cpp FAST_t<16>) with the corner score of fast_score.cpp: 0 for a non-corner, else (max over the sixteen 9-pixel arcs of
// the arc's weakest contrast against the centre, either polarity) - 1, which is the largest threshold the pixel still passes.
__global__ void __launch_bounds__(kFastTileW * kFastTileH) orb_fast_kernel(const uint8_t* __restrict__ pyr, OrbLevels L, uint8_t* __restrict__ score) {
    __shared__ uint8_t tile[kFastTileH + 6][kFastTileW + 8];
    const int l = level_of_tile(L, blockIdx.x);
    const int w = L.w[l], h = L.h[l], pitch = L.pitch[l];
    const uint8_t* img = pyr + L.off[l];
    const int tiles_x = (w + kFastTileW - 1) / kFastTileW;
    const int t = blockIdx.x - L.tile0[l];
    const int bx = (t % tiles_x) * kFastTileW, by = (t / tiles_x) * kFastTileH;
    const int tid = threadIdx.y * kFastTileW + threadIdx.x;
    for (int i = tid; i < (kFastTileH + 6) * (kFastTileW + 6); i += kFastTileW * kFastTileH) {
        const int ty = i / (kFastTileW + 6), tx = i % (kFastTileW + 6);
        const int gx = min(max(bx + tx - 3, 0), w - 1), gy = min(max(by + ty - 3, 0), h - 1);
        tile[ty][tx] = img[(size_t)gy * pitch + gx];
    }
    __syncthreads();
    const int x = bx + threadIdx.x, y = by + threadIdx.y;
    if (x >= w || y >= h) return;
    int out = 0;
    if (x >= 3 && x < w - 3 && y >= 3 && y < h - 3) {
        const int cx = threadIdx.x + 3, cy = threadIdx.y + 3;
        const int v = tile[cy][cx];
        int d[16];
        d[0] = v - tile[cy + 3][cx];      d[1] = v - tile[cy + 3][cx + 1];  d[2] = v - tile[cy + 2][cx + 2];  d[3] = v - tile[cy + 1][cx + 3];
        d[4] = v - tile[cy][cx + 3];      d[5] = v - tile[cy - 1][cx + 3];  d[6] = v - tile[cy - 2][cx + 2];  d[7] = v - tile[cy - 3][cx + 1];
        d[8] = v - tile[cy - 3][cx];      d[9] = v - tile[cy - 3][cx - 1];  d[10] = v - tile[cy - 2][cx - 2]; d[11] = v - tile[cy - 1][cx - 3];
        d[12] = v - tile[cy][cx - 3];     d[13] = v - tile[cy + 1][cx - 3]; d[14] = v - tile[cy + 2][cx - 2]; d[15] = v - tile[cy + 3][cx - 1];
        unsigned bright = 0, dark = 0;     // centre brighter / darker than the circle pixel by more than the threshold
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            bright |= (unsigned)(d[k] > kOrbFastThr) << k;
            dark |= (unsigned)(d[k] < -kOrbFastThr) << k;
        }
        auto has_arc9 = [](unsigned m) {
            unsigned m2 = m | (m << 16);
            unsigned r = m2 & (m2 >> 1);
            r &= r >> 2;
            r &= r >> 4;
            r &= m2 >> 8;
            return r != 0;
        };
        if (has_arc9(bright) || has_arc9(dark)) {
            int best = 0;
#pragma unroll
            for (int pol = 0; pol < 2; ++pol) {
                int m2[16], m4[16], m8[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) m2[k] = min(d[k], d[(k + 1) & 15]);
#pragma unroll
                for (int k = 0; k < 16; ++k) m4[k] = min(m2[k], m2[(k + 2) & 15]);
#pragma unroll
                for (int k = 0; k < 16; ++k) m8[k] = min(m4[k], m4[(k + 4) & 15]);
#pragma unroll
                for (int k = 0; k < 16; ++k) best = max(best, min(m8[k], d[(k + 8) & 15]));
#pragma unroll
                for (int k = 0; k < 16; ++k) d[k] = -d[k];
            }
            out = best - 1;         // best > threshold here, so 20 <= out <= 254
        }
    }
    score[L.off[l] + (size_t)y * pitch + x] = (uint8_t)out;
}

// Non-maximum suppression (fast.cpp: strictly greater than all eight neighbours' scores) and KeyPointsFilter::runByImageBorder(31): one
// warp per image row, one bit per pixel, and the row's corner count.
__global__ void orb_nms_kernel(const uint8_t* __restrict__ score, OrbLevels L, unsigned* __restrict__ bits, int* __restrict__ row_cnt) {
    const int grow = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (grow >= L.row0[L.n]) return;
    int l = 0;
    while (l + 1 < L.n && grow >= L.row0[l + 1]) ++l;
    const int y = grow - L.row0[l], w = L.w[l], h = L.h[l], pitch = L.pitch[l];
    unsigned* wrow = bits + L.word0[l] + (size_t)y * L.words[l];
    int cnt = 0;
    const bool row_in = y >= kOrbEdge && y < h - kOrbEdge && w > 2 * kOrbEdge && h > 2 * kOrbEdge;
    const uint8_t* s = score + L.off[l] + (size_t)y * pitch;
    for (int wi = 0; wi < L.words[l]; ++wi) {
        const int x = wi * 32 + lane;
        bool keep = false;
        if (row_in && x >= kOrbEdge && x < w - kOrbEdge) {
            const int c = s[x];
            keep = c > 0 && c > s[x - 1] && c > s[x + 1] && c > s[x - pitch - 1] && c > s[x - pitch] && c > s[x - pitch + 1] &&
                   c > s[x + pitch - 1] && c > s[x + pitch] && c > s[x + pitch + 1];
        }
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) wrow[wi] = m;
        cnt += __popc(m);
    }
    if (lane == 0) row_cnt[grow] = cnt;
}

// Exclusive scan of the row counts (one block), plus where every level starts.
__global__ void __launch_bounds__(1024) orb_scan_kernel(const int* __restrict__ row_cnt, int n_rows, int* __restrict__ row_off, OrbLevels L, int* __restrict__ level_off) {
    __shared__ int part[1024];
    const int per = (n_rows + 1023) / 1024;
    const int lo = min(threadIdx.x * per, n_rows), hi = min(lo + per, n_rows);
    int s = 0;
    for (int i = lo; i < hi; ++i) s += row_cnt[i];
    part[threadIdx.x] = s;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
        const int v = threadIdx.x >= d ? part[threadIdx.x - d] : 0;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    int run = part[threadIdx.x] - s;
    for (int i = lo; i < hi; ++i) { row_off[i] = run; run += row_cnt[i]; }
    if (threadIdx.x == 1023) row_off[n_rows] = part[1023];
    __syncthreads();
    if (threadIdx.x <= L.n) level_off[threadIdx.x] = row_off[L.row0[threadIdx.x]];
}

// core fastAtan2 (degrees), float32, no contraction
__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
    const float scale = (float)(180.0 / 3.141592653589793238462643383279502884);
    const float p1 = __fmul_rn(0.9997878412794807f, scale), p3 = __fmul_rn(-0.3258083974640975f, scale);
    const float p5 = __fmul_rn(0.1555786518463281f, scale), p7 = __fmul_rn(-0.04432655554792128f, scale);
    const float eps = 2.2204460492503131e-16f;
    const float ax = fabsf(x), ay = fabsf(y);
    float a;
    if (ax >= ay) {
        const float c = __fdiv_rn(ay, __fadd_rn(ax, eps));
        const float c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        const float c = __fdiv_rn(ax, __fadd_rn(ay, eps));
        const float c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0.f) a = __fsub_rn(180.f, a);
    if (y < 0.f) a = __fsub_rn(360.f, a);
    return a;
}

// One warp per image row: the corners of the row, left to right, get the slots row_off[row] .. and their packed position
// (x | y << 14 | level << 28) -- raster order is the order cv::FAST emits corners in.
__global__ void orb_pos_kernel(OrbLevels L, const unsigned* __restrict__ bits, const int* __restrict__ row_off, OrbCand* __restrict__ cand) {
    const int grow = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (grow >= L.row0[L.n]) return;
    int at = row_off[grow];
    if (row_off[grow + 1] == at) return;
    int l = 0;
    while (l + 1 < L.n && grow >= L.row0[l + 1]) ++l;
    const int y0 = grow - L.row0[l];
    const unsigned* wrow = bits + L.word0[l] + (size_t)y0 * L.words[l];
    for (int base = 0; base < L.words[l]; base += 32) {
        const int wi = base + lane;
        unsigned m = wi < L.words[l] ? wrow[wi] : 0u;
        const int cnt = __popc(m);
        int incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += v;
        }
        int slot = at + incl - cnt;
        while (m) {
            const int x0 = wi * 32 + __ffs(m) - 1;
            m &= m - 1;
            cand[slot++].xy = x0 | (y0 << 14) | (l << 28);
        }
        at += __shfl_sync(0xffffffffu, incl, 31);
    }
}

// One warp per corner: FAST score, Harris response (orb.cpp HarrisResponses: block 7, k 0.04) and orientation (orb.cpp ICAngles).
__global__ void __launch_bounds__(256) orb_cand_kernel(const uint8_t* __restrict__ pyr, const uint8_t* __restrict__ score, OrbLevels L, int n_cand,
                                                      float harris_k, float scale4, OrbCand* __restrict__ cand) {
    const int idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (idx >= n_cand) return;
    const int xy = cand[idx].xy;
    const int x0 = xy & 0x3fff, y0 = (xy >> 14) & 0x3fff, l = xy >> 28;
    const int pitch = L.pitch[l];
    const uint8_t* c = pyr + L.off[l] + (size_t)y0 * pitch + x0;
    // Harris: 49 block pixels over the lanes
    int a = 0, b = 0, cc = 0;
    for (int t = lane; t < 49; t += 32) {
        const uint8_t* p = c + (t / 7 - 3) * pitch + (t % 7 - 3);
        const int ix = ((int)p[1] - (int)p[-1]) * 2 + ((int)p[-pitch + 1] - (int)p[-pitch - 1]) + ((int)p[pitch + 1] - (int)p[pitch - 1]);
        const int iy = ((int)p[pitch] - (int)p[-pitch]) * 2 + ((int)p[pitch - 1] - (int)p[-pitch - 1]) + ((int)p[pitch + 1] - (int)p[-pitch + 1]);
        a += ix * ix; b += iy * iy; cc += ix * iy;
    }
    // intensity centroid: lane <-> column u = lane - 15 of the disc
    int m10 = 0, m01 = 0;
    if (lane < 31) {
        const int u = lane - kOrbHalfPatch, au = abs(u);
        m10 = u * (int)c[u];
        for (int v = 1; v <= kOrbHalfPatch; ++v) {
            if (au <= c_umax[v]) {
                const int plus = c[v * pitch + u], minus = c[-v * pitch + u];
                m01 += v * (plus - minus);
                m10 += u * (plus + minus);
            }
        }
    }
    a = __reduce_add_sync(0xffffffffu, a);
    b = __reduce_add_sync(0xffffffffu, b);
    cc = __reduce_add_sync(0xffffffffu, cc);
    m10 = __reduce_add_sync(0xffffffffu, m10);
    m01 = __reduce_add_sync(0xffffffffu, m01);
    if (lane == 0) {
        const float fa = (float)a, fb = (float)b, fc = (float)cc;
        const float tr = __fadd_rn(fa, fb);
        OrbCand o;
        o.xy = xy;
        o.score = (float)score[L.off[l] + (size_t)y0 * pitch + x0];
        o.harris = __fmul_rn(__fsub_rn(__fsub_rn(__fmul_rn(fa, fb), __fmul_rn(fc, fc)), __fmul_rn(__fmul_rn(harris_k, tr), tr)), scale4);
        o.angle = fast_atan2_deg((float)m01, (float)m10);
        cand[idx] = o;
    }
}

__device__ __forceinline__ int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
    return i;
}

// GaussianBlur(7 x 7, sigma 2, BORDER_REFLECT_101) as ORB's sub-matrix call executes it (filter.simd.hpp, float32 kernel):
// row pass s = k0 p0; s = fma(k_i, p_i, s); column pass s = k3 r0; s = fma(k_{3+j}, r_{+j} + r_{-j}, s); round half to even, saturate.
__global__ void __launch_bounds__(kBlurTile * 8) orb_blur_kernel(const uint8_t* __restrict__ pyr, OrbLevels L, const float4 k03, const float4 k46,
                                                               uint8_t* __restrict__ blurred) {
    __shared__ uint8_t raw[kBlurTile + 6][kBlurTile + 8];
    __shared__ float rows[kBlurTile + 6][kBlurTile + 1];
    const int l = level_of_tile(L, blockIdx.x);
    if (!((L.blur_mask >> l) & 1)) return;
    const int w = L.w[l], h = L.h[l], pitch = L.pitch[l];
    const uint8_t* img = pyr + L.off[l];
    const int tiles_x = (w + kBlurTile - 1) / kBlurTile;
    const int t = blockIdx.x - L.tile0[l];
    const int bx = (t % tiles_x) * kBlurTile, by = (t / tiles_x) * kBlurTile;
    const int tid = threadIdx.y * kBlurTile + threadIdx.x;
    for (int i = tid; i < (kBlurTile + 6) * (kBlurTile + 6); i += kBlurTile * 8) {
        const int ty = i / (kBlurTile + 6), tx = i % (kBlurTile + 6);
        raw[ty][tx] = img[(size_t)reflect101(by + ty - 3, h) * pitch + reflect101(bx + tx - 3, w)];
    }
    __syncthreads();
    const float k[7] = {k03.x, k03.y, k03.z, k03.w, k46.x, k46.y, k46.z};
    for (int ty = threadIdx.y; ty < kBlurTile + 6; ty += 8) {
        const uint8_t* p = &raw[ty][threadIdx.x];
        float s = __fmul_rn(k[0], (float)p[0]);
#pragma unroll
        for (int i = 1; i < 7; ++i) s = __fmaf_rn(k[i], (float)p[i], s);
        rows[ty][threadIdx.x] = s;
    }
    __syncthreads();
    const int x = bx + threadIdx.x;
    for (int ty = threadIdx.y; ty < kBlurTile; ty += 8) {
        const int y = by + ty;
        if (x >= w || y >= h) continue;
        float s = __fmul_rn(k[3], rows[ty + 3][threadIdx.x]);
#pragma unroll
        for (int j = 1; j <= 3; ++j) s = __fmaf_rn(k[3 + j], __fadd_rn(rows[ty + 3 + j][threadIdx.x], rows[ty + 3 - j][threadIdx.x]), s);
        const int v = __float2int_rn(s);
        blurred[L.off[l] + (size_t)y * pitch + x] = (uint8_t)min(max(v, 0), 255);
    }
}

// orb.cpp computeOrbDescriptors, WTA_K 2: one warp per key point, lane = descriptor byte.
__global__ void __launch_bounds__(256) orb_desc_kernel(const uint8_t* __restrict__ blurred, OrbLevels L, const OrbKp* __restrict__ kps, int n,
                                                      uint8_t* __restrict__ desc) {
    __shared__ signed char pat[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) pat[i] = c_orb_pattern[i];
    __syncthreads();
    const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (k >= n) return;
    const OrbKp kp = kps[k];
    const int pitch = L.pitch[kp.level];
    const uint8_t* c = blurred + L.off[kp.level] + (size_t)kp.cy * pitch + kp.cx;
    const float a = kp.a, b = kp.b;
    unsigned byte = 0;
#pragma unroll
    for (int bit = 0; bit < 8; ++bit) {
        const signed char* p = &pat[(lane * 8 + bit) * 4];
        int t[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const float px = (float)p[2 * e], py = (float)p[2 * e + 1];
            const float x = __fsub_rn(__fmul_rn(px, a), __fmul_rn(py, b));
            const float y = __fadd_rn(__fmul_rn(px, b), __fmul_rn(py, a));
            t[e] = c[__float2int_rn(y) * pitch + __float2int_rn(x)];
        }
        byte |= (unsigned)(t[0] < t[1]) << bit;
    }
    desc[(size_t)k * 32 + lane] = (uint8_t)byte;
}

using orbhost::RespItem;
using orbhost::retain_best;

template <typename T>
int grow_dev(esfm_ctx* ctx, T** p, size_t* cap, size_t need) {
    if (*cap >= need && *p) return ESFM_OK;
    if (*p) cudaFreeAsync(*p, ctx->stream);
    *p = nullptr; *cap = 0;
    const size_t want = need + need / 4 + 256;
    CUDA_TRY(cudaMallocFromPoolAsync((void**)p, want * sizeof(T), ctx->mempool, ctx->stream));
    *cap = want;
    return ESFM_OK;
}

template <typename T>
int grow_host(T** p, size_t* cap, size_t need) {
    if (*cap >= need && *p) return ESFM_OK;
    if (*p) cudaFreeHost(*p);
    *p = nullptr; *cap = 0;
    const size_t want = need + need / 4 + 256;
    CUDA_TRY(cudaMallocHost((void**)p, want * sizeof(T)));
    *cap = want;
    return ESFM_OK;
}

// A few persistent host threads for the per-level selection (levels are independent; level 0 holds ~31 % of the corners).
class TaskPool {
public:
    explicit TaskPool(int helpers) {
        for (int k = 0; k < helpers; ++k) threads_.emplace_back([this] { worker(); });
    }
    ~TaskPool() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
            ++epoch_;
        }
        cv_work_.notify_all();
        for (auto& t : threads_) t.join();
    }
    // fn(i) for i in [0, count), on the helpers and the calling thread; returns when all are done
    void run(int count, const std::function<void(int)>& fn) {
        if (threads_.empty() || count <= 1) {
            for (int i = 0; i < count; ++i) fn(i);
            return;
        }
        {
            std::lock_guard<std::mutex> lk(mu_);
            fn_ = &fn;
            count_ = count;
            next_.store(0);
            active_ = (int)threads_.size();
            ++epoch_;
        }
        cv_work_.notify_all();
        for (int i; (i = next_.fetch_add(1)) < count;) fn(i);
        std::unique_lock<std::mutex> lk(mu_);
        cv_done_.wait(lk, [&] { return active_ == 0; });
        fn_ = nullptr;
    }
private:
    void worker() {
        uint64_t seen = 0;
        for (;;) {
            const std::function<void(int)>* fn;
            int count;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_work_.wait(lk, [&] { return epoch_ != seen; });
                seen = epoch_;
                if (stop_) return;
                fn = fn_;
                count = count_;
            }
            for (int i; (i = next_.fetch_add(1)) < count;) (*fn)(i);
            {
                std::lock_guard<std::mutex> lk(mu_);
                --active_;
            }
            cv_done_.notify_one();
        }
    }
    std::vector<std::thread> threads_;
    std::mutex mu_;
    std::condition_variable cv_work_, cv_done_;
    const std::function<void(int)>* fn_ = nullptr;
    std::atomic<int> next_{0};
    int count_ = 0, active_ = 0;
    uint64_t epoch_ = 0;
    bool stop_ = false;
};

}  // namespace

// Per-context scratch of the extractor: kept between calls, so a run over many frames of one size allocates once.
struct OrbState {
    uint8_t* d_src = nullptr;    size_t src_cap = 0;
    uint8_t* d_pyr = nullptr;    size_t pyr_cap = 0;
    uint8_t* d_score = nullptr;  size_t score_cap = 0;
    uint8_t* d_blur = nullptr;   size_t blur_cap = 0;
    unsigned* d_bits = nullptr;  size_t bits_cap = 0;
    int* d_rows = nullptr;       size_t rows_cap = 0;       // row_cnt | row_off | level_off
    OrbCand* d_cand = nullptr;   size_t cand_cap = 0;
    OrbKp* d_kp = nullptr;       size_t kp_cap = 0;
    uint8_t* d_desc = nullptr;   size_t desc_cap = 0;
    uint8_t* h_src = nullptr;    size_t h_src_cap = 0;      // pinned
    OrbCand* h_cand = nullptr;   size_t h_cand_cap = 0;
    OrbKp* h_kp = nullptr;       size_t h_kp_cap = 0;
    int* h_level_off = nullptr;  size_t h_level_cap = 0;
    OrbLevels levels{};
    bool have_levels = false;
    TaskPool* pool = nullptr;                               // host threads of the selection step
    std::vector<esfm_keypoint_t> sel_kp[kOrbLevels];
    std::vector<OrbKp> sel_dk[kOrbLevels];
    double phase_ms[5] = {0, 0, 0, 0, 0};                   // last call: front end | corner records | host selection | descriptors | total
    int last_corners = 0;
};

void orb_state_destroy(esfm_ctx* ctx) {
    OrbState* s = ctx->orb;
    if (!s) return;
    void* dev[] = {s->d_src, s->d_pyr, s->d_score, s->d_blur, s->d_bits, s->d_rows, s->d_cand, s->d_kp, s->d_desc};
    for (void* p : dev) if (p) cudaFreeAsync(p, ctx->stream);
    void* host[] = {s->h_src, s->h_cand, s->h_kp, s->h_level_off};
    for (void* p : host) if (p) cudaFreeHost(p);
    delete s->pool;
    delete s;
    ctx->orb = nullptr;
}

int orb_extract_impl(esfm_ctx* ctx, const unsigned char* image, int rows, int cols, int channels, size_t row_stride, int max_features,
                     esfm_keypoint_t* keypoints, unsigned char* h_desc, int capacity, int* n_out, const OrbSink& sink) {
    if (!ctx || !image || !n_out) return fail(ESFM_ERR_INVALID, "esfm_orb_extract: NULL argument");
    if (rows < 1 || cols < 1 || rows > 16383 || cols > 16383) return fail(ESFM_ERR_INVALID, "image size %d x %d out of range", cols, rows);
    if (channels != 1 && channels != 3) return fail(ESFM_ERR_INVALID, "image must have 1 (gray) or 3 (BGR) 8-bit channels, got %d", channels);
    if (row_stride < (size_t)cols * channels) return fail(ESFM_ERR_INVALID, "row_stride %zu smaller than a row", row_stride);
    if (max_features < 0 || capacity < 0 || (capacity > 0 && !keypoints)) return fail(ESFM_ERR_INVALID, "bad max_features / capacity / keypoints");
    if (int rc = set_device(ctx)) return rc;
    if (!ctx->orb) {
        ctx->orb = new OrbState();
        const char* e = getenv("ESFM_ORB_THREADS");
        const int want = e ? atoi(e) : std::min(kOrbLevels, (int)std::thread::hardware_concurrency());
        ctx->orb->pool = new TaskPool(std::max(want, 1) - 1);
    }
    OrbState& S = *ctx->orb;
    cudaStream_t st = ctx->stream;
    *n_out = 0;
    using clk = std::chrono::steady_clock;
    auto ms_since = [](clk::time_point t) { return std::chrono::duration<double, std::milli>(clk::now() - t).count(); };
    const clk::time_point t_start = clk::now();
    clk::time_point t_phase = t_start;

    // ---- level geometry (orb.cpp detectAndCompute): scale_l = (float)pow(1.2f as double, l); size = cvRound(cols * (1.f / scale_l)) ----
    static_assert(kOrbLevels == orbhost::kLevels, "level count");
    float scale[kOrbLevels];
    OrbLevels L{};
    L.n = kOrbLevels;
    long long off = 0, word = 0;
    int row = 0, tile_fast = 0;
    for (int l = 0; l < kOrbLevels; ++l) {
        scale[l] = orbhost::level_scale(l);
        L.w[l] = orbhost::level_extent(cols, l);
        L.h[l] = orbhost::level_extent(rows, l);
        if (L.w[l] < 1 || L.h[l] < 1) return fail(ESFM_ERR_INVALID, "image %d x %d is too small for 8 pyramid levels", cols, rows);
        L.pitch[l] = (L.w[l] + 15) & ~15;
        L.off[l] = off;
        off += (long long)L.pitch[l] * L.h[l];
        off = (off + 255) & ~255LL;
        L.row0[l] = row; row += L.h[l];
        L.words[l] = (L.w[l] + 31) / 32;
        L.word0[l] = word; word += (long long)L.words[l] * L.h[l];
        L.tile0[l] = tile_fast;
        tile_fast += ((L.w[l] + kFastTileW - 1) / kFastTileW) * ((L.h[l] + kFastTileH - 1) / kFastTileH);
    }
    L.row0[kOrbLevels] = row;
    L.word0[kOrbLevels] = word;
    L.tile0[kOrbLevels] = tile_fast;
    const size_t pyr_bytes = (size_t)off;
    const int n_rows = row;

    // features per level (orb.cpp computeKeyPoints), float arithmetic as written there
    int per_level[kOrbLevels];
    orbhost::features_per_level(max_features, per_level);

    // ---- upload, gray, pyramid, FAST, NMS, scan ----
    const size_t src_row = (size_t)cols * channels, src_bytes = src_row * rows;
    if (int rc = grow_dev(ctx, &S.d_src, &S.src_cap, src_bytes)) return rc;
    if (int rc = grow_host(&S.h_src, &S.h_src_cap, src_bytes)) return rc;
    if (int rc = grow_dev(ctx, &S.d_pyr, &S.pyr_cap, pyr_bytes)) return rc;
    if (int rc = grow_dev(ctx, &S.d_score, &S.score_cap, pyr_bytes)) return rc;
    if (int rc = grow_dev(ctx, &S.d_blur, &S.blur_cap, pyr_bytes)) return rc;
    if (int rc = grow_dev(ctx, &S.d_bits, &S.bits_cap, (size_t)word)) return rc;
    if (int rc = grow_dev(ctx, &S.d_rows, &S.rows_cap, (size_t)2 * n_rows + 2 + kOrbLevels + 1)) return rc;
    if (int rc = grow_host(&S.h_level_off, &S.h_level_cap, (size_t)kOrbLevels + 1)) return rc;
    int* d_row_cnt = S.d_rows;
    int* d_row_off = S.d_rows + n_rows;
    int* d_level_off = S.d_rows + 2 * n_rows + 1;
    {
        // pageable image -> pinned staging -> device in a few row bands, so the host copy of band k + 1 overlaps the upload of band k
        const int bands = src_bytes > ((size_t)1 << 20) ? 4 : 1;
        for (int k = 0; k < bands; ++k) {
            const int r0 = (int)((long long)rows * k / bands), r1 = (int)((long long)rows * (k + 1) / bands);
            uint8_t* dst = S.h_src + (size_t)r0 * src_row;
            if (row_stride == src_row) ctx->copier->copy(dst, image + (size_t)r0 * row_stride, (size_t)(r1 - r0) * src_row);
            else for (int r = r0; r < r1; ++r) memcpy(S.h_src + (size_t)r * src_row, image + (size_t)r * row_stride, src_row);
            CUDA_TRY(cudaMemcpyAsync(S.d_src + (size_t)r0 * src_row, dst, (size_t)(r1 - r0) * src_row, cudaMemcpyHostToDevice, st));
        }
        ctx->stats.h2d_bytes += src_bytes;
    }
    {
        dim3 blk(32, 8), grd((cols + 31) / 32, (rows + 7) / 8);
        orb_gray_kernel<<<grd, blk, 0, st>>>(S.d_src, src_row, channels, cols, rows, S.d_pyr + L.off[0], L.pitch[0]);
        for (int l = 1; l < kOrbLevels; ++l) {
            dim3 g2((L.w[l] + 31) / 32, (L.h[l] + 7) / 8);
            orb_resize_kernel<<<g2, blk, 0, st>>>(S.d_pyr + L.off[l - 1], L.w[l - 1], L.h[l - 1], L.pitch[l - 1], S.d_pyr + L.off[l], L.w[l],
                                                  L.h[l], L.pitch[l]);
        }
        orb_fast_kernel<<<tile_fast, dim3(kFastTileW, kFastTileH), 0, st>>>(S.d_pyr, L, S.d_score);
        const int warps_per_block = 8;
        orb_nms_kernel<<<(n_rows + warps_per_block - 1) / warps_per_block, warps_per_block * 32, 0, st>>>(S.d_score, L, S.d_bits, d_row_cnt);
        orb_scan_kernel<<<1, 1024, 0, st>>>(d_row_cnt, n_rows, d_row_off, L, d_level_off);
    }
    CUDA_TRY(cudaMemcpyAsync(S.h_level_off, d_level_off, (kOrbLevels + 1) * sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaGetLastError());
    const int n_cand = S.h_level_off[kOrbLevels];
    S.phase_ms[0] = ms_since(t_phase); t_phase = clk::now();
    S.last_corners = n_cand;

    // ---- per-corner responses and angles ----
    if (int rc = grow_dev(ctx, &S.d_cand, &S.cand_cap, (size_t)std::max(n_cand, 1))) return rc;
    if (int rc = grow_host(&S.h_cand, &S.h_cand_cap, (size_t)std::max(n_cand, 1))) return rc;
    if (n_cand > 0) {
        const float scale4 = orbhost::harris_scale4();
        orb_pos_kernel<<<(n_rows + 7) / 8, 256, 0, st>>>(L, S.d_bits, d_row_off, S.d_cand);
        orb_cand_kernel<<<(n_cand + 7) / 8, 256, 0, st>>>(S.d_pyr, S.d_score, L, n_cand, 0.04f, scale4, S.d_cand);
        CUDA_TRY(cudaMemcpyAsync(S.h_cand, S.d_cand, (size_t)n_cand * sizeof(OrbCand), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        CUDA_TRY(cudaGetLastError());
        ctx->stats.d2h_bytes += (size_t)n_cand * sizeof(OrbCand);
    }

    S.phase_ms[1] = ms_since(t_phase); t_phase = clk::now();

    // ---- selection (orb.cpp computeKeyPoints: retainBest(2 n) on the FAST score, Harris, retainBest(n)), level by level ----
    S.pool->run(kOrbLevels, [&](int l) {
        std::vector<esfm_keypoint_t>& okp = S.sel_kp[l];
        std::vector<OrbKp>& odk = S.sel_dk[l];
        okp.clear();
        odk.clear();
        const int lo = S.h_level_off[l], hi = S.h_level_off[l + 1];
        if (hi == lo) return;
        std::vector<RespItem> a((size_t)(hi - lo)), b;
        for (int i = lo; i < hi; ++i) a[i - lo] = RespItem{S.h_cand[i].score, i};
        retain_best(a, 2 * per_level[l]);
        b.resize(a.size());
        for (size_t i = 0; i < a.size(); ++i) b[i] = RespItem{S.h_cand[a[i].index].harris, a[i].index};
        retain_best(b, per_level[l]);
        const float sc = scale[l];
        okp.reserve(b.size());
        odk.reserve(b.size());
        for (const RespItem& it : b) {
            const OrbCand& c = S.h_cand[it.index];
            const orbhost::KeyPointOut ko = orbhost::keypoint_of(c.xy & 0x3fff, (c.xy >> 14) & 0x3fff, sc);
            esfm_keypoint_t k;
            k.x = ko.x;
            k.y = ko.y;
            k.size = ko.size;
            k.angle = c.angle;
            k.response = c.harris;
            k.octave = l;
            okp.push_back(k);
            // orb.cpp computeOrbDescriptors re-derives the level position and the rotation from the key point it is handed
            const orbhost::SampleFrame sf = orbhost::sample_frame_of(ko, c.angle, sc);
            OrbKp d;
            d.level = l;
            d.cx = sf.cx;
            d.cy = sf.cy;
            d.a = sf.a;
            d.b = sf.b;
            odk.push_back(d);
        }
    });
    std::vector<esfm_keypoint_t> out;
    std::vector<OrbKp> kps;
    unsigned blur_mask = 0;
    for (int l = 0; l < kOrbLevels; ++l) {
        if (!S.sel_kp[l].empty()) blur_mask |= 1u << l;
        out.insert(out.end(), S.sel_kp[l].begin(), S.sel_kp[l].end());
        kps.insert(kps.end(), S.sel_dk[l].begin(), S.sel_dk[l].end());
    }
    const int n = (int)out.size();
    *n_out = n;
    S.phase_ms[2] = ms_since(t_phase); t_phase = clk::now();
    S.phase_ms[3] = 0.0;
    S.phase_ms[4] = ms_since(t_start);
    S.levels = L;
    S.levels.blur_mask = blur_mask;
    S.have_levels = true;
    if (n > capacity) return fail(ESFM_ERR_CAPACITY, "esfm_orb_extract: %d key points, capacity %d", n, capacity);
    if (n) memcpy(keypoints, out.data(), (size_t)n * sizeof(esfm_keypoint_t));
    uint8_t* d_dst = nullptr;
    if (sink) {
        if (int rc = sink(n, &d_dst)) return rc;
    }
    if (n == 0) return ESFM_OK;
    if (!d_dst) {
        if (int rc = grow_dev(ctx, &S.d_desc, &S.desc_cap, (size_t)n * 32)) return rc;
        d_dst = S.d_desc;
    }

    // ---- blur the levels that have key points, sample the descriptors ----
    if (int rc = grow_dev(ctx, &S.d_kp, &S.kp_cap, (size_t)n)) return rc;
    if (int rc = grow_host(&S.h_kp, &S.h_kp_cap, (size_t)n)) return rc;
    memcpy(S.h_kp, kps.data(), (size_t)n * sizeof(OrbKp));
    CUDA_TRY(cudaMemcpyAsync(S.d_kp, S.h_kp, (size_t)n * sizeof(OrbKp), cudaMemcpyHostToDevice, st));
    ctx->stats.h2d_bytes += (size_t)n * sizeof(OrbKp);
    {
        float kf[7];
        orbhost::gaussian_kernel_7(kf);
        OrbLevels B = S.levels;
        int tiles = 0;
        for (int l = 0; l < kOrbLevels; ++l) {
            B.tile0[l] = tiles;
            tiles += ((B.w[l] + kBlurTile - 1) / kBlurTile) * ((B.h[l] + kBlurTile - 1) / kBlurTile);
        }
        B.tile0[kOrbLevels] = tiles;
        orb_blur_kernel<<<tiles, dim3(kBlurTile, 8), 0, st>>>(S.d_pyr, B, make_float4(kf[0], kf[1], kf[2], kf[3]), make_float4(kf[4], kf[5], kf[6], 0.f),
                                                             S.d_blur);
        orb_desc_kernel<<<(n + 7) / 8, 256, 0, st>>>(S.d_blur, S.levels, S.d_kp, n, d_dst);
    }
    if (h_desc) {
        CUDA_TRY(cudaMemcpyAsync(h_desc, d_dst, (size_t)n * 32, cudaMemcpyDeviceToHost, st));
        ctx->stats.d2h_bytes += (size_t)n * 32;
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaGetLastError());
    S.phase_ms[3] = ms_since(t_phase);
    S.phase_ms[4] = ms_since(t_start);
    return ESFM_OK;
}

}  // namespace esfm

using namespace esfm;

extern "C" int esfm_orb_extract(esfm_ctx_t* ctx, const unsigned char* image, int rows, int cols, int channels, size_t row_stride, int max_features,
                                esfm_keypoint_t* keypoints, unsigned char* descriptors, int capacity, int* n_out) {
    return orb_extract_impl(ctx, image, rows, cols, channels, row_stride, max_features, keypoints, descriptors, capacity, n_out, OrbSink());
}

extern "C" int esfm_orb_last_timing(esfm_ctx_t* ctx, double* phase_ms, int* corners) {
    if (!ctx || !phase_ms) return fail(ESFM_ERR_INVALID, "esfm_orb_last_timing: NULL argument");
    if (!ctx->orb || !ctx->orb->have_levels) return fail(ESFM_ERR_STATE, "esfm_orb_last_timing: no esfm_orb_extract call on this context yet");
    for (int i = 0; i < 5; ++i) phase_ms[i] = ctx->orb->phase_ms[i];
    if (corners) *corners = ctx->orb->last_corners;
    return ESFM_OK;
}

extern "C" int esfm_orb_debug_level(esfm_ctx_t* ctx, int level, int blurred, unsigned char* out, size_t out_bytes, int* rows, int* cols) {
    if (!ctx || !rows || !cols) return fail(ESFM_ERR_INVALID, "esfm_orb_debug_level: NULL argument");
    if (!ctx->orb || !ctx->orb->have_levels) return fail(ESFM_ERR_STATE, "esfm_orb_debug_level: no esfm_orb_extract call on this context yet");
    OrbState& S = *ctx->orb;
    if (level < 0 || level >= S.levels.n) return fail(ESFM_ERR_INVALID, "level %d out of range", level);
    *rows = S.levels.h[level];
    *cols = S.levels.w[level];
    if (!out) return ESFM_OK;
    if (blurred && !((S.levels.blur_mask >> level) & 1)) return fail(ESFM_ERR_STATE, "level %d had no key points and was not blurred", level);
    if (out_bytes < (size_t)*rows * *cols) return fail(ESFM_ERR_CAPACITY, "esfm_orb_debug_level: out holds %zu bytes, level needs %zu", out_bytes, (size_t)*rows * *cols);
    if (int rc = set_device(ctx)) return rc;
    const uint8_t* src = (blurred ? S.d_blur : S.d_pyr) + S.levels.off[level];
    CUDA_TRY(cudaMemcpy2DAsync(out, *cols, src, S.levels.pitch[level], *cols, *rows, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return ESFM_OK;
}

// bank.cu -- derived device layouts of the descriptor bank (run once per esfm_bank_commit).
//
// F32X64: rows_f32[total_rows][64] (row-major, as the reference's cv::Mat holds SURF descriptors,
// cpp_code/include/utility.h:31) -> k-major 128-row tiles + half squared norms, the operand layout the
// sweep kernel bulk-copies into shared memory with one cp.async.bulk per tile.
#include "esfm_internal.cuh"

namespace esfm {

__global__ void __launch_bounds__(256) pack_f32_kernel(const float* __restrict__ rows, const int* __restrict__ frame_rows,
                                                       const int* __restrict__ frame_row_off,
                                                       const int* __restrict__ frame_tile_off, int n_frames,
                                                       float* __restrict__ kmajor) {
    __shared__ float tile[kTile][kDim + 1];
    const int t = blockIdx.x;
    // frame owning tile t: last f with frame_tile_off[f] <= t
    int lo = 0, hi = n_frames - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (frame_tile_off[mid] <= t) lo = mid; else hi = mid - 1;
    }
    const int f = lo;
    const int row0 = (t - frame_tile_off[f]) * kTile;
    int valid = frame_rows[f] - row0;
    valid = valid < 0 ? 0 : (valid > kTile ? kTile : valid);
    const float4* src = reinterpret_cast<const float4*>(rows + ((size_t)frame_row_off[f] + row0) * kDim);
    for (int idx = threadIdx.x; idx < kTile * (kDim / 4); idx += blockDim.x) {
        const int r = idx / (kDim / 4), c4 = idx % (kDim / 4);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < valid) v = src[(size_t)r * (kDim / 4) + c4];
        tile[r][c4 * 4 + 0] = v.x;
        tile[r][c4 * 4 + 1] = v.y;
        tile[r][c4 * 4 + 2] = v.z;
        tile[r][c4 * 4 + 3] = v.w;
    }
    __syncthreads();
    float* out = kmajor + (size_t)t * kTileFloats;
    for (int idx = threadIdx.x; idx < kDim * kTile; idx += blockDim.x) {
        const int k = idx / kTile, r = idx % kTile;
        out[idx] = tile[r][k];
    }
    if (threadIdx.x < kTile) {
        const int r = threadIdx.x;
        float s = 0.f;
#pragma unroll 8
        for (int k = 0; k < kDim; ++k) s = __fmaf_rn(tile[r][k], tile[r][k], s);
        // pad rows: +inf half-norm makes every distance involving them +inf, so they never pass a threshold
        out[kDim * kTile + r] = (r < valid) ? 0.5f * s : __int_as_float(0x7f800000);
    }
}

cudaError_t launch_pack_f32(const float* rows, const int* frame_rows, const int* frame_row_off, const int* frame_tile_off,
                            int n_frames, int n_tiles_total, float* kmajor, cudaStream_t s) {
    if (n_tiles_total <= 0) return cudaSuccess;
    pack_f32_kernel<<<n_tiles_total, 256, 0, s>>>(rows, frame_rows, frame_row_off, frame_tile_off, n_frames, kmajor);
    return cudaGetLastError();
}

}  // namespace esfm

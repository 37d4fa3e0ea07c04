// orb_host.h -- the host arithmetic of the ORB extractor (orb.cu), kept free of CUDA so that tests/host/orb_host.cpp can compile it with g++
// and compare it with oracle/orb_oracle.py without a GPU.  Each function restates a few lines of OpenCV 4.13.0's features2d/src/orb.cpp or
// keypoint.cpp with the float / double types written there (the reference reaches them through cv::ORB::create(max_num), feature_matching.cpp:16).
#pragma once
#include <algorithm>
#include <cmath>
#include <vector>

namespace esfm {
namespace orbhost {

constexpr int kLevels = 8;

// ORB::create takes scaleFactor as float 1.2f and keeps it in a double
inline double scale_factor() { return (double)1.2f; }

// orb.cpp getScale: (float)pow(scaleFactor, level - firstLevel)
inline float level_scale(int level) { return (float)std::pow(scale_factor(), (double)level); }

// orb.cpp detectAndCompute: Size sz(cvRound(image.cols * inv_scale), cvRound(image.rows * inv_scale)) with float inv_scale = 1.f / scale
inline int level_extent(int full, int level) {
    if (level == 0) return full;
    const float inv = 1.f / level_scale(level);
    return (int)std::lrint((double)((float)full * inv));
}

// orb.cpp computeKeyPoints: the geometric split of nfeatures over the levels
inline void features_per_level(int max_features, int* per_level) {
    const float factor = (float)(1.0 / scale_factor());
    float desired = max_features * (1 - factor) / (1 - (float)std::pow((double)factor, (double)kLevels));
    int sum = 0;
    for (int l = 0; l < kLevels - 1; ++l) {
        per_level[l] = (int)std::lrint((double)desired);
        sum += per_level[l];
        desired *= factor;
    }
    per_level[kLevels - 1] = std::max(max_features - sum, 0);
}

// KeyPointsFilter::retainBest (keypoint.cpp): std::nth_element on "response greater", then std::partition of the tail on
// "response >= the n-th response", so every tie of the boundary response survives.  The order the survivors are left in is the C++
// library's, and it is observable (it becomes the row order of the descriptors), so it is replayed with the same two library calls.
struct RespItem { float response; int index; };
inline void retain_best(std::vector<RespItem>& v, int n_points) {
    if (n_points >= 0 && v.size() > (size_t)n_points) {
        if (n_points == 0) { v.clear(); return; }
        std::nth_element(v.begin(), v.begin() + n_points - 1, v.end(), [](const RespItem& a, const RespItem& b) { return a.response > b.response; });
        const float ambiguous = v[n_points - 1].response;
        auto new_end = std::partition(v.begin() + n_points, v.end(), [ambiguous](const RespItem& a) { return a.response >= ambiguous; });
        v.resize(new_end - v.begin());
    }
}

// What orb.cpp makes of a corner at integer level coordinates (x, y): the key point (computeKeyPoints: pt *= scale, size = patchSize * scale)
// and what computeOrbDescriptors re-derives from that key point -- the level position cvRound(pt * (1.f / scale)) and cos / sin of the angle.
struct KeyPointOut { float x, y, size; };
struct SampleFrame { int cx, cy; float a, b; };
inline KeyPointOut keypoint_of(int x, int y, float scale) { return KeyPointOut{(float)x * scale, (float)y * scale, 31.f * scale}; }
inline SampleFrame sample_frame_of(const KeyPointOut& k, float angle_deg, float scale) {
    const float inv = 1.f / scale;
    SampleFrame s;
    s.cx = (int)std::lrint((double)(k.x * inv));
    s.cy = (int)std::lrint((double)(k.y * inv));
    float ang = angle_deg;
    ang *= (float)(3.141592653589793238462643383279502884 / 180.f);
    s.a = (float)std::cos((double)ang);
    s.b = (float)std::sin((double)ang);
    return s;
}

// HarrisResponses' scale (orb.cpp): scale = 1.f / ((1 << 2) * blockSize * 255.f), to the fourth power
inline float harris_scale4() {
    const float s = 1.f / ((1 << 2) * 7 * 255.f);
    return s * s * s * s;
}

// getGaussianKernel(7, 2, CV_32F): exp(-x^2 / (2 sigma^2)) in double, normalised, cast
inline void gaussian_kernel_7(float* k) {
    double kd[7], sum = 0;
    for (int i = 0; i < 7; ++i) { const double x = i - 3.0; kd[i] = std::exp(-(x * x) / 8.0); sum += kd[i]; }
    for (int i = 0; i < 7; ++i) k[i] = (float)(kd[i] / sum);
}

}  // namespace orbhost
}  // namespace esfm

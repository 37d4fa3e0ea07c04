// Host build of easysfm_b200/csrc/orb_host.h for the CPU tests (tests/test_orb_host.py): the SAME functions orb.cu calls between its
// kernels, compiled with g++, so that the level geometry, the per-level budgets, the retainBest replay and the key-point / sampling-frame
// arithmetic can be compared with oracle/orb_oracle.py (itself pinned to cv2) without a GPU.  Test infrastructure: the product does not load it.
#include "../../easysfm_b200/csrc/orb_host.h"

using namespace esfm::orbhost;

extern "C" {
void obh_levels(int cols, int rows, float* scale, int* w, int* h) {
    for (int l = 0; l < kLevels; ++l) { scale[l] = level_scale(l); w[l] = level_extent(cols, l); h[l] = level_extent(rows, l); }
}
void obh_features_per_level(int max_features, int* out) { features_per_level(max_features, out); }
int obh_retain_best(const float* responses, int n, int n_points, int* out_index) {
    std::vector<RespItem> v((size_t)n);
    for (int i = 0; i < n; ++i) v[(size_t)i] = RespItem{responses[i], i};
    retain_best(v, n_points);
    for (size_t i = 0; i < v.size(); ++i) out_index[i] = v[i].index;
    return (int)v.size();
}
// out: x, y, size, a, b (floats) | cx, cy (ints)
void obh_keypoint(int x, int y, int level, float angle_deg, float* fout, int* iout) {
    const float sc = level_scale(level);
    const KeyPointOut k = keypoint_of(x, y, sc);
    const SampleFrame s = sample_frame_of(k, angle_deg, sc);
    fout[0] = k.x; fout[1] = k.y; fout[2] = k.size; fout[3] = s.a; fout[4] = s.b;
    iout[0] = s.cx; iout[1] = s.cy;
}
float obh_harris_scale4() { return harris_scale4(); }
void obh_gaussian_kernel_7(float* k) { gaussian_kernel_7(k); }
}

// tests/shim/feature_matching.h -- MINIMAL STAND-IN for EasySFM's cpp_code/include/{feature_matching.h,utility.h} and the
// slice of OpenCV they pull in, so easysfm_b200/shim/feature_matching_gpu.cpp can be compiled and run in an image that
// has no OpenCV / PCL / Eigen.  Declarations only; names, member order and defaults follow the reference
// (feature_matching.h:8-30, utility.h:21-54) and OpenCV's public cv::Mat / cv::DMatch layout for the members the shim uses.
#ifndef TESTS_SHIM_FEATURE_MATCHING_H_
#define TESTS_SHIM_FEATURE_MATCHING_H_
#include <cstddef>
#include <string>
#include <vector>

#define CV_8UC1 0
#define CV_32FC1 5

namespace cv {
struct MatStep {
    size_t v = 0;
    operator size_t() const { return v; }
};
struct Mat {
    int flags_type = CV_8UC1;
    int rows = 0, cols = 0;
    unsigned char* data = nullptr;
    MatStep step;
    int type() const { return flags_type; }
};
struct DMatch {
    int queryIdx = -1, trainIdx = -1, imgIdx = -1;
    float distance = 0.f;
};
struct Point2f {
    float x = 0.f, y = 0.f;
};
struct KeyPoint {
    Point2f pt;
};
}  // namespace cv

// the slice of Eigen the two-view shim touches: fixed-size float matrices with element access (utility.h:36-37, estimate_motion.h:17-27)
namespace Eigen {
template <int N> struct MatrixNf {
    float m[N][N] = {};
    float& operator()(int r, int c) { return m[r][c]; }
    float operator()(int r, int c) const { return m[r][c]; }
};
typedef MatrixNf<3> Matrix3f;
typedef MatrixNf<4> Matrix4f;
}  // namespace Eigen

namespace p3dv {
struct frame_t {
    unsigned int frame_id = 0;
    std::string image_file_path;
    std::vector<cv::KeyPoint> keypoints;
    cv::Mat descriptors;
    std::vector<int> unique_pixel_ids;
    Eigen::Matrix3f K_cam;       // intrinsic elements of the camera (utility.h:37)
};

class FeatureMatching {
public:
    bool matchFeaturesORB(frame_t& cur_frame_1, frame_t& cur_frame_2, std::vector<cv::DMatch>& matches,
                          double ratio_thre = 0.8, bool show = false);
    bool matchFeaturesSURF(frame_t& cur_frame_1, frame_t& cur_frame_2, std::vector<cv::DMatch>& matches,
                           double ratio_thre = 0.5, bool show = false);
};
bool esfm_prepare_all_pairs(std::vector<frame_t>& frames, char feature, double ratio_thre, bool cross_check);
bool esfm_save_matches(const char* path);
bool esfm_load_matches(std::vector<frame_t>& frames, char feature, double ratio_thre, const char* path);
}  // namespace p3dv
#endif

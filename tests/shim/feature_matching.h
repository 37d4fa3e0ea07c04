// tests/shim/feature_matching.h -- MINIMAL STAND-IN for EasySFM's cpp_code/include/{feature_matching.h,utility.h} and the
// slice of OpenCV they pull in, so easysfm_b200/shim/feature_matching_gpu.cpp can be compiled and run in an image that
// has no OpenCV / PCL / Eigen.  Declarations only; names, member order and defaults follow the reference
// (feature_matching.h:8-30, utility.h:21-54) and OpenCV's public cv::Mat / cv::DMatch layout for the members the shim uses.
#ifndef TESTS_SHIM_FEATURE_MATCHING_H_
#define TESTS_SHIM_FEATURE_MATCHING_H_
#include <cstddef>
#include <memory>
#include <string>
#include <vector>

#define CV_8UC1 0
#define CV_8UC3 16
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
    std::shared_ptr<std::vector<unsigned char> > owned;   // cv::Mat's reference-counted buffer, for create()
    int type() const { return flags_type; }
    int channels() const { return (flags_type >> 3) + 1; }
    bool empty() const { return rows == 0 || cols == 0 || !data; }
    void create(int r, int c, int type) {                 // cv::Mat::create: a continuous r x c matrix of `type`
        const size_t elem = (size_t)((type >> 3) + 1) * ((type & 7) == 5 ? 4 : 1);
        owned = std::make_shared<std::vector<unsigned char> >((size_t)r * c * elem + 1);
        flags_type = type; rows = r; cols = c; data = owned->data(); step.v = (size_t)c * elem;
    }
};
struct DMatch {
    int queryIdx = -1, trainIdx = -1, imgIdx = -1;
    float distance = 0.f;
};
struct Point2f {
    float x = 0.f, y = 0.f;
};
struct KeyPoint {             // member order of cv::KeyPoint
    Point2f pt;
    float size = 0.f, angle = -1.f, response = 0.f;
    int octave = 0, class_id = -1;
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
    cv::Mat rgb_image;           // utility.h:25
    std::vector<cv::KeyPoint> keypoints;
    cv::Mat descriptors;
    std::vector<int> unique_pixel_ids;
    Eigen::Matrix3f K_cam;       // intrinsic elements of the camera (utility.h:37)
};

class FeatureMatching {
public:
    bool detectFeaturesORB(frame_t& cur_frame, int max_num = 5000, bool show = false);
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

// tests/shim/estimate_motion.h -- MINIMAL STAND-IN for EasySFM's cpp_code/include/estimate_motion.h: the two methods of p3dv::MotionEstimator
// that easysfm_b200/shim/estimate_motion_gpu.cpp implements, with the reference's signatures and defaults (estimate_motion.h:17-20, :26-27).
#ifndef TESTS_SHIM_ESTIMATE_MOTION_H_
#define TESTS_SHIM_ESTIMATE_MOTION_H_
#include "feature_matching.h"

namespace p3dv {
class MotionEstimator {
public:
    bool estimate2D2D_E5P_RANSAC(frame_t& cur_frame_1, frame_t& cur_frame_2, std::vector<cv::DMatch>& matches, std::vector<cv::DMatch>& inlier_matches,
                                 Eigen::Matrix4f& T, double ransac_thre = 1.0, double ransac_prob = 0.99, bool show = false);
    bool getDepthFast(frame_t& cur_frame_1, frame_t& cur_frame_2, Eigen::Matrix4f& T_21, const std::vector<cv::DMatch>& matches, double& appro_depth,
                      int random_rate = 20);
};
}  // namespace p3dv
#endif

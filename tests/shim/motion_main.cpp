// Drives the two-view shim the way cpp_code/test/sfm.cpp:163-166 drives the reference: estimate2D2D_E5P_RANSAC then getDepthFast per pair.
// usage: motion_main <in file> <out file>
//   in : int32 n_pairs; double K[9]; per pair: int32 n_kp1, n_kp2, n_matches; float kp1[n_kp1][2]; float kp2[n_kp2][2]; DMatch matches[n_matches]
//   out: per pair: int32 ok, n_inliers; float T[16]; double depth; DMatch inliers[n_inliers]
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "estimate_motion.h"

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    FILE* f = std::fopen(argv[1], "rb");
    FILE* o = std::fopen(argv[2], "wb");
    if (!f || !o) return 3;
    int n_pairs = 0;
    double K[9];
    if (std::fread(&n_pairs, 4, 1, f) != 1 || std::fread(K, 8, 9, f) != 9) return 4;
    p3dv::MotionEstimator ee;
    for (int p = 0; p < n_pairs; ++p) {
        int n1, n2, nm;
        if (std::fread(&n1, 4, 1, f) != 1 || std::fread(&n2, 4, 1, f) != 1 || std::fread(&nm, 4, 1, f) != 1) return 4;
        p3dv::frame_t f1, f2;
        f1.frame_id = 10 + 2 * p;
        f2.frame_id = 11 + 2 * p;
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) f1.K_cam(r, c) = f2.K_cam(r, c) = (float)K[3 * r + c];
        f1.keypoints.resize(n1);
        f2.keypoints.resize(n2);
        for (int i = 0; i < n1; ++i) if (std::fread(&f1.keypoints[i].pt, 4, 2, f) != 2) return 4;
        for (int i = 0; i < n2; ++i) if (std::fread(&f2.keypoints[i].pt, 4, 2, f) != 2) return 4;
        std::vector<cv::DMatch> matches(nm), inliers;
        if (nm && std::fread(matches.data(), sizeof(cv::DMatch), nm, f) != (size_t)nm) return 4;
        Eigen::Matrix4f T;
        double depth = 0.0;
        const bool ok = ee.estimate2D2D_E5P_RANSAC(f1, f2, matches, inliers, T, 1.0);
        if (ok && !inliers.empty()) ee.getDepthFast(f1, f2, T, inliers, depth);
        const int iok = ok ? 1 : 0, ni = (int)inliers.size();
        std::fwrite(&iok, 4, 1, o);
        std::fwrite(&ni, 4, 1, o);
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) { const float v = T(r, c); std::fwrite(&v, 4, 1, o); }
        std::fwrite(&depth, 8, 1, o);
        if (ni) std::fwrite(inliers.data(), sizeof(cv::DMatch), ni, o);
    }
    std::fclose(f);
    std::fclose(o);
    return 0;
}

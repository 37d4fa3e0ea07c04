// Drives FeatureMatching::detectFeaturesORB of the C++ shim the way cpp_code/test/sfm.cpp:116 does.
// usage: orb_main <raw image file> <rows> <cols> <channels> <max_num> <out file>
// out: int32 n | n x (x, y, size, angle, response: float32; octave, class_id: int32) | n x 32 descriptor bytes
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "feature_matching.h"

int main(int argc, char** argv) {
    if (argc < 7) return 2;
    const int rows = std::atoi(argv[2]), cols = std::atoi(argv[3]), ch = std::atoi(argv[4]), max_num = std::atoi(argv[5]);
    p3dv::frame_t frame;
    frame.rgb_image.create(rows, cols, ch == 3 ? CV_8UC3 : CV_8UC1);
    FILE* f = std::fopen(argv[1], "rb");
    const size_t bytes = (size_t)rows * cols * ch;
    if (!f || std::fread(frame.rgb_image.data, 1, bytes, f) != bytes) return 3;
    std::fclose(f);
    p3dv::FeatureMatching fm;
    if (!fm.detectFeaturesORB(frame, max_num)) return 4;
    const int n = (int)frame.keypoints.size();
    if (frame.descriptors.rows != n || (n && (frame.descriptors.cols != 32 || frame.descriptors.type() != CV_8UC1))) return 5;
    FILE* o = std::fopen(argv[6], "wb");
    std::fwrite(&n, sizeof n, 1, o);
    for (const cv::KeyPoint& k : frame.keypoints) {
        const float v[5] = {k.pt.x, k.pt.y, k.size, k.angle, k.response};
        const int w[2] = {k.octave, k.class_id};
        std::fwrite(v, sizeof v, 1, o);
        std::fwrite(w, sizeof w, 1, o);
    }
    if (n) std::fwrite(frame.descriptors.data, 32, (size_t)n, o);
    std::fclose(o);
    return 0;
}

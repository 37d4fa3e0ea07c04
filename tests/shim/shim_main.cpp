// Drives the C++ shim the way cpp_code/test/sfm.cpp:140-161 drives the reference: for i, for j < i: matchFeaturesX(frames[i], frames[j], out).
// usage: shim_main <O|S> <n_frames> <rows_0> ... <rows_{n-1}> <descriptor file (raw, frames back to back)> <mode> <out file>
//   mode 0: per-call path   1: esfm_prepare_all_pairs first   2: prepare + esfm_save_matches(<out file>.matches)
//   mode 3: esfm_load_matches(<out file>.matches) instead of matching (no device needed)
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "feature_matching.h"

int main(int argc, char** argv) {
    if (argc < 6) return 2;
    const char feature = argv[1][0];
    const int n = std::atoi(argv[2]);
    std::vector<int> rows(n);
    for (int i = 0; i < n; ++i) rows[i] = std::atoi(argv[3 + i]);
    const char* path = argv[3 + n];
    const int mode = std::atoi(argv[4 + n]);
    const bool prepare = mode == 1 || mode == 2;
    const char* out_path = argv[5 + n];
    const size_t rb = feature == 'O' ? 32 : 256;
    size_t total = 0;
    for (int r : rows) total += (size_t)r;
    std::vector<unsigned char> blob(total * rb + 1);
    FILE* f = std::fopen(path, "rb");
    if (!f || std::fread(blob.data(), 1, total * rb, f) != total * rb) return 3;
    std::fclose(f);
    std::vector<p3dv::frame_t> frames(n);
    size_t off = 0;
    for (int i = 0; i < n; ++i) {
        frames[i].frame_id = 100 + i;  // ids need not be 0..n-1
        frames[i].descriptors.flags_type = feature == 'O' ? CV_8UC1 : CV_32FC1;
        frames[i].descriptors.rows = rows[i];
        frames[i].descriptors.cols = feature == 'O' ? 32 : 64;
        frames[i].descriptors.data = blob.data() + off;
        frames[i].descriptors.step.v = rb;
        off += (size_t)rows[i] * rb;
    }
    p3dv::FeatureMatching fm;
    if (prepare && !p3dv::esfm_prepare_all_pairs(frames, feature, feature == 'O' ? 0.8 : 0.5, false)) return 4;
    const std::string match_file = std::string(out_path) + ".matches";
    if (mode == 2 && !p3dv::esfm_save_matches(match_file.c_str())) return 6;
    if (mode == 3 && !p3dv::esfm_load_matches(frames, feature, feature == 'O' ? 0.8 : 0.5, match_file.c_str())) return 7;
    FILE* o = std::fopen(out_path, "wb");
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < i; ++j) {
            std::vector<cv::DMatch> m;
            const bool ok = feature == 'O' ? fm.matchFeaturesORB(frames[i], frames[j], m) : fm.matchFeaturesSURF(frames[i], frames[j], m);
            if (!ok) return 5;
            const int cnt = (int)m.size();
            std::fwrite(&cnt, sizeof cnt, 1, o);
            if (cnt) std::fwrite(m.data(), sizeof(cv::DMatch), m.size(), o);
        }
    std::fclose(o);
    return 0;
}

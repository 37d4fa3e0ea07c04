/*
 * oracle/tracks_oracle.c -- TEST INFRASTRUCTURE ONLY (the checker, never the product).
 *
 * LITERAL CPU restatement of the reference's track building and co-visibility scoring, the
 * step that consumes the pair matches right after the matching hot path:
 *   - unique-id propagation over the pair loop      cpp_code/test/sfm.cpp:140-217
 *       (for i, for j < i: inlier matches of (i, j) label frame i's keypoints with frame j's ids unless the id is
 *        already used in frame i -- found by a LINEAR scan of frame i, :181-188; unlabeled keypoints get fresh ids
 *        after the row, :199-213; the feature track matrix is a dense frames x points bool matrix, :136,210)
 *   - initial frame pair                             cpp_code/src/feature_matching.cpp:160-233
 *       (per pair: sum over the points seen by BOTH frames of the number of frames that see the point; pairs whose
 *        depth / baseline ratio exceeds the limit are skipped; '>=' so the LAST best pair in loop order wins)
 *   - next frame                                     cpp_code/src/feature_matching.cpp:235-268
 *       (the unprocessed frame that sees most of the current 3D points; strict '>' so the FIRST best frame wins)
 * Integer work throughout: the product (easysfm_b200/csrc/tracks.cu) must reproduce it bit for bit.
 *
 * The reference has no tests or golden vectors for this step and the code is plain C++ over
 * std::vector -- no third-party arithmetic: this file follows it line by line (dense bool matrix, linear duplicate
 * scans) and is therefore only usable at small sizes; tests/test_tracks.py pins the product against it on seeded
 * random match graphs.  PINNING: parity unpinned by reference fixtures (none exist); pinned by construction (literal
 * restatement of ~80 lines of integer code).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int32_t queryIdx, trainIdx, imgIdx;
    float distance;
} dmatch_t;

/* sfm.cpp:140-217.  kp[f] = keypoints of frame f; pair p = i (i - 1) / 2 + j holds n_in[p] inlier matches starting at
 * in_off[p] in `inl`.  Outputs: ids (frames back to back, frame f at kp_off[f]), has_match, and the dense track
 * matrix track[f * total_kp + id].  Returns the number of unique points. */
int64_t oracle_tracks_build(int n_frames, const int32_t* kp, const int64_t* in_off, const int32_t* n_in, const dmatch_t* inl,
                            int32_t* ids, uint8_t* has_match, uint8_t* track) {
    int64_t total_kp = 0;
    int64_t* kp_off = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n_frames + 1));
    for (int f = 0; f < n_frames; ++f) { kp_off[f] = total_kp; total_kp += kp[f]; }
    kp_off[n_frames] = total_kp;
    for (int64_t k = 0; k < total_kp; ++k) { ids[k] = -1; has_match[k] = 0; }      /* utility.h:44-51 init_pixel_ids */
    if (track) memset(track, 0, (size_t)n_frames * (size_t)total_kp);
    int64_t cur = 0;                                                                /* sfm.cpp:138 */
    for (int i = 0; i < n_frames; ++i) {                                            /* :140 */
        int32_t* idi = ids + kp_off[i];
        for (int j = 0; j < i; ++j) {                                               /* :143 */
            const int64_t p = (int64_t)i * (i - 1) / 2 + j;
            const int32_t* idj = ids + kp_off[j];
            for (int k = 0; k < n_in[p]; ++k) {                                     /* :172 */
                const dmatch_t m = inl[in_off[p] + k];
                if (idi[m.queryIdx] < 0 || idi[m.queryIdx] != idj[m.trainIdx]) {    /* :174-175 */
                    int dup = 0;
                    for (int q = 0; q < kp[i]; ++q)                                 /* :178-185 check duplication */
                        if (idj[m.trainIdx] == idi[q]) { dup = 1; break; }
                    if (!dup) {                                                     /* :186-190 */
                        idi[m.queryIdx] = idj[m.trainIdx];
                        has_match[kp_off[i] + m.queryIdx] = 1;
                    }
                }
            }
        }
        int64_t fresh = 0;                                                          /* :200-213 */
        for (int k = 0; k < kp[i]; ++k) {
            if (idi[k] < 0) { idi[k] = (int32_t)(cur + fresh); ++fresh; }
            if (track) track[(size_t)i * (size_t)total_kp + (size_t)idi[k]] = 1;
        }
        cur += fresh;
    }
    free(kp_off);
    return cur;
}

/* feature_matching.cpp:160-233 on the dense matrix.  depth[p] = img_match_graph[i][j].appro_depth.
 * Returns 1 and (f1, f2, depth_init, best) when a pair was found, else 0 with the default (1, 0). */
int oracle_find_init_pair(int n_frames, int64_t n_points, const uint8_t* track, const double* depth, int min_track_num_init,
                          double max_depth_baseline_ratio_init, int* f1, int* f2, double* depth_init, int64_t* best) {
    int* cnt = (int*)calloc((size_t)n_points, sizeof(int));
    for (int i = 0; i < n_frames; ++i)
        for (int64_t k = 0; k < n_points; ++k) cnt[k] += track[(size_t)i * (size_t)n_points + (size_t)k];      /* :177-185 */
    *f1 = 0; *f2 = 0;
    int max_sum = min_track_num_init;                                                                          /* :188 */
    double ratio_init = 0.0;
    for (int i = 0; i < n_frames; ++i)
        for (int j = 0; j < i; ++j) {
            const double r = depth[(int64_t)i * (i - 1) / 2 + j];
            if (r > max_depth_baseline_ratio_init) continue;                                                   /* :198-199 */
            int sum = 0;
            for (int64_t k = 0; k < n_points; ++k)
                if (track[(size_t)i * (size_t)n_points + (size_t)k] && track[(size_t)j * (size_t)n_points + (size_t)k]) sum += cnt[k];
            if (sum >= max_sum) { max_sum = sum; ratio_init = r; *f1 = i; *f2 = j; }                            /* :208-214 */
        }
    free(cnt);
    *best = max_sum;
    if (*f1 == *f2) {                                                                                          /* :219-226 */
        *f1 = 1; *f2 = 0;
        *depth_init = n_frames > 1 ? depth[0] : 0.0;
        return 0;
    }
    *depth_init = ratio_init;
    return 1;
}

/* feature_matching.cpp:235-268.  next_frame is left untouched when no frame sees any point (as in the reference). */
int oracle_find_next_frame(int n_frames, int64_t n_points, const uint8_t* track, const uint8_t* to_process, const int32_t* point_ids,
                           int64_t n_ids, int* next_frame) {
    int max_common = 0;
    for (int i = 0; i < n_frames; ++i) {
        if (!to_process[i]) continue;
        int common = 0;
        for (int64_t j = 0; j < n_ids; ++j)
            if (track[(size_t)i * (size_t)n_points + (size_t)point_ids[j]]) ++common;
        if (common > max_common) { max_common = common; *next_frame = i; }
    }
    return max_common;
}

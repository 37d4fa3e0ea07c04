"""Track building + co-visibility scoring at the BASELINE configs[3] shape: 1000 frames x 8000 keypoints.
  * scoring kernel (findInitializeFramePair's O(N^2 P) loop, feature_matching.cpp:193-215): all 499,500 pairs on the device;
  * unique-id propagation (sfm.cpp:172-217) on the host: matches per second on a full pair graph of a smaller frame set;
  * the reference's literal loops (oracle/tracks_oracle.c) timed on a small case and extrapolated (dense bool matrix: O(N^2 P)).
usage: python tools/tracks_bench.py [n_frames] [keypoints]      -> one JSON line"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import easysfm_b200 as esfm
import oracle

n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
n_kp = int(sys.argv[2]) if len(sys.argv) > 2 else 8000
rng = np.random.default_rng(7)
W = 4 * n_kp


def observe(n):
    return [rng.choice(W, size=n_kp, replace=False) for _ in range(n)]


def link(a, b):
    common, qi, tj = np.intersect1d(a, b, return_indices=True)
    order = np.argsort(qi, kind="stable")
    m = np.zeros(len(common), esfm.DMATCH_DTYPE)
    m["queryIdx"], m["trainIdx"] = qi[order], tj[order]
    return m

# ---- scoring kernel at full size: a chain graph labels the frames (the kernel's work does not depend on the labels) ----
lm = observe(n_frames)
t = esfm.Tracks([n_kp] * n_frames)
for i in range(n_frames):
    if i:
        t.add_pair(i, i - 1, link(lm[i], lm[i - 1]))
    t.finish_frame(i)
ctx = esfm.Context(0)
scores, ms = t.pair_scores(ctx)          # warm-up (uploads the id lists)
best = min(t.pair_scores(ctx)[1] for _ in range(5))
n_pairs = n_frames * (n_frames - 1) // 2
found, f1, f2, depth, best_score = t.find_init_pair(ctx, None, min_track_num_init=100)
searches = n_pairs * n_kp
lds = searches * np.ceil(np.log2(n_kp))
sms = ctx.sm_count

# ---- host labelling rate on a full pair graph (every pair has matches) ----
n_small = 120
lm2 = observe(n_small)
pairs = [link(lm2[i], lm2[j]) for i in range(n_small) for j in range(i)]
n_matches = sum(len(m) for m in pairs)
t0 = time.perf_counter()
t2 = esfm.Tracks([n_kp] * n_small)
k = 0
for i in range(n_small):
    for j in range(i):
        t2.add_pair(i, j, pairs[k]); k += 1
    t2.finish_frame(i)
host_s = time.perf_counter() - t0

# ---- the reference's literal loops on a small case ----
n_ref, kp_ref = 24, 1500
lm3 = [rng.choice(4 * kp_ref, size=kp_ref, replace=False) for _ in range(n_ref)]
pairs3 = [link(lm3[i], lm3[j]) for i in range(n_ref) for j in range(i)]
t0 = time.perf_counter()
ids_ref, has_ref, track, npts = oracle.tracks_build([kp_ref] * n_ref, pairs3)
ref_build_s = time.perf_counter() - t0
t0 = time.perf_counter()
oracle.find_init_pair(track, np.ones(len(pairs3)), 100)
ref_score_s = time.perf_counter() - t0
steps = (n_ref * (n_ref - 1) // 2) * track.shape[1]
P_full = t.counts()[1]
print(json.dumps({
    "shape": f"{n_frames} frames x {n_kp} keypoints, {n_pairs} pairs, {P_full} unique points",
    "covis_kernel_ms": best, "pairs_per_s": n_pairs / (best * 1e-3), "id_lookups_per_s": searches / (best * 1e-3),
    "bound": "shared-memory reads of the binary search (ceil(log2 F) dependent LDS per looked-up id; 32 lanes/clk/SM)",
    "lds_lane_ops_per_s": lds / (best * 1e-3), "lds_peak_lane_ops_per_s": sms * 32 * 1.965e9, "frac_of_lds_peak": lds / (best * 1e-3) / (sms * 32 * 1.965e9),
    "l2_to_sm_bytes_algorithmic": int(n_pairs * n_kp * 4 * (1 + 1 / 8)), "init_pair": [found, f1, f2, best_score],
    "host_labelling": {"frames": n_small, "pairs": len(pairs), "matches": int(n_matches), "seconds": host_s, "matches_per_s": n_matches / host_s,
                       "note": "esfm_tracks_add_pair / finish_frame through ctypes, one call per pair (hash lookup instead of the reference's linear duplicate scan)"},
    "reference_literal": {"case": f"{n_ref} frames x {kp_ref} keypoints", "tracks_build_s": ref_build_s, "find_init_pair_s": ref_score_s,
                          "bool_steps_per_s": steps / ref_score_s,
                          "extrapolated_find_init_pair_s_at_full_shape": n_pairs * float(n_frames * n_kp) / (steps / ref_score_s),
                          "note": "dense frames x total-keypoints bool matrix walked per pair (feature_matching.cpp:201-207); 1 host core"},
}))

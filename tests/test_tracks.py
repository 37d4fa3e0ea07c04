"""Track building and co-visibility against the literal restatement of the reference's loops
(oracle/tracks_oracle.c: cpp_code/test/sfm.cpp:140-217, cpp_code/src/feature_matching.cpp:160-268).
CPU: unique-id propagation and findNextFrame are host code in the library -- bit-exact on seeded random match graphs with
repeated keypoints, conflicting labels, empty pairs and empty frames.  GPU: the pair-score kernel and findInitializeFramePair."""
import numpy as np
import pytest

import easysfm_b200 as esfm
import oracle


def random_graph(seed, kp, p_pair=0.7, max_matches=60, landmarks=None):
    """Inlier matches per pair in loop order.  With `landmarks`, keypoints observe world points and matches link equal
    landmarks (consistent tracks) plus a few random wrong links (conflicts that the duplicate rule must resolve)."""
    rng = np.random.default_rng(seed)
    n = len(kp)
    lm = [rng.choice(landmarks, size=k, replace=False) if landmarks else None for k in kp]
    pairs = []
    for i in range(n):
        for j in range(i):
            if kp[i] == 0 or kp[j] == 0 or rng.random() > p_pair:
                pairs.append(np.zeros(0, esfm.DMATCH_DTYPE))
                continue
            if landmarks:
                common, qi, tj = np.intersect1d(lm[i], lm[j], return_indices=True)
                sel = rng.random(len(common)) < 0.8
                q, t = qi[sel], tj[sel]
                nbad = int(rng.integers(0, 4))
                q = np.concatenate([q, rng.integers(0, kp[i], nbad)])
                t = np.concatenate([t, rng.integers(0, kp[j], nbad)])
            else:
                m = int(rng.integers(0, max_matches))
                q = rng.integers(0, kp[i], m)                 # repeated query keypoints on purpose
                t = rng.integers(0, kp[j], m)
            order = np.argsort(q, kind="stable")              # the matcher's output order: ascending queryIdx
            mm = np.zeros(len(q), esfm.DMATCH_DTYPE)
            mm["queryIdx"], mm["trainIdx"] = q[order], t[order]
            pairs.append(mm)
    return pairs


def build_lib(kp, pairs):
    t = esfm.Tracks(kp)
    k = 0
    for i in range(len(kp)):
        for j in range(i):
            if len(pairs[k]):
                t.add_pair(i, j, pairs[k])
            k += 1
        t.finish_frame(i)
    return t


@pytest.mark.parametrize("seed,kp,landmarks", [(1, [30, 25, 0, 40, 33, 28], None), (2, [50] * 9, 120), (3, [80, 1, 64, 70, 75, 2, 90], 200),
                                               (4, [12, 12, 12], None), (5, [200, 180, 220, 190, 210, 205, 195, 185], 600)])
def test_unique_id_propagation_equals_reference_loops(seed, kp, landmarks):
    pairs = random_graph(seed, kp, landmarks=landmarks)
    ids_ref, has_ref, track, npts_ref = oracle.tracks_build(kp, pairs)
    t = build_lib(kp, pairs)
    done, npts = t.counts()
    assert done == len(kp) and npts == npts_ref
    for f in range(len(kp)):
        ids, has = t.frame(f)
        np.testing.assert_array_equal(ids, ids_ref[f])
        np.testing.assert_array_equal(has, has_ref[f])
        assert len(set(ids.tolist())) == len(ids)             # an id occurs at most once per frame
    # findNextFrame on random processed sets / point sets
    rng = np.random.default_rng(seed + 100)
    for _ in range(6):
        todo = (rng.random(len(kp)) < 0.6).astype(np.uint8)
        pts = rng.integers(0, max(npts, 1), int(rng.integers(0, 80))).astype(np.int32)   # duplicates allowed, as in the reference
        assert t.find_next_frame(todo, pts, next_frame=-7) == oracle.find_next_frame(track[:, :max(npts, 1)] if npts else track, todo, pts, next_frame=-7)


def test_pairs_must_arrive_in_loop_order():
    t = esfm.Tracks([5, 5, 5])
    t.finish_frame(0)
    m = np.zeros(1, esfm.DMATCH_DTYPE)
    with pytest.raises(esfm.EsfmError):
        t.add_pair(2, 0, m)                                   # frame 1 is in progress
    t.add_pair(1, 0, m)
    with pytest.raises(esfm.EsfmError):
        t.add_pair(1, 0, m)                                   # train frames must ascend
    bad = m.copy(); bad["queryIdx"] = 9
    t.finish_frame(1)
    with pytest.raises(esfm.EsfmError):
        t.add_pair(2, 0, bad)                                 # keypoint index out of range
    with pytest.raises(esfm.EsfmError):
        t.finish_frame(1)


@pytest.mark.gpu
@pytest.mark.parametrize("seed,kp,landmarks", [(11, [60, 55, 0, 70, 65, 58, 61], 150), (12, [300] * 12, 900), (13, [40, 45], 60), (14, [500, 30, 700, 2, 650, 640, 10, 620, 600], 1500)])
def test_pair_scores_and_init_pair_equal_reference(ctx, seed, kp, landmarks):
    pairs = random_graph(seed, kp, landmarks=landmarks)
    ids_ref, has_ref, track, npts = oracle.tracks_build(kp, pairs)
    t = build_lib(kp, pairs)
    n = len(kp)
    scores, ms = t.pair_scores(ctx)
    cnt = track.sum(axis=0).astype(np.int64)
    ref = np.array([int((cnt * (track[i] & track[j])).sum()) for i in range(n) for j in range(i)], np.int64)
    np.testing.assert_array_equal(scores, ref)
    rng = np.random.default_rng(seed)
    depth = rng.uniform(1.0, 80.0, len(ref))                  # some pairs beyond the 50.0 baseline limit
    for min_tracks in (0, 100, int(ref.max()) if len(ref) else 0, 10 ** 9):
        got = t.find_init_pair(ctx, depth, min_track_num_init=min_tracks)
        want = oracle.find_init_pair(track, depth, min_track_num_init=min_tracks)
        assert got == want
    assert t.find_init_pair(ctx, None, min_track_num_init=1)[:3] == oracle.find_init_pair(track, np.ones(len(ref)), min_track_num_init=1)[:3]


@pytest.mark.gpu
def test_tracks_from_gpu_matches(ctx):
    """The whole chain on real matcher output: all-pairs matches -> esfm_tracks_build (pairs above 20 matches) == the reference's
    loops fed with the same matches; keypoints observing the same landmark end up with the same id."""
    from easysfm_b200 import synth
    rows = [400, 380, 0, 420, 390, 410]
    frames = synth.orb_like(len(rows), rows, seed=41)
    res = ctx.bank_from_frames(frames).match_all_pairs(0.8, True)
    t = esfm.Tracks(rows)
    t.build(res, min_pair_matches=20)
    pairs = [res.pair_at(k)[2] if len(res.pair_at(k)[2]) > 20 else np.zeros(0, esfm.DMATCH_DTYPE) for k in range(res.n_pairs)]
    ids_ref, has_ref, track, npts = oracle.tracks_build(rows, pairs)
    assert t.counts() == (len(rows), npts)
    for f in range(len(rows)):
        ids, has = t.frame(f)
        np.testing.assert_array_equal(ids, ids_ref[f])
        np.testing.assert_array_equal(has, has_ref[f])
    assert t.find_init_pair(ctx, None, min_track_num_init=100) == oracle.find_init_pair(track, np.ones(res.n_pairs), min_track_num_init=100)

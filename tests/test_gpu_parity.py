"""GPU parity tests proper: the CUDA path, called through the C ABI, against the oracle on the same inputs.

Oracles: oracle.match / oracle.knn2 (C restatement, oracle/bf_oracle.c) and oracle.cv2_oracle (the
reference's own cv2.BFMatcher calls).  Hamming must be bit-exact; L2 within 1e-5 relative (util.REL_TOL).
"""
import numpy as np
import pytest

import oracle
from oracle import cv2_oracle
from easysfm_b200 import synth
from util import assert_matches_equal, check_knn_l2, dist64, justify_l2

pytestmark = pytest.mark.gpu

RATIOS = (0.5, 0.8)


# --------------------------------------------------------------------------------------------------
# ORB / Hamming: bit-exact
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("nq,nt", [(1, 2), (2, 2), (37, 53), (256, 256), (257, 1025), (1500, 700), (2049, 3000)])
def test_hamming_pair_bit_exact(ctx, nq, nt):
    fr = synth.orb_like(2, [nq, nt], seed=nq * 7 + nt)
    Q, T = fr
    for ratio in RATIOS:
        for cc in (False, True):
            got = ctx.match_descriptors(Q, T, ratio, cc)
            ref = oracle.match(Q, T, ratio, cc)
            assert_matches_equal(got, ref)
    ref_cv = cv2_oracle.match(Q, T, 0.8, True)
    assert_matches_equal(ctx.match_descriptors(Q, T, 0.8, True), ref_cv)


def test_hamming_knn2_ties_lowest_index(ctx):
    rng = np.random.default_rng(5)
    T = rng.integers(0, 256, (600, 32), dtype=np.uint8)
    T[7] = T[3]; T[400] = T[3]; T[599] = T[3]          # exact duplicates: ranks 1,2 must be 3 then 7
    Q = T[[3, 10, 20]].copy()
    Q[1, 0] ^= 1
    bank = ctx.bank_from_frames([Q, T])
    idx, dist = bank.knn2_pair(0, 1)
    ridx, rdist = oracle.knn2(Q, T)
    np.testing.assert_array_equal(idx, ridx)
    np.testing.assert_array_equal(dist, rdist)
    assert idx[0].tolist() == [3, 7] and dist[0].tolist() == [0.0, 0.0]
    cidx, cdist = cv2_oracle.knn2(Q, T)
    np.testing.assert_array_equal(idx, cidx)
    np.testing.assert_array_equal(dist, cdist)


def test_hamming_all_pairs_ragged(ctx):
    rows = [900, 0, 1, 2, 1300, 257, 1024]
    fr = synth.orb_like(len(rows), rows, seed=11)
    bank = ctx.bank_from_frames(fr)
    for cc in (False, True):
        res = bank.match_all_pairs(0.8, cc)
        assert res.n_pairs == len(rows) * (len(rows) - 1) // 2
        k = 0
        for i in range(len(rows)):
            for j in range(i):
                q, t, m = res.pair_at(k)
                assert (q, t) == (i, j)
                assert_matches_equal(m, oracle.match(fr[i], fr[j], 0.8, cc))
                assert_matches_equal(res.pair(i, j), m)
                k += 1


# --------------------------------------------------------------------------------------------------
# SURF / L2
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("nq,nt", [(1, 2), (3, 2), (100, 130), (256, 128), (257, 385), (1000, 2000), (2000, 777)])
def test_l2_pair_vs_oracle(ctx, nq, nt):
    Q, T = synth.surf_like(2, [nq, nt], seed=nq + 3 * nt)
    D = dist64(Q, T)
    bank = ctx.bank_from_frames([Q, T])
    idx, dist = bank.knn2_pair(0, 1)
    check_knn_l2(Q, T, idx, dist, D)
    ridx, rdist = oracle.knn2(Q, T)
    same = (idx == ridx).all(axis=1)
    # where the candidates agree the refinement reproduces the oracle's arithmetic bit for bit
    np.testing.assert_array_equal(dist[same], rdist[same])
    assert same.mean() > 0.999
    for ratio in RATIOS:
        for cc in (False, True):
            got = ctx.match_descriptors(Q, T, ratio, cc)
            ref = oracle.match(Q, T, ratio, cc)
            ndiff = justify_l2(Q, T, ratio, cc, got, ref, D)
            assert ndiff <= max(1, len(ref) // 500)
            ref_cv = cv2_oracle.match(Q, T, ratio, cc)
            justify_l2(Q, T, ratio, cc, got, ref_cv, D)


def test_l2_duplicates_and_zero_distance(ctx):
    Q, T = synth.surf_like(2, [300, 500], seed=77)
    T[40] = T[17]; T[300] = T[17]; T[499] = T[17]      # three identical train rows (A1 probe)
    Q[5] = T[17]                                        # d = 0 to all of them
    Q[6] = Q[5]                                         # duplicate queries: only the lower index survives cross-check
    bank = ctx.bank_from_frames([Q, T])
    idx, dist = bank.knn2_pair(0, 1)
    assert idx[5].tolist() == [17, 40] and dist[5].tolist() == [0.0, 0.0]
    ridx, rdist = oracle.knn2(Q, T)
    np.testing.assert_array_equal(idx, ridx)
    np.testing.assert_array_equal(dist, rdist)
    cidx, _ = cv2_oracle.knn2(Q, T)
    np.testing.assert_array_equal(idx, cidx)
    for cc in (False, True):
        got = ctx.match_descriptors(Q, T, 0.8, cc)
        assert_matches_equal(got, oracle.match(Q, T, 0.8, cc))
        ref_cv = cv2_oracle.match(Q, T, 0.8, cc)
        assert_matches_equal(got, ref_cv, exact_distance=False)


def test_l2_all_pairs_config2_subset(ctx):
    """BASELINE configs[1] shape (25 x 2k SURF-like), reduced to 6 frames here; ratios 0.5 / 0.8, cross_check 0 / 1."""
    fr = synth.surf_like(6, 2000, seed=2)
    bank = ctx.bank_from_frames(fr)
    for ratio in RATIOS:
        for cc in (False, True):
            res = bank.match_all_pairs(ratio, cc)
            for k in range(res.n_pairs):
                i, j, m = res.pair_at(k)
                ref = oracle.match(fr[i], fr[j], ratio, cc)
                if len(m) == len(ref) and (m["trainIdx"] == ref["trainIdx"]).all():
                    np.testing.assert_array_equal(m["distance"], ref["distance"])
                else:
                    justify_l2(fr[i], fr[j], ratio, cc, m, ref)


def test_mutual_nn_mode(ctx):
    """ratio = +inf: plain mutual nearest neighbour == cv2.BFMatcher(crossCheck=True).match (feature_match.py:26-27)."""
    Q, T = synth.surf_like(2, [400, 350], seed=9)
    got = ctx.match_descriptors(Q, T, float("inf"), True)
    assert_matches_equal(got, oracle.mutual_nn(Q, T))
    assert_matches_equal(got, cv2_oracle.mutual_nn(Q, T), exact_distance=False)
    Qb, Tb = synth.orb_like(2, [300, 1], seed=4)    # a single train row still has a nearest neighbour
    got = ctx.match_descriptors(Qb, Tb, float("inf"), True)
    assert_matches_equal(got, cv2_oracle.mutual_nn(Qb, Tb))


def test_edge_cases(ctx):
    Q, T = synth.surf_like(2, [50, 1], seed=1)
    assert len(ctx.match_descriptors(Q, T, 0.8, False)) == 0          # train rows < 2 => no matches (F7)
    assert len(ctx.match_descriptors(Q[:0], Q, 0.8, True)) == 0       # empty query
    assert len(ctx.match_descriptors(Q, Q[:0], 0.8, True)) == 0       # empty train
    Qb = synth.orb_like(1, 64, seed=2)[0]
    assert len(ctx.match_descriptors(Qb, Qb[:1], 0.8, False)) == 0
    m = ctx.match_descriptors(Qb, Qb, 0.8, True)                       # identical frames: every row matches itself at d = 0
    assert (m["queryIdx"] == m["trainIdx"]).all() and (m["distance"] == 0).all()
    assert_matches_equal(m, oracle.match(Qb, Qb, 0.8, True))


def test_error_behaviour(ctx):
    import easysfm_b200 as esfm
    b = ctx.bank(esfm.KIND_F32X64, 2)
    with pytest.raises(esfm.EsfmError):
        b.set_frame(0, np.zeros((4, 32), np.float32))                  # wrong width
    with pytest.raises(esfm.EsfmError):
        b.set_frame(5, np.zeros((4, 64), np.float32))                  # frame id out of range
    with pytest.raises(esfm.EsfmError):
        b.match_all_pairs(0.8)                                         # not committed
    b.set_frame(0, np.zeros((4, 64), np.float32))
    with pytest.raises(esfm.EsfmError):
        b.commit()                                                     # frame 1 never set


def test_set_frame_pinned_equals_set_frame(ctx):
    """esfm_bank_set_frame_pinned (no host-side copy, commit reads the caller's page-locked memory) gives the same bank as
    esfm_bank_set_frame, for both kinds and with a mix of the two calls; pageable memory is refused."""
    import torch
    import easysfm_b200 as esfm
    for frames in (synth.orb_like(3, [300, 129, 512], seed=11), synth.surf_like(3, [300, 129, 512], seed=11)):
        kind = esfm.KIND_B256 if frames[0].dtype == np.uint8 else esfm.KIND_F32X64
        ref_bank = ctx.bank_from_frames(frames)
        ref = ref_bank.match_all_pairs(0.8, True)
        pinned = [torch.from_numpy(f.copy()).pin_memory() for f in frames]
        b = ctx.bank(kind, 3)
        b.set_frame_pinned(0, pinned[0].numpy())
        b.set_frame(1, frames[1])                      # mixing the two calls is allowed
        b.set_frame_pinned(2, pinned[2].numpy())
        b.commit()
        got = b.match_all_pairs(0.8, True)
        for k in range(ref.n_pairs):
            i, j, m = ref.pair_at(k)
            assert_matches_equal(got.pair(i, j), m)
        b2 = ctx.bank(kind, 1)
        with pytest.raises(esfm.EsfmError):
            b2.set_frame_pinned(0, frames[0])          # pageable numpy memory


@pytest.mark.parametrize("nq,nt,levels", [(300, 700, 3), (129, 1025, 2), (515, 260, 5)])
def test_l2_exact_ties_on_quantised_descriptors(ctx, nq, nt, levels):
    """Descriptors quantised to a few levels (k/8): every distance is exact in fp32 in every implementation (OpenCV, the C
    oracle, the FP32-FMA sweep, the 3xTF32 tensor-core sweep), so the data is full of EXACT ties at rank 1 and 2 and the
    lowest-index rule must hold index for index -- forward knn-2, ratio test and mutual cross-check."""
    rng = np.random.default_rng(nq * 31 + nt)
    Q = (rng.integers(-levels, levels + 1, (nq, 64)) / 8.0).astype(np.float32)
    T = (rng.integers(-levels, levels + 1, (nt, 64)) / 8.0).astype(np.float32)
    T[7] = T[3]; T[nt - 1] = T[3]; Q[11] = T[3]; Q[12] = T[3]
    bank = ctx.bank_from_frames([Q, T])
    idx, dist = bank.knn2_pair(0, 1)
    ridx, rdist = oracle.knn2(Q, T)
    np.testing.assert_array_equal(idx, ridx)
    np.testing.assert_array_equal(dist, rdist)
    cidx, cdist = cv2_oracle.knn2(Q, T)
    np.testing.assert_array_equal(idx, cidx)
    np.testing.assert_array_equal(dist, cdist)
    for ratio in (0.8, 1.0, float("inf")):       # +inf = no ratio test: the duplicates (d2 = 0) must survive it
        for cc in (False, True):
            assert_matches_equal(ctx.match_descriptors(Q, T, ratio, cc), oracle.match(Q, T, ratio, cc))
    import cv2
    want = sorted((m.queryIdx, m.trainIdx, m.distance) for m in cv2.BFMatcher(cv2.NORM_L2, crossCheck=True).match(Q, T))
    got = ctx.match_descriptors(Q, T, float("inf"), True)
    assert [(int(m["queryIdx"]), int(m["trainIdx"]), float(m["distance"])) for m in got] == want    # feature_match.py:26-27 itself


def test_hamming_engines_are_byte_identical(ctx):
    """XOR + POPC and the FP8 +-1 tensor-core dot product give the same bytes: matches (ratio / cross-check / mutual-NN mode)
    and raw knn-2, on a ragged bank with empty, 1-row, non-multiple-of-128 frames and exact duplicates."""
    rows = [700, 0, 1, 129, 1025, 2, 512, 300]
    frames = synth.orb_like(len(rows), rows, seed=9)
    frames[4][10] = frames[4][3]; frames[6][5] = frames[4][3]; frames[6][7] = frames[4][3]
    keep = ctx.hamming_engine()
    out = {}
    try:
        for eng in ("popc", "tc", "tc16"):
            ctx.set_hamming_engine(eng)
            bank = ctx.bank_from_frames(frames)
            per = []
            for ratio, cc in ((0.8, True), (0.8, False), (float("inf"), True)):
                res = bank.match_all_pairs(ratio, cc)
                per += [res.pair_at(k)[2].tobytes() for k in range(res.n_pairs)]
            for (i, j) in ((4, 6), (6, 4), (0, 7), (3, 2), (0, 5), (2, 0)):
                idx, dist = bank.knn2_pair(i, j)
                per += [idx.tobytes(), dist.tobytes()]
            out[eng] = per
            bank.close()
    finally:
        ctx.set_hamming_engine(keep)
    assert out["popc"] == out["tc"]
    assert out["popc"] == out["tc16"]
    # the +-1 operand encoding with the generic epilogue ($ESFM_ORB_Z=0, read at esfm_init): what frames with more than 32768
    # rows fall back to (the default Z encoding packs the column index into 15 bits of the accumulator)
    import os
    import easysfm_b200 as esfm
    old = os.environ.get("ESFM_ORB_Z")
    os.environ["ESFM_ORB_Z"] = "0"
    try:
        with esfm.Context(0) as c2:
            c2.set_hamming_engine("tc")
            bank = c2.bank_from_frames(frames)
            per = []
            for ratio, cc in ((0.8, True), (0.8, False), (float("inf"), True)):
                res = bank.match_all_pairs(ratio, cc)
                per += [res.pair_at(k)[2].tobytes() for k in range(res.n_pairs)]
            for (i, j) in ((4, 6), (6, 4), (0, 7), (3, 2), (0, 5), (2, 0)):
                idx, dist = bank.knn2_pair(i, j)
                per += [idx.tobytes(), dist.tobytes()]
            bank.close()
    finally:
        if old is None:
            os.environ.pop("ESFM_ORB_Z", None)
        else:
            os.environ["ESFM_ORB_Z"] = old
    assert per == out["popc"]


# --------------------------------------------------------------------------------------------------
# cross-check under contention: many query rows compete for the same train rows
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["orb", "surf"])
def test_cross_check_contention(ctx, kind):
    """The tc16 sweeps decide the cross-check without column minima: survivors of the ratio test claim their train row, histograms
    of every row's best / second-best value tell which survivors could still be beaten, and only those go through a verification
    sweep (finalize.cu, two-phase).  Here several query rows -- also ones that FAIL the ratio test, and ones whose second neighbour
    is the contested train row -- sit at graded distances from the same train rows, and exact duplicates tie: every path must agree
    with the oracle, for all engines, with and without the ratio test."""
    rng = np.random.default_rng(17)
    nt, nq = 1100, 1500
    if kind == "orb":
        T = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
        Q = rng.integers(0, 256, (nq, 32), dtype=np.uint8)

        def flip(row, nbits):
            r = row.copy()
            for b in rng.choice(256, nbits, replace=False):
                r[b >> 3] ^= np.uint8(1 << (b & 7))
            return r
        for k in range(300):                      # 2-5 queries per contested train row, at 0..60 flipped bits
            t = int(rng.integers(0, 200))
            Q[int(rng.integers(0, nq))] = flip(T[t], int(rng.integers(0, 60)))
        for k in range(40):                       # exact duplicates of one train row in several query rows (index tie-break)
            Q[int(rng.integers(0, nq))] = T[int(rng.integers(200, 220))]
        for k in range(60):                       # a query between TWO train rows: fails the ratio test but is near both
            a, b = int(rng.integers(0, 200)), int(rng.integers(220, 260))
            T[b] = flip(T[a], 6)
            Q[int(rng.integers(0, nq))] = flip(T[a], 3)
    else:
        T = rng.standard_normal((nt, 64)).astype(np.float32)
        T /= np.linalg.norm(T, axis=1, keepdims=True)
        Q = rng.standard_normal((nq, 64)).astype(np.float32)
        Q /= np.linalg.norm(Q, axis=1, keepdims=True)

        def jitter(row, s):
            r = row + s * rng.standard_normal(64).astype(np.float32)
            return (r / np.linalg.norm(r)).astype(np.float32)
        for k in range(300):
            t = int(rng.integers(0, 200))
            Q[int(rng.integers(0, nq))] = jitter(T[t], float(rng.uniform(0.0, 0.12)))
        for k in range(40):
            Q[int(rng.integers(0, nq))] = T[int(rng.integers(200, 220))]
        for k in range(60):
            a, b = int(rng.integers(0, 200)), int(rng.integers(220, 260))
            T[b] = jitter(T[a], 0.01)
            Q[int(rng.integers(0, nq))] = jitter(T[a], 0.005)
    removed = 0
    for ratio in (0.8, 1.0, float("inf")):
        ref = oracle.match(Q, T, ratio, True)
        got = ctx.match_descriptors(Q, T, ratio, True)
        if kind == "orb":
            assert_matches_equal(got, ref)
        else:
            justify_l2(Q, T, ratio, True, got, ref)
        removed += len(oracle.match(Q, T, ratio, False)) - len(ref)
    assert removed > 300          # the cross-check really had something to remove
    # all pairs of a small bank built from the same rows (one launch, several pairs, both orientations of the contention)
    frames = [Q[:700], T[:500], Q[700:], T[500:]]
    bank = ctx.bank_from_frames(frames)
    res = bank.match_all_pairs(0.8, True)
    for k in range(res.n_pairs):
        i, j, m = res.pair_at(k)
        ref = oracle.match(frames[i], frames[j], 0.8, True)
        if kind == "orb":
            assert_matches_equal(m, ref)
        else:
            justify_l2(frames[i], frames[j], 0.8, True, m, ref)
    bank.close()

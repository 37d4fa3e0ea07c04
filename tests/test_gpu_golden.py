"""CUDA path (through the C ABI) against the committed golden fixtures generated from the reference's cv2 calls."""
import numpy as np
import pytest

from golden_util import as_matches, frames_of, load
from util import assert_matches_equal, justify_l2

pytestmark = pytest.mark.gpu


def _check(Q, T, ratio, cc, got, ref):
    if Q.dtype == np.uint8:
        assert_matches_equal(got, ref)
    elif len(Q) and len(T):
        n = justify_l2(Q, T, ratio, bool(cc), got, ref)
        assert n <= max(1, len(ref) // 500)
    else:
        assert len(got) == len(ref) == 0


def test_kat_cases_gpu(ctx):
    z = load("kat_cases.npz")
    for name in z["names"]:
        Q, T = z[f"{name}_Q"], z[f"{name}_T"]
        for r in (50, 80):
            for cc in (0, 1):
                got = ctx.match_descriptors(Q, T, r / 100.0, bool(cc))
                _check(Q, T, r / 100.0, cc, got, as_matches(z[f"{name}_r{r}_c{cc}_idx"], z[f"{name}_r{r}_c{cc}_dist"]))
        bank = ctx.bank_from_frames([Q, T])
        idx, dist = bank.knn2_pair(0, 1)
        np.testing.assert_array_equal(idx, z[f"{name}_knn_idx"])
        if Q.dtype == np.uint8:
            np.testing.assert_array_equal(dist, z[f"{name}_knn_dist"])
        else:
            np.testing.assert_allclose(dist, z[f"{name}_knn_dist"], rtol=1e-5)
        m = ctx.match_descriptors(Q, T, float("inf"), True)
        assert_matches_equal(m, as_matches(z[f"{name}_mutual_idx"], z[f"{name}_mutual_dist"]), exact_distance=(Q.dtype == np.uint8))


@pytest.mark.parametrize("kind", ["surf", "orb"])
def test_synth_all_pairs_gpu(ctx, kind):
    z = load(f"synth_{kind}.npz")
    fr = frames_of(z)
    bank = ctx.bank_from_frames(fr)
    for r in (50, 80):
        for cc in (0, 1):
            res = bank.match_all_pairs(r / 100.0, bool(cc))
            for k in range(res.n_pairs):
                i, j, m = res.pair_at(k)
                _check(fr[i], fr[j], r / 100.0, cc, m, as_matches(z[f"m_r{r}_c{cc}_{i}_{j}_idx"], z[f"m_r{r}_c{cc}_{i}_{j}_dist"]))


def test_fountain_orb_all_55_pairs_gpu(ctx):
    """BASELINE configs[0] stand-in: ORB(8000) on the 11 bundled fountain images, all 55 pairs, bit-exact vs cv2,
    cross_check 0 (== the reference's matchFeaturesORB, feature_matching.cpp:71-92) and 1."""
    z = load("fountain_orb.npz")
    fr = frames_of(z)
    bank = ctx.bank_from_frames(fr)
    for cc in (0, 1):
        res = bank.match_all_pairs(0.8, bool(cc))
        assert res.n_pairs == 55
        total = 0
        for k in range(res.n_pairs):
            i, j, m = res.pair_at(k)
            assert_matches_equal(m, as_matches(z[f"m_r80_c{cc}_{i}_{j}_idx"], z[f"m_r80_c{cc}_{i}_{j}_dist"]))
            total += len(m)
        assert total > 1000


def test_reference_interface_mirror_gpu(ctx):
    """FeatureMatching.matchFeaturesORB/SURF keep the reference's call shape: append to the caller's list, return True."""
    import easysfm_b200 as esfm
    z = load("fountain_orb.npz")
    fr = frames_of(z)[:3]
    frames = [esfm.Frame(i, d) for i, d in enumerate(fr)]
    fm = esfm.FeatureMatching(ctx=ctx)
    out = []
    assert fm.matchFeaturesORB(frames[1], frames[0], out) is True            # default ratio 0.8 (feature_matching.h:18)
    ref = as_matches(z["m_r80_c0_1_0_idx"], z["m_r80_c0_1_0_dist"])
    assert_matches_equal(np.array(out, dtype=ref.dtype), ref)
    n0 = len(out)
    fm.matchFeaturesORB(frames[2], frames[1], out)                            # appends, never clears (:90)
    assert len(out) == n0 + len(z["m_r80_c0_2_1_idx"])
    fm.prepare(frames, 0.8)                                                   # all-pairs pre-pass, then lookups
    out2 = []
    fm.matchFeaturesORB(frames[1], frames[0], out2)
    assert_matches_equal(np.array(out2, dtype=ref.dtype), ref)
    zs = load("synth_surf.npz")
    fs = frames_of(zs)
    outs = []
    assert fm.matchFeaturesSURF(esfm.Frame(1, fs[1]), esfm.Frame(0, fs[0]), outs) is True   # default ratio 0.5 (:21)
    refs = as_matches(zs["m_r50_c0_1_0_idx"], zs["m_r50_c0_1_0_dist"])
    justify_l2(fs[1], fs[0], 0.5, False, np.array(outs, dtype=refs.dtype), refs)
    m = esfm.pairwise_match_descriptors(fs[1], fs[0], "ratio_test", ctx=ctx)               # nn_ratio 0.7 (feature_match.py:5)
    assert len(m) > 0

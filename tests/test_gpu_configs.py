"""Config-level parity on the GPU (BASELINE.json configs[1], configs[2]) and size-independent properties at full frame size."""
import numpy as np
import pytest

import oracle
from oracle import cv2_oracle
from easysfm_b200 import scheduler, synth
from util import assert_matches_equal, justify_l2

pytestmark = pytest.mark.gpu


def test_config2_surf_25x2k_all_300_pairs(ctx):
    """configs[1]: synthetic SURF 64-d fp32, 25 images x 2k features, all-pairs 2-NN + ratio 0.8 + cross-check vs the oracle
    (all 300 pairs), plus ratio 0.5 / cross_check 0 on a sample, plus cv2 itself on a sample."""
    fr = synth.surf_like(25, 2000, seed=2)
    bank = ctx.bank_from_frames(fr)
    res = bank.match_all_pairs(0.8, True)
    assert res.n_pairs == 300
    ndiff = 0
    for k in range(res.n_pairs):
        i, j, m = res.pair_at(k)
        ref = oracle.match(fr[i], fr[j], 0.8, True)
        if len(m) == len(ref) and (m["trainIdx"] == ref["trainIdx"]).all() and (m["queryIdx"] == ref["queryIdx"]).all():
            np.testing.assert_array_equal(m["distance"], ref["distance"])     # same candidates => bit-identical distances
        else:
            ndiff += justify_l2(fr[i], fr[j], 0.8, True, m, ref)
    assert ndiff <= 3
    rng = np.random.default_rng(0)
    pairs = scheduler.all_pairs(25)
    for ratio, cc in ((0.5, False), (0.5, True), (0.8, False)):
        sel = pairs[rng.choice(len(pairs), 12, replace=False)]
        r2 = bank.match_pairs(sel, ratio, cc)
        for k in range(r2.n_pairs):
            i, j, m = r2.pair_at(k)
            justify_l2(fr[i], fr[j], ratio, cc, m, oracle.match(fr[i], fr[j], ratio, cc))
    for (i, j) in [(3, 1), (24, 0), (17, 16)]:
        justify_l2(fr[i], fr[j], 0.8, True, res.pair(i, j), cv2_oracle.match(fr[i], fr[j], 0.8, True))


def test_config3_orb_8k_bit_exact_sample(ctx):
    """configs[2] shape (ORB 256-bit, 8k features per image): 40 frames here (780 pairs on the GPU), every pair against the
    C oracle on a seeded sample of 24 pairs and against cv2 on 6 -- indices AND distances bit-exact."""
    fr = synth.orb_like(40, 8000, seed=3)
    bank = ctx.bank_from_frames(fr)
    res = bank.match_all_pairs(0.8, True)
    assert res.n_pairs == 780
    rng = np.random.default_rng(1)
    pairs = scheduler.all_pairs(40)
    for n, k in enumerate(rng.choice(len(pairs), 24, replace=False)):
        i, j = pairs[k]
        m = res.pair(int(i), int(j))
        assert_matches_equal(m, oracle.match(fr[i], fr[j], 0.8, True))
        if n < 6:
            assert_matches_equal(m, cv2_oracle.match(fr[i], fr[j], 0.8, True))
    r0 = bank.match_pairs(pairs[:8], 0.8, False)      # cross_check = 0 == the reference's matchFeaturesORB
    for k in range(8):
        i, j, m = r0.pair_at(k)
        assert_matches_equal(m, cv2_oracle.match(fr[i], fr[j], 0.8, False))


@pytest.mark.parametrize("kind", ["surf", "orb"])
def test_mutual_nn_is_symmetric_at_full_size(ctx, kind):
    """Size-independent property at the bench frame size: with ratio = +inf and cross_check, (q, t) is a match of (A -> B)
    iff (t, q) is a match of (B -> A), with equal distances (mutual nearest neighbours, lowest index on ties both ways)."""
    n = 8000 if kind == "surf" else 4000
    A, B = (synth.surf_like if kind == "surf" else synth.orb_like)(2, [n, n - 37], seed=5)
    bank = ctx.bank_from_frames([A, B])
    r = bank.match_pairs([[0, 1], [1, 0]], float("inf"), True)
    ab, ba = r.pair_at(0)[2], r.pair_at(1)[2]
    assert len(ab) == len(ba) > 0
    fwd = {(int(m["queryIdx"]), int(m["trainIdx"])): float(m["distance"]) for m in ab}
    bwd = {(int(m["trainIdx"]), int(m["queryIdx"])): float(m["distance"]) for m in ba}
    assert fwd == bwd
    assert (np.diff(ab["queryIdx"]) > 0).all() and (np.diff(ba["queryIdx"]) > 0).all()   # ascending queryIdx, no duplicates


def test_results_are_deterministic_and_chunk_invariant(ctx):
    """The same pairs in one batch, in two batches, and one by one give byte-identical matches (atomics only pick winners by key)."""
    fr = synth.surf_like(6, [1500, 900, 1300, 700, 1100, 1000], seed=8)
    bank = ctx.bank_from_frames(fr)
    pairs = scheduler.all_pairs(6)
    whole = bank.match_pairs(pairs, 0.8, True)
    a = bank.match_pairs(pairs[:7], 0.8, True)
    b = bank.match_pairs(pairs[7:], 0.8, True)
    for k in range(len(pairs)):
        ref = whole.pair_at(k)[2]
        part = a.pair_at(k)[2] if k < 7 else b.pair_at(k - 7)[2]
        assert part.tobytes() == ref.tobytes()
        i, j = pairs[k]
        assert bank.match_pair(int(i), int(j), 0.8, True).tobytes() == ref.tobytes()     # split over many CTAs (units_per_pair > 1)

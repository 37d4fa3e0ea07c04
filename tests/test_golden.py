"""The oracle (C restatement) against the committed golden fixtures that tests/golden/make_golden.py produced by
running the reference's cv2.BFMatcher calls.  CPU only; no cv2, no /root/reference needed at test time."""
import numpy as np
import pytest

import oracle
from golden_util import as_matches, frames_of, load
from util import assert_matches_equal, justify_l2


def _check(Q, T, ratio, cc, ref):
    got = oracle.match(Q, T, ratio, bool(cc))
    if Q.dtype == np.uint8:
        assert_matches_equal(got, ref)           # Hamming: bit-exact
    elif len(Q) and len(T):
        justify_l2(Q, T, ratio, bool(cc), got, ref)
    else:
        assert len(got) == len(ref) == 0


def test_kat_cases():
    z = load("kat_cases.npz")
    for name in z["names"]:
        Q, T = z[f"{name}_Q"], z[f"{name}_T"]
        for r in (50, 80):
            for cc in (0, 1):
                _check(Q, T, r / 100.0, cc, as_matches(z[f"{name}_r{r}_c{cc}_idx"], z[f"{name}_r{r}_c{cc}_dist"]))
        idx, dist = oracle.knn2(Q, T)
        np.testing.assert_array_equal(idx, z[f"{name}_knn_idx"])
        if Q.dtype == np.uint8:
            np.testing.assert_array_equal(dist, z[f"{name}_knn_dist"])
        else:
            np.testing.assert_allclose(dist, z[f"{name}_knn_dist"], rtol=1e-5)
        m = oracle.mutual_nn(Q, T)
        ref = as_matches(z[f"{name}_mutual_idx"], z[f"{name}_mutual_dist"])
        assert_matches_equal(m, ref, exact_distance=(Q.dtype == np.uint8))


@pytest.mark.parametrize("kind", ["surf", "orb"])
def test_synth_all_pairs(kind):
    z = load(f"synth_{kind}.npz")
    fr = frames_of(z)
    for r in (50, 80):
        for cc in (0, 1):
            for i in range(len(fr)):
                for j in range(i):
                    _check(fr[i], fr[j], r / 100.0, cc, as_matches(z[f"m_r{r}_c{cc}_{i}_{j}_idx"], z[f"m_r{r}_c{cc}_{i}_{j}_dist"]))


def test_fountain_orb_sample():
    """BASELINE configs[0] stand-in (ORB on the 11 bundled fountain images): a sample of the 55 pairs on CPU;
    all 55 are checked on the GPU in test_gpu_golden.py."""
    z = load("fountain_orb.npz")
    fr = frames_of(z)
    assert len(fr) == 11
    for (i, j) in [(1, 0), (5, 2), (10, 9)]:
        for cc in (0, 1):
            _check(fr[i], fr[j], 0.8, cc, as_matches(z[f"m_r80_c{cc}_{i}_{j}_idx"], z[f"m_r80_c{cc}_{i}_{j}_dist"]))

"""oracle/orb_oracle.py (ORB extraction) against cv2's own output -- the committed golden vectors
(tests/golden/orb_extract.npz, made by tests/golden/make_golden_orb.py with cv2 4.13.0) and, when cv2 is importable, cv2 itself on further
seeded images.  Bit for bit: key-point coordinates, size, angle, Harris response, octave, ORDER, and all 256 descriptor bits."""
import os

import numpy as np
import pytest

from oracle import orb_oracle as oo
from orb_util import CASES, image

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "orb_extract.npz"))


def assert_same(kp, desc, ref_kp, ref_desc):
    assert len(kp) == len(ref_kp)
    for f in ref_kp.dtype.names:
        assert np.array_equal(kp[f], ref_kp[f]), f
    assert np.array_equal(desc, ref_desc)


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_equals_cv2_golden(name):
    seed, h, w, shapes, bgr, nf = CASES[name]
    kp, desc = oo.detect_and_compute(image(seed, h, w, shapes, bgr), nf)
    assert_same(kp, desc, GOLD[name + "_kp"], GOLD[name + "_desc"])


def test_oracle_equals_cv2_on_a_photograph():
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "orb_fountain.npz"))
    kp, desc = oo.detect_and_compute(g["image"], int(g["max_features"]))
    assert_same(kp, desc, g["kp"], g["desc"])


def test_pieces_known_answers():
    # BGR -> gray weights sum to 1 << 15; the measured sampling pattern stays inside the 31 x 31 patch and has no degenerate test
    g = oo.bgr_to_gray(np.full((2, 2, 3), 200, np.uint8))
    assert (g == 200).all()
    p = oo.pattern()
    assert p.shape == (256, 4) and np.abs(p).max() <= 15 and not ((p[:, 0] == p[:, 2]) & (p[:, 1] == p[:, 3])).any()
    assert list(p[0]) == [8, -3, 9, 5]
    assert oo.features_per_level(5000) == [1086, 905, 754, 628, 524, 436, 364, 303]
    assert oo.UMAX == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    # a flat image has no corners; resizing a flat image and blurring it keep it flat
    flat = np.full((100, 120), 77, np.uint8)
    assert oo.fast_scores(flat).max() == 0
    assert (oo.resize_linear_exact(flat, 100, 83) == 77).all() and (oo.gaussian_blur_7x7(flat) == 77).all()
    # fastAtan2: the four axes and the documented 0.3 degree accuracy
    assert oo.fast_atan2(0, 1) == 0 and abs(oo.fast_atan2(1, 0) - 90) < 1e-4 and abs(oo.fast_atan2(-1, 0) - 270) < 1e-4
    for a in np.linspace(0.1, 359.9, 97):
        assert abs(oo.fast_atan2(np.sin(np.radians(a)), np.cos(np.radians(a))) - a) < 0.3


def test_retain_best_keeps_boundary_ties():
    r = np.array([5, 1, 3, 3, 9, 3, 0, 3], np.float32)
    keep = oo.retain_best(r, 3)
    assert sorted(keep.tolist()) == [0, 2, 3, 4, 5, 7]            # 9, 5 and every 3
    assert oo.retain_best(r, 8).tolist() == list(range(8))        # nothing to drop: order untouched
    assert len(oo.retain_best(r, 0)) == 0


def test_oracle_equals_cv2_live():
    cv2 = pytest.importorskip("cv2")
    for seed, (h, w), nf in [(101, (240, 320), 1500), (102, (333, 257), 800), (103, (480, 640), 3000)]:
        img = image(seed, h, w, 40)
        k = cv2.ORB_create(nf).detect(img, None)
        k, d = cv2.ORB_create(nf).compute(img, k)
        ref = np.zeros(len(k), oo.KP_DTYPE)
        for i, p in enumerate(k):
            ref[i] = (p.pt[0], p.pt[1], p.size, p.angle, p.response, p.octave)
        kp, desc = oo.detect_and_compute(img, nf)
        assert_same(kp, desc, ref, d)

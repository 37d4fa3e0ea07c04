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


# ---- stage-by-stage pins against the cv2 functions ORB is built from (live; skipped where cv2 is absent) --------------------------------
def test_stages_equal_cv2_live():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(7)
    # color_rgb BGR2GRAY
    bgr = rng.integers(0, 256, (97, 131, 3), dtype=np.uint8)
    assert np.array_equal(oo.bgr_to_gray(bgr), cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY))
    # resize.cpp INTER_LINEAR_EXACT, including the 1-pixel-smaller and strongly reduced cases
    img = rng.integers(0, 256, (480, 640), dtype=np.uint8)
    for dw, dh in [(533, 400), (444, 333), (639, 479), (100, 77), (179, 134)]:
        assert np.array_equal(oo.resize_linear_exact(img, dw, dh), cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR_EXACT)), (dw, dh)
    # fast.cpp + fast_score.cpp: corners, order and scores of cv2's FAST-9/16 with non-maximum suppression
    tex = image(303, 200, 260, 30)
    ref = cv2.FastFeatureDetector_create(oo.FAST_THRESHOLD, True).detect(tex, None)
    ys, xs = oo.fast_nms(oo.fast_scores(tex))
    assert len(ref) == len(xs) > 100
    assert [(int(k.pt[0]), int(k.pt[1])) for k in ref] == list(zip(xs.tolist(), ys.tolist()))          # raster order
    assert [k.response for k in ref] == oo.fast_scores(tex)[ys, xs].astype(np.float32).tolist()
    # without suppression: every corner pixel
    allc = cv2.FastFeatureDetector_create(oo.FAST_THRESHOLD, False).detect(tex, None)
    yy, xx = np.nonzero(oo.fast_scores(tex))
    assert sorted((int(k.pt[0]), int(k.pt[1])) for k in allc) == sorted(zip(xx.tolist(), yy.tolist()))
    # core fastAtan2
    for _ in range(2000):
        y, x = (float(v) for v in rng.integers(-70000, 70000, 2))
        assert oo.fast_atan2(y, x) == np.float32(cv2.fastAtan2(y, x)), (y, x)
    # getGaussianKernel(7, 2, CV_32F)
    assert np.array_equal(oo.gaussian_kernel_7(), cv2.getGaussianKernel(7, 2, cv2.CV_32F).ravel())


def test_compute_on_given_keypoints_equals_cv2_live():
    """ORB::compute alone (the reference's second call, feature_matching.cpp:22): hand-placed key points with arbitrary angles and octaves."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(9)
    img = image(404, 300, 400, 40)
    levels, scales = oo.build_pyramid(img)
    kps, arr = [], np.zeros(300, oo.KP_DTYPE)
    for i in range(300):
        l = int(rng.integers(0, 5))
        h, w = levels[l].shape
        x = float(np.float32(rng.integers(40, w - 40)) * scales[l])
        y = float(np.float32(rng.integers(40, h - 40)) * scales[l])
        ang = float(np.float32(rng.random() * 360.0))
        kps.append(cv2.KeyPoint(x, y, 31.0 * float(scales[l]), ang, 1.0, l, -1))
        arr[i] = (x, y, 31.0 * float(scales[l]), ang, 1.0, l)
    out_kps, desc = cv2.ORB_create(500).compute(img, kps)
    assert len(out_kps) == 300
    # key points that are not sorted by level come back grouped by level (a stable regrouping inside ORB); detect()'s output, which is what
    # the reference passes, already is
    order = np.argsort(arr["octave"], kind="stable")
    assert [(k.pt[0], k.pt[1], k.octave) for k in out_kps] == [(float(arr["x"][i]), float(arr["y"][i]), int(arr["octave"][i])) for i in order]
    assert np.array_equal(oo.compute_descriptors(levels, scales, arr[order]), desc)

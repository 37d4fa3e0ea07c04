"""GPU: esfm_orb_extract / esfm_bank_set_frame_from_image through the C ABI against cv2's golden output
(tests/golden/orb_extract.npz) and the numpy oracle -- bit for bit: key-point order, coordinates, size, angle, response, octave, descriptors;
the intermediate pyramid and blurred levels; and the bank filled on the device against a bank filled from the host."""
import os

import numpy as np
import pytest

from easysfm_b200 import capi
from easysfm_b200.feature_matching import FeatureMatching, Frame
from oracle import orb_oracle as oo
from orb_util import CASES, image

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "orb_extract.npz"))


@pytest.fixture(scope="module")
def octx():
    return capi.Context(0)


def assert_same(kp, desc, ref_kp, ref_desc):
    assert len(kp) == len(ref_kp)
    for f in ref_kp.dtype.names:
        assert np.array_equal(kp[f], ref_kp[f]), f
    assert np.array_equal(desc, ref_desc)


@pytest.mark.parametrize("name", list(CASES))
def test_extract_equals_cv2_golden(octx, name):
    seed, h, w, shapes, bgr, nf = CASES[name]
    kp, desc = octx.orb_extract(image(seed, h, w, shapes, bgr), nf)
    assert_same(kp, desc, GOLD[name + "_kp"], GOLD[name + "_desc"])


def test_extract_equals_cv2_on_a_photograph(octx):
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "orb_fountain.npz"))
    kp, desc = octx.orb_extract(g["image"], int(g["max_features"]))
    assert_same(kp, desc, g["kp"], g["desc"])


def test_levels_equal_oracle(octx):
    img = image(31, 301, 417, 30, True)
    kp, desc = octx.orb_extract(img, 1500)
    gray = oo.bgr_to_gray(img)
    levels, scales = oo.build_pyramid(gray)
    used = set(int(o) for o in kp["octave"])
    for l, ref in enumerate(levels):
        got = octx.orb_debug_level(l)
        assert got.shape == ref.shape and np.array_equal(got, ref), l
        if l in used:
            assert np.array_equal(octx.orb_debug_level(l, True), oo.gaussian_blur_7x7(ref)), l
    rkp, rdesc = oo.detect_and_compute(img, 1500)
    assert_same(kp, desc, rkp, rdesc)


@pytest.mark.parametrize("seed,h,w,nf", [(41, 233, 377, 700), (42, 480, 640, 20000), (43, 64, 64, 50), (44, 90, 1200, 400), (45, 1080, 1920, 5000)])
def test_extract_equals_oracle(octx, seed, h, w, nf):
    img = image(seed, h, w, 50)
    kp, desc = octx.orb_extract(img, nf)
    rkp, rdesc = oo.detect_and_compute(img, nf)
    assert_same(kp, desc, rkp, rdesc)


def test_flat_and_strided_images(octx):
    kp, desc = octx.orb_extract(np.full((200, 300), 90, np.uint8), 500)
    assert len(kp) == 0 and desc.shape == (0, 32)
    big = image(51, 260, 400, 30)
    view = big[10:250, 20:380]                      # row stride > width
    kp, desc = octx.orb_extract(view, 600)
    rkp, rdesc = oo.detect_and_compute(np.ascontiguousarray(view), 600)
    assert_same(kp, desc, rkp, rdesc)
    kp0, desc0 = octx.orb_extract(big, 0)
    assert len(kp0) == 0


def test_bank_from_images_equals_bank_from_host(octx):
    imgs = [image(60 + k, 240, 320, 30) for k in range(4)]
    fm = FeatureMatching(octx, cross_check=True)
    frames = [Frame(k, rgb_image=im) for k, im in enumerate(imgs)]
    res_dev = fm.prepare_from_images(frames, 0.8, max_num=800)
    host = [Frame(k, rgb_image=im) for k, im in enumerate(imgs)]
    for f in host:
        assert fm.detectFeaturesORB(f, 800)
    for f, g in zip(frames, host):
        assert np.array_equal(f.keypoints, g.keypoints)
    res_host = octx.bank_from_frames([f.descriptors for f in host]).match_all_pairs(0.8, True)
    assert res_dev.n_pairs == res_host.n_pairs == 6
    for i in range(4):
        for j in range(i):
            assert np.array_equal(res_dev.pair(i, j), res_host.pair(i, j))
            ref = __import__("oracle").match(host[i].descriptors, host[j].descriptors, 0.8, True)
            assert np.array_equal(res_dev.pair(i, j), ref)


def test_errors(octx):
    img = image(70, 120, 160, 10)
    with pytest.raises(TypeError):
        octx.orb_extract(img.astype(np.float32), 100)
    lib = capi.load_library()
    n = capi.c_int(0)
    kps = np.zeros(4, capi.KEYPOINT_DTYPE)
    rc = lib.esfm_orb_extract(octx._h, img.ctypes.data, 120, 160, 2, 320, 100, kps.ctypes.data, None, 4, capi.ctypes.byref(n))
    assert rc == 1 and b"channels" in lib.esfm_last_error()
    rc = lib.esfm_orb_extract(octx._h, img.ctypes.data, 120, 160, 1, 160, 300, kps.ctypes.data, None, 4, capi.ctypes.byref(n))
    assert rc == capi.ERR_CAPACITY and n.value > 4
    fb = capi.Bank(octx, capi.KIND_F32X64, 2)
    with pytest.raises(capi.EsfmError):
        fb.set_frame_from_image(0, img, 100)

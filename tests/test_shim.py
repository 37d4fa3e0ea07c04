"""The C++ shim that keeps p3dv::FeatureMatching::matchFeaturesORB/SURF's signatures (easysfm_b200/shim) compiles against a
stand-in for the EasySFM/OpenCV headers and links against the C-ABI library (CPU); on the GPU it is driven like
cpp_code/test/sfm.cpp:140-161 and must reproduce the oracle."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "easysfm_b200", "shim", "feature_matching_gpu.cpp")
LIBDIR = os.path.join(ROOT, "easysfm_b200", "lib")


def build_shim(tmp):
    exe = os.path.join(tmp, "shim_main")
    cmd = ["g++", "-std=c++14", "-O1", "-Wall", "-I", os.path.join(ROOT, "tests", "shim"), "-I", os.path.join(ROOT, "include"),
           SHIM, os.path.join(ROOT, "tests", "shim", "shim_main.cpp"), "-L", LIBDIR, "-lesfm_match", f"-Wl,-rpath,{LIBDIR}", "-o", exe]
    subprocess.check_call(cmd)
    return exe


def test_shim_compiles_and_links(tmp_path):
    if not os.path.exists(os.path.join(LIBDIR, "libesfm_match.so")):
        sys.path.insert(0, ROOT)
        import __graft_entry__
        __graft_entry__.build()
    exe = build_shim(str(tmp_path))
    assert os.path.exists(exe)


@pytest.mark.gpu
@pytest.mark.parametrize("feature,prepare", [("O", 0), ("O", 1), ("S", 0), ("S", 1)])
def test_shim_matches_oracle(tmp_path, feature, prepare):
    import oracle
    from easysfm_b200 import synth
    from util import justify_l2
    exe = build_shim(str(tmp_path))
    rows = [300, 0, 257, 190]
    frames = (synth.orb_like if feature == "O" else synth.surf_like)(len(rows), rows, seed=12)
    blob = os.path.join(str(tmp_path), "desc.bin")
    with open(blob, "wb") as f:
        for fr in frames:
            f.write(np.ascontiguousarray(fr).tobytes())
    out = os.path.join(str(tmp_path), "out.bin")
    subprocess.check_call([exe, feature, str(len(rows))] + [str(r) for r in rows] + [blob, str(prepare), out], stdout=subprocess.DEVNULL)
    raw = open(out, "rb").read()
    pos = 0
    ratio = 0.8 if feature == "O" else 0.5   # the header defaults the reference's call sites rely on (feature_matching.h:17-21)
    for i in range(len(rows)):
        for j in range(i):
            cnt = int(np.frombuffer(raw, np.int32, 1, pos)[0]); pos += 4
            m = np.frombuffer(raw, oracle.DMATCH_DTYPE, cnt, pos); pos += 16 * cnt
            ref = oracle.match(frames[i], frames[j], ratio, False)
            if feature == "O":
                assert len(m) == len(ref) and (m["trainIdx"] == ref["trainIdx"]).all() and (m["distance"] == ref["distance"]).all()
                assert (m["imgIdx"] == 0).all()
            elif len(frames[i]) and len(frames[j]):
                justify_l2(frames[i], frames[j], ratio, False, m, ref)
            else:
                assert cnt == 0
    assert pos == len(raw)


# --------------------------------------------------------------------------------------------------
# the two-view shim: p3dv::MotionEstimator::estimate2D2D_E5P_RANSAC / getDepthFast
# --------------------------------------------------------------------------------------------------
MOTION_SHIM = os.path.join(ROOT, "easysfm_b200", "shim", "estimate_motion_gpu.cpp")


def build_motion_shim(tmp):
    exe = os.path.join(tmp, "motion_main")
    cmd = ["g++", "-std=c++14", "-O1", "-Wall", "-I", os.path.join(ROOT, "tests", "shim"), "-I", os.path.join(ROOT, "include"),
           MOTION_SHIM, os.path.join(ROOT, "tests", "shim", "motion_main.cpp"), "-L", LIBDIR, "-lesfm_match", f"-Wl,-rpath,{LIBDIR}", "-o", exe]
    subprocess.check_call(cmd)
    return exe


def test_motion_shim_compiles_and_links(tmp_path):
    assert os.path.exists(build_motion_shim(str(tmp_path)))


@pytest.mark.gpu
def test_motion_shim_matches_oracle(tmp_path):
    """The C++ drop-in keeps the reference's signatures (estimate_motion.h:17-20, :26-27); driven like sfm.cpp:165-166 it must give the
    oracle's inlier matches, T = [R t; 0 1] (float) and getDepthFast's mean depth (every 20th inlier)."""
    import oracle
    from oracle import two_view_oracle as tvo
    from two_view_util import scene
    exe = build_motion_shim(str(tmp_path))
    rng = np.random.default_rng(8)
    cases = []
    inp = os.path.join(str(tmp_path), "in.bin")
    with open(inp, "wb") as f:
        specs = [(400, 0.4, 0.3), (250, 0.8, 0.1), (4, 0.0, 0.0)]
        K = scene(5, 0, 0, 0)[0]
        f.write(np.int32(len(specs)).tobytes())
        f.write(np.ascontiguousarray(K, np.float64).tobytes())
        for k, (n, noise, outl) in enumerate(specs):
            _, x1, x2, R, t = scene(n, noise, outl, 40 + k)
            p1, p2 = rng.permutation(n), rng.permutation(n)
            kp1, kp2 = np.zeros((n, 2), np.float32), np.zeros((n, 2), np.float32)
            kp1[p1], kp2[p2] = x1, x2
            m = np.zeros(n, oracle.DMATCH_DTYPE)
            m["queryIdx"], m["trainIdx"], m["distance"] = p1, p2, rng.random(n).astype(np.float32)
            f.write(np.array([n, n, n], np.int32).tobytes())
            f.write(kp1.tobytes()); f.write(kp2.tobytes()); f.write(m.tobytes())
            cases.append((x1, x2, m, K))
    out = os.path.join(str(tmp_path), "out.bin")
    subprocess.check_call([exe, inp, out], stdout=subprocess.DEVNULL)
    raw = open(out, "rb").read()
    pos = 0
    for k, (x1, x2, m, K) in enumerate(cases):
        ok, ni = np.frombuffer(raw, np.int32, 2, pos); pos += 8
        T = np.frombuffer(raw, np.float32, 16, pos).reshape(4, 4); pos += 64
        depth = float(np.frombuffer(raw, np.float64, 1, pos)[0]); pos += 8
        inl = np.frombuffer(raw, oracle.DMATCH_DTYPE, int(ni), pos); pos += 16 * int(ni)
        pair_key = ((10 + 2 * k) << 32) | (11 + 2 * k)                       # the shim's sampler key: the two frame ids
        ref = tvo.estimate_two_view(x1, x2, K, 0.99, 1.0, 1000, seed=0, pair=pair_key)
        assert ok == 1
        if ref["E"] is None:
            assert ni == 0
            continue
        want = m[ref["mask"].astype(bool)]
        assert abs(int(ni) - len(want)) <= 1
        if int(ni) == len(want):
            assert inl.tobytes() == want.tobytes()
            np.testing.assert_allclose(T[:3, :3], ref["R"], atol=1e-5)
            np.testing.assert_allclose(T[:3, 3], ref["t"], atol=1e-5)
            assert T[3].tolist() == [0.0, 0.0, 0.0, 1.0]
            # getDepthFast: every 20th INLIER with the float T the caller holds
            d_ref = tvo.mean_depth(T[:3, :3].astype(np.float64), T[:3, 3].astype(np.float64), x1, x2, K, ref["mask"], 20)
            np.testing.assert_allclose(depth, d_ref, rtol=1e-6)
    assert pos == len(raw)


# --------------------------------------------------------------------------------------------------
# detectFeaturesORB through the shim, against cv2's golden output
# --------------------------------------------------------------------------------------------------
def build_orb_shim(tmp):
    exe = os.path.join(tmp, "orb_main")
    cmd = ["g++", "-std=c++14", "-O1", "-Wall", "-I", os.path.join(ROOT, "tests", "shim"), "-I", os.path.join(ROOT, "include"),
           SHIM, os.path.join(ROOT, "tests", "shim", "orb_main.cpp"), "-L", LIBDIR, "-lesfm_match", f"-Wl,-rpath,{LIBDIR}", "-o", exe]
    subprocess.check_call(cmd)
    return exe


def test_orb_shim_compiles_and_links(tmp_path):
    assert os.path.exists(build_orb_shim(str(tmp_path)))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["odd_bgr_2000", "vga_500"])
def test_orb_shim_equals_cv2_golden(tmp_path, name):
    from orb_util import CASES, image
    gold = np.load(os.path.join(ROOT, "tests", "golden", "orb_extract.npz"))
    seed, h, w, shapes, bgr, nf = CASES[name]
    img = image(seed, h, w, shapes, bgr)
    exe = build_orb_shim(str(tmp_path))
    raw_path, out = os.path.join(str(tmp_path), "img.bin"), os.path.join(str(tmp_path), "kp.bin")
    with open(raw_path, "wb") as f:
        f.write(np.ascontiguousarray(img).tobytes())
    subprocess.check_call([exe, raw_path, str(h), str(w), "3" if bgr else "1", str(nf), out], stdout=subprocess.DEVNULL)
    raw = open(out, "rb").read()
    n = int(np.frombuffer(raw, np.int32, 1, 0)[0])
    rec = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])
    kp = np.frombuffer(raw, rec, n, 4)
    desc = np.frombuffer(raw, np.uint8, n * 32, 4 + n * rec.itemsize).reshape(n, 32)
    ref = gold[name + "_kp"]
    assert n == len(ref) and (kp["class_id"] == -1).all()
    for fld in ref.dtype.names:
        assert np.array_equal(kp[fld], ref[fld]), fld
    assert np.array_equal(desc, gold[name + "_desc"])

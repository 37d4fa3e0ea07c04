"""The host arithmetic of the ORB extractor (easysfm_b200/csrc/orb_host.h, compiled with g++: tests/host/orb_host.cpp) against the numpy
oracle, without a GPU: level geometry, per-level budgets, the retainBest replay, key-point coordinates and the sampling frame."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import orb_oracle as oo

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def lib():
    out = os.path.join(HERE, "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "liborbhost.so")
    src = os.path.join(HERE, "host", "orb_host.cpp")
    hdr = os.path.join(HERE, "..", "easysfm_b200", "csrc", "orb_host.h")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, src])
    lib = ctypes.CDLL(so)
    lib.obh_retain_best.restype = ctypes.c_int
    lib.obh_harris_scale4.restype = ctypes.c_float
    return lib


def ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def test_level_geometry(lib):
    rng = np.random.default_rng(1)
    sizes = [(640, 480), (700, 297), (545, 525), (194, 537), (1920, 1080), (16383, 16383), (8, 8)] + [tuple(int(v) for v in rng.integers(8, 5000, 2)) for _ in range(400)]
    ref_scales = np.array(oo.level_scales(), np.float32)
    for cols, rows in sizes:
        scale, w, h = np.zeros(8, np.float32), np.zeros(8, np.int32), np.zeros(8, np.int32)
        lib.obh_levels(cols, rows, ptr(scale), ptr(w), ptr(h))
        assert np.array_equal(scale, ref_scales)
        for l in range(8):      # build_pyramid's sizes without resizing anything
            inv = np.float32(1.0) / ref_scales[l]
            assert (w[l], h[l]) == ((cols, rows) if l == 0 else (oo.c_round(np.float32(cols) * inv), oo.c_round(np.float32(rows) * inv))), (cols, rows, l)


def test_features_per_level(lib):
    for n in [0, 1, 7, 8, 100, 499, 500, 2000, 5000, 8000, 20000, 123457]:
        out = np.zeros(8, np.int32)
        lib.obh_features_per_level(n, ptr(out))
        assert out.tolist() == oo.features_per_level(n), n


def test_retain_best_replay(lib):
    rng = np.random.default_rng(2)
    for trial in range(200):
        n = int(rng.integers(1, 3000))
        # integer-valued responses (FAST scores) tie heavily; floats (Harris) rarely
        r = (rng.integers(20, 60, n) if trial % 2 else rng.random(n)).astype(np.float32)
        k = int(rng.integers(0, n + 10))
        out = np.zeros(n, np.int32)
        m = lib.obh_retain_best(ptr(r), n, k, ptr(out))
        assert out[:m].tolist() == oo.retain_best(r, k).tolist(), (trial, n, k)
        if 0 < k < n:       # the kept SET is {response >= the k-th largest}, whatever the order
            cut = np.sort(r)[::-1][k - 1]
            assert sorted(out[:m].tolist()) == np.flatnonzero(r >= cut).tolist()


def test_keypoint_and_sampling_frame(lib):
    rng = np.random.default_rng(3)
    scales = oo.level_scales()
    for _ in range(3000):
        l = int(rng.integers(0, 8))
        x, y = int(rng.integers(31, 3000)), int(rng.integers(31, 3000))
        ang = np.float32(rng.random() * 360.0)
        f, i = np.zeros(5, np.float32), np.zeros(2, np.int32)
        lib.obh_keypoint(x, y, l, ctypes.c_float(float(ang)), ptr(f), ptr(i))
        kx, ky = np.float32(x) * scales[l], np.float32(y) * scales[l]
        inv = np.float32(1.0) / scales[l]
        a = np.float32(ang) * np.float32(np.pi / 180.0)
        want = (kx, ky, np.float32(31) * scales[l], np.float32(np.cos(np.float64(a))), np.float32(np.sin(np.float64(a))))
        assert tuple(f) == want, (x, y, l, ang)
        assert tuple(i) == (oo.c_round(kx * inv), oo.c_round(ky * inv))
        assert tuple(i) == (x, y)       # scaling up and back never moves a corner off its pixel


def test_constants(lib):
    k = np.zeros(7, np.float32)
    lib.obh_gaussian_kernel_7(ptr(k))
    assert np.array_equal(k, oo.gaussian_kernel_7())
    s = np.float32(1.0) / (np.float32(28) * np.float32(255.0))
    assert np.float32(lib.obh_harris_scale4()) == s * s * s * s

"""Generates tests/golden/orb_extract.npz by running the REFERENCE's ORB calls (cv::ORB::create(max_num) -> detect -> compute, exactly as
cpp_code/src/feature_matching.cpp:16-22 issues them) through cv2 4.13.0 in the build container.

    python tests/golden/make_golden_orb.py

Stored per case of tests/orb_util.py:CASES: the key points (x, y, size, angle, response, octave) in cv2's output order and the 32-byte
descriptors.  The images are regenerated from their seeds by the tests, which need neither cv2 nor /root/reference."""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from orb_util import CASES, image  # noqa: E402

KP = np.dtype([("x", "f4"), ("y", "f4"), ("size", "f4"), ("angle", "f4"), ("response", "f4"), ("octave", "i4")])
out = {}
for name, (seed, h, w, shapes, bgr, nf) in CASES.items():
    img = image(seed, h, w, shapes, bgr)
    det = cv2.ORB_create(nf)
    ext = cv2.ORB_create(nf)
    kps = det.detect(img, None)
    kps, desc = ext.compute(img, kps)
    arr = np.zeros(len(kps), KP)
    for i, k in enumerate(kps):
        arr[i] = (k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave)
    if desc is None:
        desc = np.zeros((0, 32), np.uint8)
    out[name + "_kp"], out[name + "_desc"] = arr, desc
    print(name, img.shape, len(kps), np.bincount(arr["octave"], minlength=8))
np.savez_compressed(os.path.join(HERE, "orb_extract.npz"), **out)

# A real photograph: the central 384 x 512 window of frame 0004 of the reference's bundled fountain set (test_data/images_25), as gray (what
# ORB reduces the image to first), with cv2's output for it.  Stored with its pixels because /root/reference does not exist where the GPU
# tests run.
src = "/root/reference/test_data/images_25/0004.png"
if os.path.exists(src):
    gray = cv2.cvtColor(cv2.imread(src), cv2.COLOR_BGR2GRAY)
    h, w = gray.shape
    crop = np.ascontiguousarray(gray[(h - 384) // 2:(h - 384) // 2 + 384, (w - 512) // 2:(w - 512) // 2 + 512])
    kps = cv2.ORB_create(2000).detect(crop, None)
    kps, desc = cv2.ORB_create(2000).compute(crop, kps)
    arr = np.zeros(len(kps), KP)
    for i, k in enumerate(kps):
        arr[i] = (k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave)
    np.savez_compressed(os.path.join(HERE, "orb_fountain.npz"), image=crop, kp=arr, desc=desc, max_features=np.int32(2000))
    print("fountain crop", crop.shape, len(kps), np.bincount(arr["octave"], minlength=8))

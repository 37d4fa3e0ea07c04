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

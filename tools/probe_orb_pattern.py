"""Recovers the 256 x 2 sampling pairs of cv2's ORB descriptor (WTA_K = 2) by probing cv2.ORB.compute with impulse images.

The learned rBRIEF test pattern is part of OpenCV (features2d, orb.cpp: bit_pattern_31_), which is a dependency of the reference
(cpp_code/src/feature_matching.cpp:16-22 calls cv::ORB::create(max_num)->detect/compute) and is not under /root/reference.  The pattern is
not exposed through cv2 either, so it is measured: one key point at angle 0 on octave 0, a single white pixel on black (family A) or a
single black pixel on white (family B) at every offset P of a 39 x 39 window.  Bit k of the descriptor is blur(p0_k) < blur(p1_k), with
blur = the 7 x 7 sigma 2 fixed-point Gaussian (oracle/orb_oracle.py:gaussian_blur_7x7, pinned to cv2.GaussianBlur); the pair that reproduces
both response maps exactly is the pattern entry.  Output: easysfm_b200/csrc/orb_pattern.inc (C initialiser) and oracle/orb_pattern.npy.

Run in a container that has cv2 (4.13.0 here); the outputs are committed.
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
K7 = np.array([18, 34, 48, 56, 48, 34, 18], np.int64)
R = 19          # probe window radius
C = 60          # key point position
N = 2 * C + 1


def blur(img):
    a = np.pad(img.astype(np.int64), 3, mode="reflect")
    h, w = img.shape
    hp = sum(K7[i] * a[:, i:i + w] for i in range(7))
    v = sum(K7[i] * hp[i:i + h, :] for i in range(7))
    return (v + (1 << 15)) >> 16


def main():
    orb = cv2.ORB_create(500)
    kp = [cv2.KeyPoint(float(C), float(C), 31.0, 0.0, 1.0, 0, -1)]
    bits = np.zeros((2, 2 * R + 1, 2 * R + 1, 256), np.uint8)
    for fam in range(2):
        for dy in range(-R, R + 1):
            for dx in range(-R, R + 1):
                img = np.full((N, N), 255 * fam, np.uint8)
                img[C + dy, C + dx] = 255 * (1 - fam)
                k2, d = orb.compute(img, kp)
                assert len(k2) == 1
                bits[fam, dy + R, dx + R] = np.unpackbits(d[0], bitorder="little")
    # impulse responses of the blur: value at q for an impulse at P is resp[fam][q - P]
    base = np.zeros((15, 15), np.uint8); base[7, 7] = 255
    respA = blur(base)                                    # white pixel on black
    respB = blur(255 - base)                              # black pixel on white

    def val(resp, q, P):
        d = (q[0] - P[0] + 7, q[1] - P[1] + 7)
        if 0 <= d[0] < 15 and 0 <= d[1] < 15:
            return resp[d[0], d[1]]
        return resp[0, 0]

    pattern = np.zeros((256, 4), np.int32)                # x0 y0 x1 y1
    for k in range(256):
        SA = np.argwhere(bits[0, :, :, k]) - R            # (dy, dx) where p1 is the brighter one
        SB = np.argwhere(bits[1, :, :, k]) - R            # ... where p0 is the darker one
        assert len(SA) and len(SB), k

        def cands(S):
            lo = S.max(0) - 3; hi = S.min(0) + 3
            return [(y, x) for y in range(lo[0], hi[0] + 1) for x in range(lo[1], hi[1] + 1)]
        found = []
        for p1 in cands(SA):
            for p0 in cands(SB):
                ok = True
                for fam, resp in ((0, respA), (1, respB)):
                    for dy in range(-R, R + 1):
                        for dx in range(-R, R + 1):
                            want = bits[fam, dy + R, dx + R, k]
                            got = val(resp, p0, (dy, dx)) < val(resp, p1, (dy, dx))
                            if got != want:
                                ok = False; break
                        if not ok: break
                    if not ok: break
                if ok:
                    found.append((p0, p1))
        assert len(found) == 1, (k, found)
        (y0, x0), (y1, x1) = found[0]
        pattern[k] = (x0, y0, x1, y1)
    assert np.abs(pattern).max() <= 15
    np.save(os.path.join(ROOT, "oracle", "orb_pattern.npy"), pattern.astype(np.int8))
    with open(os.path.join(ROOT, "easysfm_b200", "csrc", "orb_pattern.inc"), "w") as f:
        f.write("// x0, y0, x1, y1 of the 256 rBRIEF tests (31 x 31 patch) as cv2 %s applies them; written by tools/probe_orb_pattern.py\n" % cv2.__version__)
        for k in range(256):
            f.write("%d,%d,%d,%d,%s" % (*pattern[k], "\n" if k % 8 == 7 else " "))
    print("pattern recovered; |coord| max", np.abs(pattern).max())


if __name__ == "__main__":
    sys.exit(main())

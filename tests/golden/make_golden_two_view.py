"""Generates tests/golden/two_view.npz by running the REFERENCE's calls of the two-view step (cv2.findEssentialMat / cv2.recoverPose /
cv2.triangulatePoints exactly as cpp_code/src/estimate_motion.cpp:48-49, :65, :262 issue them) in the build container (cv2 4.13.0).

    python tests/golden/make_golden_two_view.py

Stored: the camera matrix; 24 five-point sets with EVERY solution cv2's minimal solver returns for them (findEssentialMat on exactly five
points returns the 3k x 3 stack); 4 whole scenes with cv2's essential matrix, inlier mask, recoverPose output and the getDepthFast value
computed from cv2.triangulatePoints.  The tests need neither cv2 nor /root/reference."""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from two_view_util import scene  # noqa: E402

out = {}
k = 0
for s in range(40):
    K, x1, x2, _, _ = scene(5, 0.3, 0.0, 500 + s)
    E, _ = cv2.findEssentialMat(x1, x2, K, cv2.RANSAC, 0.99, 1.0)
    if E is None or k >= 24:
        continue
    out[f"five_x1_{k}"], out[f"five_x2_{k}"], out[f"five_E_{k}"] = x1, x2, E
    k += 1
out["n_five"] = np.int32(k)
out["K"] = K
for k, (n, noise, outl) in enumerate([(400, 0.3, 0.2), (600, 0.5, 0.4), (300, 1.0, 0.3), (500, 0.2, 0.6)]):
    K, x1, x2, R, t = scene(n, noise, outl, 700 + k)
    E, mask = cv2.findEssentialMat(x1, x2, K, cv2.RANSAC, 0.99, 1.0)
    good, Rcv, tcv, pmask = cv2.recoverPose(E, x1, x2, K, mask=mask.copy())
    sel = np.flatnonzero(mask.ravel())
    n1 = np.c_[(x1[sel, 0] - K[0, 2]) / K[0, 0], (x1[sel, 1] - K[1, 2]) / K[1, 1]].astype(np.float64)
    n2 = np.c_[(x2[sel, 0] - K[0, 2]) / K[0, 0], (x2[sel, 1] - K[1, 2]) / K[1, 1]].astype(np.float64)
    Q = cv2.triangulatePoints(np.eye(3, 4), np.c_[Rcv, tcv], n1.T, n2.T)
    X = (Q[:3] / Q[3]).T
    out[f"sc_x1_{k}"], out[f"sc_x2_{k}"], out[f"sc_E_{k}"], out[f"sc_mask_{k}"] = x1, x2, E, mask.ravel().astype(np.uint8)
    out[f"sc_R_{k}"], out[f"sc_t_{k}"], out[f"sc_good_{k}"] = Rcv, tcv.ravel(), np.int32(good)
    out[f"sc_depth_{k}"] = np.float64(np.linalg.norm(X, axis=1).mean())
out["n_scene"] = np.int32(4)
np.savez_compressed(os.path.join(HERE, "two_view.npz"), **out)
print("wrote two_view.npz:", len(out), "arrays,", os.path.getsize(os.path.join(HERE, "two_view.npz")), "bytes")

"""oracle/cv2_oracle.py -- TEST INFRASTRUCTURE ONLY: the reference's matcher calls, run through cv2.

The reference's arithmetic for this path lives in OpenCV's BFMatcher, which the reference calls at
  - cpp_code/src/feature_matching.cpp:74,80   DescriptorMatcher::create("BruteForce-Hamming")->knnMatch(.., 2)
  - python_code/feature_match.py:33-39         cv2.BFMatcher(cv2.NORM_L2, crossCheck=False).knnMatch(k=2) + ratio
  - python_code/feature_match.py:26-27         cv2.BFMatcher(cv2.NORM_L2, crossCheck=True).match
The functions below issue exactly those calls on caller-supplied descriptor arrays (the reference
file itself cannot run end to end here: cv2.xfeatures2d / SURF is absent) and compose
them the way SURVEY.md §8c defines ("2-NN + ratio + cross-check" exists nowhere in the reference as
one call, F3).  cv2 (opencv-python-headless 4.13.0) is part of the image, here and on the GPU box.
"""
from __future__ import annotations

import numpy as np

from . import DMATCH_DTYPE


def _norm(Q):
    import cv2
    if Q.dtype == np.float32:
        return cv2.NORM_L2
    if Q.dtype == np.uint8:
        return cv2.NORM_HAMMING
    raise TypeError("float32 or uint8 descriptors only")


def knn2(Q, T):
    """cv2.BFMatcher(norm).knnMatch(Q, T, k=2) -> (idx[nq,2], dist[nq,2]); -1/inf where absent."""
    import cv2
    nq = Q.shape[0]
    idx = np.full((nq, 2), -1, np.int32)
    dist = np.full((nq, 2), np.inf, np.float32)
    if nq == 0 or T.shape[0] == 0:
        return idx, dist
    res = cv2.BFMatcher(_norm(Q), crossCheck=False).knnMatch(np.ascontiguousarray(Q), np.ascontiguousarray(T), k=2)
    for q, lst in enumerate(res):
        for r, m in enumerate(lst[:2]):
            idx[q, r] = m.trainIdx
            dist[q, r] = m.distance
    return idx, dist


def match(Q, T, ratio: float, cross_check: bool):
    """The composite: forward knn-2 + double-precision ratio + optional reverse knn-1 mutual check."""
    import cv2
    out = []
    nq, nt = Q.shape[0], T.shape[0]
    # ratio = +inf: no ratio test (include/esfm_match.h) -- one neighbour is enough and `d1 < inf * d2` is NOT evaluated
    # (inf * 0 is NaN and would drop the exact duplicates that BFMatcher(crossCheck=True).match keeps)
    no_ratio = ratio == float("inf")
    if nq == 0 or nt < (1 if no_ratio else 2):
        return np.zeros(0, DMATCH_DTYPE)
    Q = np.ascontiguousarray(Q)
    T = np.ascontiguousarray(T)
    norm = _norm(Q)
    fwd = cv2.BFMatcher(norm, crossCheck=False).knnMatch(Q, T, k=1 if no_ratio else 2)
    rev = None
    if cross_check:
        r = cv2.BFMatcher(norm, crossCheck=False).knnMatch(T, Q, k=1)
        rev = np.array([lst[0].trainIdx if lst else -1 for lst in r], np.int64)
    for q, lst in enumerate(fwd):
        if len(lst) < (1 if no_ratio else 2):
            continue
        m = lst[0]
        # Python floats are doubles: the same arithmetic as `float < double * float` in C++.
        if not no_ratio and not (float(m.distance) < float(ratio) * float(lst[1].distance)):
            continue
        if rev is not None and rev[m.trainIdx] != q:
            continue
        out.append((q, m.trainIdx, 0, m.distance))
    return np.array(out, DMATCH_DTYPE) if out else np.zeros(0, DMATCH_DTYPE)


def mutual_nn(Q, T):
    """cv2.BFMatcher(norm, crossCheck=True).match(Q, T), in ascending queryIdx."""
    import cv2
    if Q.shape[0] == 0 or T.shape[0] == 0:
        return np.zeros(0, DMATCH_DTYPE)
    ms = cv2.BFMatcher(_norm(Q), crossCheck=True).match(np.ascontiguousarray(Q), np.ascontiguousarray(T))
    out = sorted(((m.queryIdx, m.trainIdx, 0, m.distance) for m in ms), key=lambda x: x[0])
    return np.array(out, DMATCH_DTYPE) if out else np.zeros(0, DMATCH_DTYPE)


def time_pair(Q, T, cross_check: bool, repeats: int = 3) -> float:
    """Best-of-N wall seconds of the reference's CPU calls for one image pair (bench.py baseline)."""
    import time
    import cv2
    norm = _norm(Q)
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        cv2.BFMatcher(norm, crossCheck=False).knnMatch(Q, T, k=2)
        if cross_check:
            cv2.BFMatcher(norm, crossCheck=False).knnMatch(T, Q, k=1)
        best = min(best, time.perf_counter() - t0)
    return best

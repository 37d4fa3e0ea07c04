"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the EasySFM matching hot path (see bf_oracle.c header for the reference
file:line map) plus the OpenCV composite that pins it (cv2_oracle.py).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package; the product package easysfm_b200 never does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

KIND_L2 = 0
KIND_HAMMING = 1

DMATCH_DTYPE = np.dtype(
    [("queryIdx", "<i4"), ("trainIdx", "<i4"), ("imgIdx", "<i4"), ("distance", "<f4")]
)


def build(force: bool = False) -> str:
    """Compile oracle/bf_oracle.c -> oracle/liboracle.so (gcc, a second)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "bf_oracle.c")
    src2 = os.path.join(_HERE, "tracks_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(src2)):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(build())
        vp, ip, fp = ctypes.c_void_p, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_float)
        lib.oracle_knn2.argtypes = [ctypes.c_int, vp, ctypes.c_int, vp, ctypes.c_int, ctypes.c_int, ip, fp]
        lib.oracle_knn2.restype = ctypes.c_int
        lib.oracle_match.argtypes = [ctypes.c_int, vp, ctypes.c_int, vp, ctypes.c_int, ctypes.c_int,
                                     ctypes.c_double, ctypes.c_int, vp, ctypes.POINTER(ctypes.c_int)]
        lib.oracle_match.restype = ctypes.c_int
        lib.oracle_mutual_nn.argtypes = [ctypes.c_int, vp, ctypes.c_int, vp, ctypes.c_int, ctypes.c_int,
                                         vp, ctypes.POINTER(ctypes.c_int)]
        lib.oracle_mutual_nn.restype = ctypes.c_int
        i64p, u8p, dp = ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_uint8), ctypes.POINTER(ctypes.c_double)
        lib.oracle_tracks_build.argtypes = [ctypes.c_int, ip, i64p, ip, vp, ip, u8p, u8p]
        lib.oracle_tracks_build.restype = ctypes.c_int64
        lib.oracle_find_init_pair.argtypes = [ctypes.c_int, ctypes.c_int64, u8p, dp, ctypes.c_int, ctypes.c_double,
                                              ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int), dp, i64p]
        lib.oracle_find_init_pair.restype = ctypes.c_int
        lib.oracle_find_next_frame.argtypes = [ctypes.c_int, ctypes.c_int64, u8p, u8p, ip, ctypes.c_int64, ctypes.POINTER(ctypes.c_int)]
        lib.oracle_find_next_frame.restype = ctypes.c_int
        lib.oracle_set_threads.argtypes = [ctypes.c_int]
        lib.oracle_get_threads.restype = ctypes.c_int
        _LIB = lib
    return _LIB


def set_threads(n: int) -> None:
    """0 = all online cores."""
    _lib().oracle_set_threads(int(n))


def get_threads() -> int:
    return int(_lib().oracle_get_threads())


def _prep(Q, T):
    Q = np.ascontiguousarray(Q)
    T = np.ascontiguousarray(T)
    if Q.dtype == np.float32 and T.dtype == np.float32:
        kind = KIND_L2
    elif Q.dtype == np.uint8 and T.dtype == np.uint8:
        kind = KIND_HAMMING
    else:
        raise TypeError("descriptors must both be float32 (L2) or uint8 (Hamming)")
    if Q.ndim != 2 or T.ndim != 2 or (Q.shape[0] and T.shape[0] and Q.shape[1] != T.shape[1]):
        raise ValueError("descriptor banks must be 2-D with equal width")
    cols = Q.shape[1] if Q.shape[0] else T.shape[1]
    return kind, Q, T, int(cols)


def knn2(Q, T):
    """Two nearest train rows per query row: (idx[nq,2] int32, dist[nq,2] float32)."""
    kind, Q, T, cols = _prep(Q, T)
    nq, nt = Q.shape[0], T.shape[0]
    idx = np.full((nq, 2), -1, np.int32)
    dist = np.full((nq, 2), np.inf, np.float32)
    rc = _lib().oracle_knn2(kind, Q.ctypes.data, nq, T.ctypes.data, nt, cols,
                            idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                            dist.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
    if rc:
        raise RuntimeError(f"oracle_knn2 failed: {rc}")
    return idx, dist


def match(Q, T, ratio: float, cross_check: bool):
    """2-NN + ratio (double) + optional mutual cross-check; structured array, ascending queryIdx."""
    kind, Q, T, cols = _prep(Q, T)
    nq, nt = Q.shape[0], T.shape[0]
    out = np.zeros(max(nq, 1), DMATCH_DTYPE)
    n = ctypes.c_int(0)
    rc = _lib().oracle_match(kind, Q.ctypes.data, nq, T.ctypes.data, nt, cols, float(ratio),
                             int(bool(cross_check)), out.ctypes.data, ctypes.byref(n))
    if rc:
        raise RuntimeError(f"oracle_match failed: {rc}")
    return out[: n.value].copy()


def mutual_nn(Q, T):
    kind, Q, T, cols = _prep(Q, T)
    nq, nt = Q.shape[0], T.shape[0]
    out = np.zeros(max(nq, 1), DMATCH_DTYPE)
    n = ctypes.c_int(0)
    rc = _lib().oracle_mutual_nn(kind, Q.ctypes.data, nq, T.ctypes.data, nt, cols, out.ctypes.data, ctypes.byref(n))
    if rc:
        raise RuntimeError(f"oracle_mutual_nn failed: {rc}")
    return out[: n.value].copy()


# ---- track building / co-visibility (oracle/tracks_oracle.c: literal restatement of sfm.cpp:140-217, feature_matching.cpp:160-268) ----
def tracks_build(keypoints_per_frame, inliers_per_pair):
    """inliers_per_pair: DMATCH arrays in the reference's loop order (i ascending, j < i).  -> (ids per frame [list of int32
    arrays], has_match per frame, dense track matrix uint8 [n_frames, total_keypoints], n_unique_points)."""
    kp = np.ascontiguousarray(keypoints_per_frame, dtype=np.int32)
    n = len(kp)
    cnt = np.array([len(m) for m in inliers_per_pair], np.int32)
    assert len(cnt) == n * (n - 1) // 2
    off = np.zeros(len(cnt) + 1, np.int64)
    np.cumsum(cnt, out=off[1:])
    inl = np.concatenate([np.ascontiguousarray(m, dtype=DMATCH_DTYPE) for m in inliers_per_pair]) if len(cnt) and off[-1] else np.zeros(1, DMATCH_DTYPE)
    total = int(kp.sum())
    ids = np.zeros(max(total, 1), np.int32)
    has = np.zeros(max(total, 1), np.uint8)
    track = np.zeros((n, max(total, 1)), np.uint8)
    i32, i64, u8 = ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_uint8)
    cnt_c = np.ascontiguousarray(cnt if len(cnt) else np.zeros(1, np.int32))
    npts = _lib().oracle_tracks_build(n, kp.ctypes.data_as(i32), off.ctypes.data_as(i64), cnt_c.ctypes.data_as(i32), inl.ctypes.data,
                                      ids.ctypes.data_as(i32), has.ctypes.data_as(u8), track.ctypes.data_as(u8))
    o = np.concatenate([[0], np.cumsum(kp)])
    return [ids[o[f]:o[f + 1]].copy() for f in range(n)], [has[o[f]:o[f + 1]].copy() for f in range(n)], track, int(npts)


def find_init_pair(track, appro_depth, min_track_num_init=100, max_depth_baseline_ratio_init=50.0):
    """-> (found, frame_1, frame_2, depth_init, best_score) of feature_matching.cpp:160-233 on the dense track matrix."""
    track = np.ascontiguousarray(track, dtype=np.uint8)
    n, npts = track.shape
    depth = np.ascontiguousarray(appro_depth, dtype=np.float64)
    if len(depth) == 0:
        depth = np.zeros(1, np.float64)
    f1, f2, d, best = ctypes.c_int(), ctypes.c_int(), ctypes.c_double(), ctypes.c_int64()
    ok = _lib().oracle_find_init_pair(n, npts, track.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), depth.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                      int(min_track_num_init), float(max_depth_baseline_ratio_init), ctypes.byref(f1), ctypes.byref(f2),
                                      ctypes.byref(d), ctypes.byref(best))
    return bool(ok), f1.value, f2.value, d.value, best.value


def find_next_frame(track, frames_to_process, point_ids, next_frame=-1):
    """-> (next_frame, common points) of feature_matching.cpp:235-268 (next_frame unchanged when nothing is seen)."""
    track = np.ascontiguousarray(track, dtype=np.uint8)
    n, npts = track.shape
    tp = np.ascontiguousarray(frames_to_process, dtype=np.uint8)
    pid = np.ascontiguousarray(point_ids, dtype=np.int32)
    nf = ctypes.c_int(int(next_frame))
    common = _lib().oracle_find_next_frame(n, npts, track.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), tp.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)),
                                           (pid if len(pid) else np.zeros(1, np.int32)).ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), len(pid),
                                           ctypes.byref(nf))
    return nf.value, int(common)
